"""Checkpoint layout of the reference's `M3PTransformerModel` (m3p_transformer.py:L613-728) as data.

Only `jointfwd` is on the hot path; the decoder / generation / pre-training branches are not
executed, but their parameters live in the state_dict and must round-trip through
`load_state_dict` / `state_dict` / `from_pretrained` (479 keys at 12 layers, 267
tensors that never receive a gradient).  The module tree is therefore built from this table as
plain parameter holders (nn.Linear / nn.LayerNorm / nn.Embedding leaves with the reference's names,
so torch.nn.utils.prune can re-parametrise them by name as train_task_sft.py:L134-215 does).
"""
from __future__ import annotations

import torch.nn as nn

N_MAX_POSITIONS = 514          # m3p_transformer.py:L16


def encoder_leaves(cfg):
    """Ordered (path, kind, args) for `bert.encoder.*`.  kind: 'emb' (num, dim, padding_idx),
    'lin' (in, out), 'ln' (dim,)."""
    d = int(getattr(cfg, "emb_dim", getattr(cfg, "hidden_size", 768)))
    L = int(getattr(cfg, "n_layers", 12))
    n_words = int(getattr(cfg, "n_words", 250002))
    n_langs = int(getattr(cfg, "n_langs", 100))
    refine = int(getattr(cfg, "refine_layers", 6))
    pad = int(getattr(cfg, "pad_index", 1))
    hid = 4 * d
    out = [("position_embeddings", "emb", (N_MAX_POSITIONS, d, None))]
    if n_langs > 1:
        out.append(("cross_lang_embeddings", "emb", (n_langs, d, None)))
    out += [("embeddings", "emb", (n_words, d, pad)), ("layer_norm_emb", "ln", (d,)),
            ("image_embeddings.image_embeddings", "lin", (2048, d)),
            ("image_embeddings.image_distbution_embeddings", "lin", (1600, d)),
            ("image_embeddings.image_location_embeddings", "lin", (5, d)),
            ("image_embeddings.LayerNorm", "ln", (d,))]
    for i in range(refine):
        p = "refine_embeddings.layers.%d." % i
        for j in range(3):
            out.append((p + "self_attn.linears.%d" % j, "lin", (d, d)))
        out.append((p + "self_attn.aoa_layer.0", "lin", (2 * d, 2 * d)))
        out += [(p + "feed_forward.lin1", "lin", (d, hid)), (p + "feed_forward.lin2", "lin", (hid, d)),
                (p + "sublayer.0.norm", "ln", (d,)), (p + "sublayer.1.norm", "ln", (d,))]
    out.append(("refine_embeddings.norm", "ln", (d,)))
    out += [("cross_alignment.att_weight_c", "lin", (d, 1)), ("cross_alignment.att_weight_q", "lin", (d, 1)),
            ("cross_alignment.att_weight_cq", "lin", (d, 1)), ("cross_alignment.align_output", "lin", (d, d)),
            ("cross_alignment.layer_norm", "ln", (d,))]
    for i in range(L):
        for n in ("q_lin", "k_lin", "v_lin", "out_lin"):
            out.append(("attentions.%d.%s" % (i, n), "lin", (d, d)))
    for i in range(L):
        out.append(("layer_norm1.%d" % i, "ln", (d,)))
    for i in range(L):
        out += [("ffns.%d.lin1" % i, "lin", (d, hid)), ("ffns.%d.lin2" % i, "lin", (hid, d))]
    for i in range(L):
        out.append(("layer_norm2.%d" % i, "ln", (d,)))
    for i in range(L):
        out.append(("layer_norm15.%d" % i, "ln", (d,)))
    for i in range(L):
        for n in ("q_lin", "k_lin", "v_lin", "out_lin"):
            out.append(("encoder_attn.%d.%s" % (i, n), "lin", (d, d)))
    for i in range(2):
        p = "latent_transforms.%d." % i
        out += [(p + "x_to_mu", "lin", (d, d)), (p + "x_to_logvar", "lin", (d, d)), (p + "out_dense", "lin", (2 * d, d))]
    for i in range(2):
        p = "original_transforms.%d." % i
        out += [(p + "dense", "lin", (d, d)), (p + "dense_mu", "lin", (d, d)), (p + "LayerNorm", "ln", (d,))]
    out += [("pooled_layer.dense", "lin", (d, d)), ("seq_relationship", "lin", (d, 1)),
            ("pooled_layer2.dense", "lin", (d, d)), ("seq_relationship2", "lin", (d, 1)),
            ("mrfr_dense", "lin", (d, 2048)), ("transformer_obj.dense", "lin", (d, d)),
            ("transformer_obj.LayerNorm", "ln", (d,))]
    return out


# parameters under bert.encoder that jointfwd touches (everything else never receives a gradient)
def live_prefixes(cfg):
    L = int(getattr(cfg, "n_layers", 12))
    live = ["position_embeddings", "embeddings", "layer_norm_emb", "image_embeddings.image_embeddings",
            "image_embeddings.image_location_embeddings", "image_embeddings.LayerNorm", "pooled_layer.dense"]
    for i in range(L):
        live += ["attentions.%d." % i, "layer_norm1.%d" % i, "ffns.%d." % i, "layer_norm2.%d" % i]
    return live


class Holder(nn.Module):
    """anonymous container; children are attached with add_module under the reference's names
    (numeric names index like an nn.ModuleList)"""

    def __getitem__(self, i):
        return self._modules[str(i)]

    def __len__(self):
        return len(self._modules)


def build_encoder(cfg) -> nn.Module:
    root = Holder()
    for path, kind, a in encoder_leaves(cfg):
        parts = path.split(".")
        cur = root
        for part in parts[:-1]:
            if part not in cur._modules:
                cur.add_module(part, Holder())
            cur = cur._modules[part]
        if kind == "emb":
            leaf = nn.Embedding(a[0], a[1], padding_idx=a[2])
        elif kind == "lin":
            leaf = nn.Linear(a[0], a[1])
        else:
            leaf = nn.LayerNorm(a[0], eps=1e-12)
        cur.add_module(parts[-1], leaf)
    return root


def state_shapes(cfg, num_labels=1842, task="TASK15"):
    """{state_dict key: shape} of M3PForVLTasks, incl. the bert.pooler alias (encoders.py:L1029)."""
    d = int(getattr(cfg, "emb_dim", getattr(cfg, "hidden_size", 768)))
    C = int(getattr(cfg, "clf_hidden_size", 1536))
    out = {}
    for path, kind, a in encoder_leaves(cfg):
        k = "bert.encoder." + path
        if kind == "emb":
            out[k + ".weight"] = (a[0], a[1])
        elif kind == "lin":
            out[k + ".weight"] = (a[1], a[0])
            out[k + ".bias"] = (a[1],)
        else:
            out[k + ".weight"] = (a[0],)
            out[k + ".bias"] = (a[0],)
    out["bert.pooler.dense.weight"] = (d, d)
    out["bert.pooler.dense.bias"] = (d,)
    c = "clfs_dict.%s.logit_fc." % task
    out[c + "0.weight"] = (C, int(getattr(cfg, "pooler_size", d)))
    out[c + "0.bias"] = (C,)
    out[c + "2.weight"] = (C,)
    out[c + "2.bias"] = (C,)
    out[c + "3.weight"] = (num_labels, C)
    out[c + "3.bias"] = (num_labels,)
    return out
