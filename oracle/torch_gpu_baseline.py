"""Stock-PyTorch restatement of the UC2 / M3P step with the strongest library ops (F.linear -> cuBLAS,
F.scaled_dot_product_attention, F.layer_norm, F.gelu, fused torch.optim.AdamW), meant to run on the SAME B200 under
torch.autocast(bf16): the "what does plain PyTorch get on this GPU" bar of bench.py --impl torch_gpu.
BASELINE INFRASTRUCTURE ONLY - like the rest of oracle/, never imported by the product path.

Same arithmetic as oracle/uc2_oracle.py / m3p_oracle.py (which follow the reference line by line and are pinned to it by the
golden fixtures); tests/test_oracle_golden.py::test_torch_gpu_baseline_matches_oracle checks that in fp32 on the CPU.  It is
deliberately FASTER than the reference's own module graph (fused QKV weight, one joint sequence, SDPA instead of 8 bmm + 2
softmax + cat per layer, fused LayerNorm / GeLU): a generous baseline.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .uc2_oracle import position_ids


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


class FusedParams:
    """q|k|v weights and biases of every layer concatenated once (what any tuned PyTorch model does)"""

    def __init__(self, sd, family, n_layers):
        self.sd = sd
        self.qkv = []
        for l in range(n_layers):
            if family == "uc2":
                a = "bert.encoder.layer.%d.attention_self." % (2 * l)
                names = [a + "query", a + "key", a + "value"]
            else:
                a = "bert.encoder.attentions.%d." % l
                names = [a + "q_lin", a + "k_lin", a + "v_lin"]
            w = torch.cat([sd[n + ".weight"] for n in names], 0).detach().requires_grad_(sd[names[0] + ".weight"].requires_grad)
            b = torch.cat([sd[n + ".bias"] for n in names], 0).detach().requires_grad_(sd[names[0] + ".bias"].requires_grad)
            self.qkv.append((w, b))

    def parameters(self, family):
        skip = ("query.", "key.", "value.", "q_lin.", "k_lin.", "v_lin.")
        seen, out = set(), []
        for k, v in self.sd.items():
            if any(s in k for s in skip) and (".attention_self." in k or ".attentions." in k):
                continue
            if id(v) not in seen and v.requires_grad:
                seen.add(id(v))
                out.append(v)
        for w, b in self.qkv:
            out += [w, b]
        return out


def _attention(x, qkv_w, qkv_b, H, attn_bias, dropout_p):
    B, S, D = x.shape
    qkv = F.linear(x, qkv_w, qkv_b).view(B, S, 3, H, D // H).permute(2, 0, 3, 1, 4)
    ctx = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=attn_bias, dropout_p=dropout_p)
    return ctx.transpose(1, 2).reshape(B, S, D)


def uc2_forward(fp: FusedParams, cfg, input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask,
                task="TASK15", train=False):
    sd = fp.sd
    E = "bert.embeddings."
    eps = cfg.layer_norm_eps
    p_h = float(cfg.hidden_dropout_prob) if train else 0.0
    p_a = float(cfg.attention_probs_dropout_prob) if train else 0.0
    pos = position_ids(input_txt, cfg.pad_token_id)
    t = (F.embedding(input_txt, sd[E + "word_embeddings.weight"]) + F.embedding(pos, sd[E + "position_embeddings.weight"])
         + F.embedding(token_type_ids, sd[E + "new_token_type_embeddings.weight"]))
    t = F.dropout(_ln(t, sd[E + "LayerNorm.weight"], sd[E + "LayerNorm.bias"], eps), p_h, train)
    img = _ln(F.linear(input_imgs, sd[E + "image_embeddings.weight"], sd[E + "image_embeddings.bias"]),
              sd[E + "image_layer_norm.weight"], sd[E + "image_layer_norm.bias"], eps)
    loc = _ln(F.linear(image_loc, sd[E + "image_location_embeddings.weight"], sd[E + "image_location_embeddings.bias"]),
              sd[E + "image_location_layer_norm.weight"], sd[E + "image_location_layer_norm.bias"], eps)
    v = img.float() + loc.float() + sd[E + "new_token_type_embeddings.weight"][1]
    v = F.dropout(_ln(v, sd[E + "v_LayerNorm.weight"], sd[E + "v_LayerNorm.bias"], eps), p_h, train)
    x = torch.cat([t, v], dim=1)
    m = torch.cat([attention_mask, image_attention_mask], dim=1).to(x.dtype)
    bias = ((1.0 - m) * -10000.0)[:, None, None, :]
    H = cfg.num_attention_heads
    for l in range(len(cfg.tt_attn_sublayers)):
        a = "bert.encoder.layer.%d." % (2 * l)
        f = "bert.encoder.layer.%d." % (2 * l + 1)
        ctx = _attention(x, fp.qkv[l][0], fp.qkv[l][1], H, bias.to(torch.bfloat16) if x.is_cuda else bias, p_a)
        y = F.dropout(F.linear(ctx, sd[a + "attention_output.dense.weight"], sd[a + "attention_output.dense.bias"]), p_h, train)
        x = _ln(y + x, sd[a + "attention_output.LayerNorm.weight"], sd[a + "attention_output.LayerNorm.bias"], eps)
        h = F.gelu(F.linear(x, sd[f + "intermediate.dense.weight"], sd[f + "intermediate.dense.bias"]))
        y = F.dropout(F.linear(h, sd[f + "output.dense.weight"], sd[f + "output.dense.bias"]), p_h, train)
        x = _ln(y + x, sd[f + "output.LayerNorm.weight"], sd[f + "output.LayerNorm.bias"], eps)
    pooled = torch.relu(F.linear(x[:, 0], sd["bert.t_pooler.dense.weight"], sd["bert.t_pooler.dense.bias"]))
    pooled = F.dropout(pooled, 0.1, train)
    c = "clfs_dict.%s.logit_fc." % task
    h = _ln(F.gelu(F.linear(pooled, sd[c + "0.weight"], sd[c + "0.bias"])), sd[c + "2.weight"], sd[c + "2.bias"], eps)
    return F.linear(h, sd[c + "3.weight"], sd[c + "3.bias"]).float()


def m3p_forward(fp: FusedParams, cfg, input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask,
                task="TASK15", train=False):
    sd = fp.sd
    P = "bert.encoder."
    B, T = input_txt.shape
    R = input_imgs.shape[1]
    S = T + R
    H = cfg.n_heads if hasattr(cfg, "n_heads") else cfg.num_attention_heads
    n_layers = int(getattr(cfg, "n_layers", 12))
    p_h = float(getattr(cfg, "dropout", 0.1)) if train else 0.0
    p_a = float(getattr(cfg, "attention_dropout", 0.1)) if train else 0.0
    I = P + "image_embeddings."
    img = (F.linear(input_imgs, sd[I + "image_embeddings.weight"], sd[I + "image_embeddings.bias"]).float()
           + F.linear(image_loc, sd[I + "image_location_embeddings.weight"], sd[I + "image_location_embeddings.bias"]).float())
    img = F.dropout(_ln(img, sd[I + "LayerNorm.weight"], sd[I + "LayerNorm.bias"], 1e-12), p_h, train)
    txt = F.embedding(input_txt, sd[P + "embeddings.weight"])
    n = attention_mask.sum(dim=1) + image_attention_mask.sum(dim=1)
    valid = torch.arange(S, device=input_txt.device).unsqueeze(0) < n.unsqueeze(1)
    vf = valid.unsqueeze(-1).to(img.dtype)
    x = (torch.cat([img, txt], dim=1) + sd[P + "position_embeddings.weight"][:S].unsqueeze(0)) * vf
    x = F.dropout(_ln(x, sd[P + "layer_norm_emb.weight"], sd[P + "layer_norm_emb.bias"], 1e-12), p_h, train)
    amask = valid[:, None, None, :]
    for i in range(n_layers):
        A = P + "attentions.%d." % i
        ctx = _attention(x, fp.qkv[i][0], fp.qkv[i][1], H, amask, p_a)
        attn = F.dropout(F.linear(ctx, sd[A + "out_lin.weight"], sd[A + "out_lin.bias"]), p_h, train)
        x = _ln(x + attn, sd[P + "layer_norm1.%d.weight" % i], sd[P + "layer_norm1.%d.bias" % i], 1e-12)
        Fp = P + "ffns.%d." % i
        h = F.gelu(F.linear(x, sd[Fp + "lin1.weight"], sd[Fp + "lin1.bias"]))
        y = F.dropout(F.linear(h, sd[Fp + "lin2.weight"], sd[Fp + "lin2.bias"]), p_h, train)
        x = _ln(x + y, sd[P + "layer_norm2.%d.weight" % i], sd[P + "layer_norm2.%d.bias" % i], 1e-12) * vf
    pooled = torch.tanh(F.linear(x[:, 0], sd[P + "pooled_layer.dense.weight"], sd[P + "pooled_layer.dense.bias"]))
    pooled = F.dropout(pooled, 0.1, train)
    c = "clfs_dict.%s.logit_fc." % task
    eps = float(getattr(cfg, "layer_norm_eps", 1e-12))
    h = _ln(F.gelu(F.linear(pooled, sd[c + "0.weight"], sd[c + "0.bias"])), sd[c + "2.weight"], sd[c + "2.bias"], eps)
    return F.linear(h, sd[c + "3.weight"], sd[c + "3.bias"]).float()
