"""Oracle: UC2 (`BertForVLTasks` with uc2_base.json) forward, restated as a plain post-LN
BERT over cat(text, regions).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (reference, paths relative to /root/reference/volta/volta):
  embeddings.py:L157-170   create_position_ids_from_input_ids
  embeddings.py:L636-669   UC2Embeddings.forward (text LN; region 3-LN form)
  encoders.py:L978-995     additive masks (1-m)*-10000
  encoders.py:L229-359     BertGatedSelfAttention.forward  (scores/8 AFTER QK^T, then + mask)
  encoders.py:L399-425     BertGatedSelfOutput.forward     (dense, dropout, +residual, LN)
  encoders.py:L487-502     BertGatedIntermediate.forward   (dense, erf-GeLU)
  encoders.py:L542-567     BertGatedOutput.forward
  encoders.py:L603-608     BertTextPooler.forward          (ReLU(dense(h[:,0])))
  encoders.py:L788-815     SimpleClassifier                (Linear, GeLU, LN, Linear)
  encoders.py:L1202-1259   BertForVLTasks.forward

With every sub-layer in `shared_sublayers` and `single_ln_sublayers` and tt/tv/vt/vv all on
(uc2_base.json:L33-168) the gated bi-stream machinery is exactly this computation.

`mm_dtype=torch.bfloat16` reproduces the "bf16 oracle" of SURVEY.md Appendix A.5: the operands
of every F.linear / torch.matmul are rounded to bf16, accumulation and everything else stay in
the working precision.  Dropout is off (eval mode / p = 0): parity is defined without it.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _r(x, mm_dtype):
    return x if mm_dtype is None else x.to(mm_dtype).to(x.dtype)


ROUND_BWD = "bf16+bwd"      # mm_dtype value: bf16 operands in the forward AND the backward contractions


def linear(x, w, b, mm_dtype=None):
    if mm_dtype == ROUND_BWD:
        from .bf16_autograd import linear_bf16
        return linear_bf16(x, w, b)
    return F.linear(_r(x, mm_dtype), _r(w, mm_dtype), b)


def matmul(a, b, mm_dtype=None):
    if mm_dtype == ROUND_BWD:
        from .bf16_autograd import matmul_bf16
        return matmul_bf16(a, b)
    return torch.matmul(_r(a, mm_dtype), _r(b, mm_dtype))


def gelu(x):
    # encoders.py:L131-137
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps):
    # BertLayerNorm python fallback, encoders.py:L49-62 (biased variance, eps inside sqrt)
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + eps)) + b


def position_ids(input_ids, pad_id):
    # embeddings.py:L168-170
    mask = input_ids.ne(pad_id).int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + pad_id


def uc2_embeddings(sd, cfg, input_txt, input_imgs, image_loc, token_type_ids, mm_dtype=None):
    E = "bert.embeddings."
    eps = cfg.layer_norm_eps
    pos = position_ids(input_txt, cfg.pad_token_id)
    t = (sd[E + "word_embeddings.weight"][input_txt] + sd[E + "position_embeddings.weight"][pos]
         + sd[E + "new_token_type_embeddings.weight"][token_type_ids])
    t = layer_norm(t, sd[E + "LayerNorm.weight"], sd[E + "LayerNorm.bias"], eps)
    img = layer_norm(linear(input_imgs, sd[E + "image_embeddings.weight"], sd[E + "image_embeddings.bias"], mm_dtype),
                     sd[E + "image_layer_norm.weight"], sd[E + "image_layer_norm.bias"], eps)
    loc = layer_norm(linear(image_loc, sd[E + "image_location_embeddings.weight"],
                            sd[E + "image_location_embeddings.bias"], mm_dtype),
                     sd[E + "image_location_layer_norm.weight"], sd[E + "image_location_layer_norm.bias"], eps)
    v = img + loc + sd[E + "new_token_type_embeddings.weight"][1]
    v = layer_norm(v, sd[E + "v_LayerNorm.weight"], sd[E + "v_LayerNorm.bias"], eps)
    return t, v


def encoder_layer(sd, cfg, l, x, add_mask, mm_dtype=None):
    a = "bert.encoder.layer.%d." % (2 * l)
    f = "bert.encoder.layer.%d." % (2 * l + 1)
    eps = cfg.layer_norm_eps
    B, S, D = x.shape
    H = cfg.num_attention_heads
    dh = D // H

    def heads(y):
        return y.view(B, S, H, dh).permute(0, 2, 1, 3)

    q = heads(linear(x, sd[a + "attention_self.query.weight"], sd[a + "attention_self.query.bias"], mm_dtype))
    k = heads(linear(x, sd[a + "attention_self.key.weight"], sd[a + "attention_self.key.bias"], mm_dtype))
    v = heads(linear(x, sd[a + "attention_self.value.weight"], sd[a + "attention_self.value.bias"], mm_dtype))
    scores = matmul(q, k.transpose(-1, -2), mm_dtype) / math.sqrt(dh)
    scores = scores + add_mask[:, None, None, :]
    probs = torch.softmax(scores, dim=-1)
    ctx = matmul(probs, v, mm_dtype).permute(0, 2, 1, 3).reshape(B, S, D)
    x = layer_norm(linear(ctx, sd[a + "attention_output.dense.weight"], sd[a + "attention_output.dense.bias"], mm_dtype) + x,
                   sd[a + "attention_output.LayerNorm.weight"], sd[a + "attention_output.LayerNorm.bias"], eps)
    h = gelu(linear(x, sd[f + "intermediate.dense.weight"], sd[f + "intermediate.dense.bias"], mm_dtype))
    x = layer_norm(linear(h, sd[f + "output.dense.weight"], sd[f + "output.dense.bias"], mm_dtype) + x,
                   sd[f + "output.LayerNorm.weight"], sd[f + "output.LayerNorm.bias"], eps)
    return x


def uc2_forward(sd, cfg, input_txt, input_imgs, image_loc, token_type_ids, attention_mask,
                image_attention_mask, task="TASK15", mm_dtype=None, return_hidden=False):
    """logits [B, num_labels].  `sd` maps the reference's state_dict keys to tensors."""
    dt = sd["bert.embeddings.LayerNorm.weight"].dtype
    input_imgs = input_imgs.to(dt)
    image_loc = image_loc.to(dt)
    t, v = uc2_embeddings(sd, cfg, input_txt, input_imgs, image_loc, token_type_ids, mm_dtype)
    x = torch.cat([t, v], dim=1)
    m = torch.cat([attention_mask, image_attention_mask], dim=1).to(dt)
    add_mask = (1.0 - m) * -10000.0
    n_layers = len(cfg.tt_attn_sublayers)
    hidden = [x]
    for l in range(n_layers):
        x = encoder_layer(sd, cfg, l, x, add_mask, mm_dtype)
        hidden.append(x)
    pooled = torch.relu(linear(x[:, 0], sd["bert.t_pooler.dense.weight"], sd["bert.t_pooler.dense.bias"], mm_dtype))
    c = "clfs_dict.%s.logit_fc." % task
    h = gelu(linear(pooled, sd[c + "0.weight"], sd[c + "0.bias"], mm_dtype))
    h = layer_norm(h, sd[c + "2.weight"], sd[c + "2.bias"], cfg.layer_norm_eps)
    logits = linear(h, sd[c + "3.weight"], sd[c + "3.bias"], mm_dtype)
    if return_hidden:
        return logits, hidden
    return logits
