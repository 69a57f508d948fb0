"""Oracle: M3P (`M3PForVLTasks` with m3p_base.json) forward.  TEST INFRASTRUCTURE ONLY.

Follows (reference, relative to /root/reference/volta/volta):
  encoders.py:L1033-1042           M3PModel.forward (lengths = mask sums; pooler on token 0)
  m3p_transformer.py:L877-964      M3PTransformerModel.jointfwd  ([img ; txt], prefix mask)
  m3p_transformer.py:L59-78        get_masks (prefix validity: arange(S) < n_img + n_txt)
  m3p_transformer.py:L149-210      MultiHeadAttention (q/sqrt(d) BEFORE QK^T, -inf mask, fp32 softmax)
  m3p_transformer.py:L222-227      TransformerFFN (lin1, erf-GeLU, lin2)
  m3p_transformer.py:L247-269      BertImageEmbeddings (feat proj + loc proj, one LN eps 1e-12)
  m3p_transformer.py:L554-560      BertPooler (tanh)
  encoders.py:L1311-1353           M3PForVLTasks.forward; SimpleClassifier L788-815

All encoder LayerNorms are nn.LayerNorm(eps=1e-12); the classifier LN uses config.layer_norm_eps
(default 1e-12).  `token_type_ids` is ignored.  Dropout off.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .uc2_oracle import gelu, linear, matmul


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def m3p_forward(sd, cfg, input_txt, input_imgs, image_loc, token_type_ids, attention_mask,
                image_attention_mask, task="TASK15", mm_dtype=None, return_hidden=False):
    P = "bert.encoder."
    dt = sd[P + "layer_norm_emb.weight"].dtype
    input_imgs = input_imgs.to(dt)
    image_loc = image_loc.to(dt)
    B, T = input_txt.shape
    R = input_imgs.shape[1]
    S = T + R
    H = cfg.n_heads if hasattr(cfg, "n_heads") else cfg.num_attention_heads
    D = cfg.hidden_size
    dh = D // H
    n_layers = int(getattr(cfg, "n_layers", 12))

    I = P + "image_embeddings."
    img = (linear(input_imgs, sd[I + "image_embeddings.weight"], sd[I + "image_embeddings.bias"], mm_dtype)
           + linear(image_loc, sd[I + "image_location_embeddings.weight"], sd[I + "image_location_embeddings.bias"], mm_dtype))
    img = _ln(img, sd[I + "LayerNorm.weight"], sd[I + "LayerNorm.bias"], 1e-12)
    txt = sd[P + "embeddings.weight"][input_txt]
    n = attention_mask.sum(dim=1) + image_attention_mask.sum(dim=1)
    valid = (torch.arange(S).unsqueeze(0) < n.unsqueeze(1))
    x = torch.cat([img, txt], dim=1) + sd[P + "position_embeddings.weight"][:S].unsqueeze(0)
    x = x * valid.unsqueeze(-1).to(dt)
    x = _ln(x, sd[P + "layer_norm_emb.weight"], sd[P + "layer_norm_emb.bias"], 1e-12)
    hidden = [x]

    def heads(y):
        return y.view(B, S, H, dh).transpose(1, 2)

    for i in range(n_layers):
        A = P + "attentions.%d." % i
        q = heads(linear(x, sd[A + "q_lin.weight"], sd[A + "q_lin.bias"], mm_dtype)) / math.sqrt(dh)
        k = heads(linear(x, sd[A + "k_lin.weight"], sd[A + "k_lin.bias"], mm_dtype))
        v = heads(linear(x, sd[A + "v_lin.weight"], sd[A + "v_lin.bias"], mm_dtype))
        scores = matmul(q, k.transpose(2, 3), mm_dtype)
        scores = scores.masked_fill(~valid[:, None, None, :], -float("inf"))
        w = torch.softmax(scores.float(), dim=-1).type_as(scores)
        ctx = matmul(w, v, mm_dtype).transpose(1, 2).contiguous().view(B, S, D)
        attn = linear(ctx, sd[A + "out_lin.weight"], sd[A + "out_lin.bias"], mm_dtype)
        x = _ln(x + attn, sd[P + "layer_norm1.%d.weight" % i], sd[P + "layer_norm1.%d.bias" % i], 1e-12)
        Fp = P + "ffns.%d." % i
        h = gelu(linear(x, sd[Fp + "lin1.weight"], sd[Fp + "lin1.bias"], mm_dtype))
        x = x + linear(h, sd[Fp + "lin2.weight"], sd[Fp + "lin2.bias"], mm_dtype)
        x = _ln(x, sd[P + "layer_norm2.%d.weight" % i], sd[P + "layer_norm2.%d.bias" % i], 1e-12)
        x = x * valid.unsqueeze(-1).to(dt)
        hidden.append(x)

    pooled = torch.tanh(linear(x[:, 0], sd[P + "pooled_layer.dense.weight"], sd[P + "pooled_layer.dense.bias"], mm_dtype))
    c = "clfs_dict.%s.logit_fc." % task
    eps = float(getattr(cfg, "layer_norm_eps", 1e-12))
    h = gelu(linear(pooled, sd[c + "0.weight"], sd[c + "0.bias"], mm_dtype))
    h = _ln(h, sd[c + "2.weight"], sd[c + "2.bias"], eps)
    logits = linear(h, sd[c + "3.weight"], sd[c + "3.bias"], mm_dtype)
    if return_hidden:
        return logits, hidden
    return logits
