"""Deterministic synthetic weights and GQA-shaped batches.

Everything is produced by an integer hash (splitmix64 over int64 tensors) followed
by exact fp32 arithmetic, so the same seed gives bit-identical tensors in the
build container and on the GPU box, independent of torch's RNG implementation.
No dataset or checkpoint is available offline; these are the inputs both arms of
the benchmark and all parity tests run on.
"""
from __future__ import annotations

import math
import zlib

import torch

_M64 = (1 << 64) - 1


def _i64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_GOLD = _i64(0x9E3779B97F4A7C15)
_C1 = _i64(0xBF58476D1CE4E5B9)
_C2 = _i64(0x94D049BB133111EB)


def _lsr(z: torch.Tensor, s: int) -> torch.Tensor:
    return (z >> s) & ((1 << (64 - s)) - 1)


def hash_u64(seed: int, n: int, offset: int = 0) -> torch.Tensor:
    """splitmix64 of (seed, index) as int64 (bit pattern of the uint64 result)."""
    idx = torch.arange(offset, offset + n, dtype=torch.int64)
    z = idx * _GOLD + _i64(seed * 0xD1B54A32D192ED03 + 0x632BE59BD9B4E019)
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    z = z ^ _lsr(z, 31)
    return z


def hash_uniform(seed: int, n: int) -> torch.Tensor:
    """fp32 uniform in [0, 1) with 24 random bits per element (exact)."""
    chunk = 1 << 24
    if n <= chunk:
        return _lsr(hash_u64(seed, n), 40).to(torch.float32) * (1.0 / 16777216.0)
    out = torch.empty(n, dtype=torch.float32)
    for o in range(0, n, chunk):
        m = min(chunk, n - o)
        out[o:o + m] = _lsr(hash_u64(seed, m, offset=o), 40).to(torch.float32) * (1.0 / 16777216.0)
    return out


def hash_randint(seed: int, n: int, low: int, high: int) -> torch.Tensor:
    """int64 uniform in [low, high)."""
    return _lsr(hash_u64(seed, n), 11) % (high - low) + low


def _name_seed(seed: int, name: str) -> int:
    return (seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFF


def hash_tensor(seed: int, name: str, shape, scale: float, center: float = 0.0) -> torch.Tensor:
    """Uniform tensor with mean `center` and standard deviation `scale` (fp32)."""
    n = 1
    for s in shape:
        n *= int(s)
    u = hash_uniform(_name_seed(seed, name), n)
    half_width = scale * math.sqrt(3.0)
    return ((u * 2.0 - 1.0) * half_width + center).reshape(*shape)


def synth_state_dict(model_or_shapes, seed: int = 0) -> dict:
    """Weights for every key of a state_dict: matrices ~ U(std 0.02), LayerNorm gains
    ~ 1 + U(std 0.1), biases / LayerNorm offsets ~ U(std 0.1) so that no bias or LN path is
    vacuous.  Aliased keys (v_query ≡ query …) get identical tensors
    because the value only depends on the canonical parameter name."""
    if hasattr(model_or_shapes, "state_dict"):
        shapes = {k: tuple(v.shape) for k, v in model_or_shapes.state_dict().items()}
    else:
        shapes = dict(model_or_shapes)
    out = {}
    for key, shape in shapes.items():
        canon = canonical_key(key)
        if len(shape) >= 2:
            t = hash_tensor(seed, canon, shape, 0.02)
        elif _is_ln_weight(canon):
            t = hash_tensor(seed, canon, shape, 0.1, center=1.0)
        else:
            t = hash_tensor(seed, canon, shape, 0.1)
        out[key] = t
    return out


def _is_ln_weight(key: str) -> bool:
    low = key.lower()
    return key.endswith(".weight") and ("layernorm" in low or "layer_norm" in low or
                                        key.endswith("logit_fc.2.weight"))


def canonical_key(key: str) -> str:
    """Map the aliased state_dict names of the reference to the primary name
    (v_query ≡ query, v_dense ≡ dense, v_LayerNorm ≡ LayerNorm inside the
    encoder, image_token_type_embeddings ≡ new_token_type_embeddings, bert.pooler ≡
    bert.encoder.pooled_layer for M3P)."""
    if ".encoder.layer." in key:
        for a, b in ((".v_query.", ".query."), (".v_key.", ".key."), (".v_value.", ".value."),
                     (".v_dense.", ".dense."), (".v_LayerNorm.", ".LayerNorm.")):
            key = key.replace(a, b)
    key = key.replace("embeddings.image_token_type_embeddings.", "embeddings.new_token_type_embeddings.")
    key = key.replace("bert.pooler.dense.", "bert.encoder.pooled_layer.dense.")
    key = key.replace(".weight_orig", ".weight")
    return key


# ----------------------------------------------------------------------------------------------
# batches
# ----------------------------------------------------------------------------------------------
def synth_batch(batch_size: int, num_regions: int = 36, num_locs: int = 7, seq_len: int = 40,
                num_labels: int = 1842, vocab_size: int = 250002, feat_dim: int = 2048,
                seed: int = 1234, ragged_regions: bool = False, normalize: bool = False) -> dict:
    """One GQA-shaped batch: the 10-tuple fields the reference's
    ForwardModelsTrain unpacks (volta/volta/task_utils.py:L315), as a dict.

    question: <s>=0 … </s>=2, pad id 1, length U{6..40}; features relu(N(0,1))-like
    (here: relu of a centred uniform with unit variance); boxes [x1,y1,x2,y2,w,h,w*h] (UC2, 7)
    or [x1,y1,x2,y2,w*h] (M3P, 5, rows L2-normalised when `normalize`)."""
    B, T, R = batch_size, seq_len, num_regions
    lens = hash_randint(seed + 1, B, 6, T + 1)
    ar = torch.arange(T).unsqueeze(0)
    ids = hash_randint(seed + 2, B * T, 3, vocab_size).reshape(B, T)
    ids = torch.where(ar == 0, torch.zeros_like(ids), ids)
    ids = torch.where(ar == (lens.unsqueeze(1) - 1), torch.full_like(ids, 2), ids)
    ids = torch.where(ar >= lens.unsqueeze(1), torch.ones_like(ids), ids)
    input_mask = (ar < lens.unsqueeze(1)).long()
    segment_ids = torch.zeros(B, T, dtype=torch.int64)

    u = hash_uniform(seed + 3, B * R * feat_dim).reshape(B, R, feat_dim)
    feats = torch.relu((u * 2.0 - 1.0) * math.sqrt(3.0))
    if normalize:
        feats = feats / feats.norm(dim=-1, keepdim=True).clamp_min(1e-12)

    bx = hash_uniform(seed + 4, B * R * 4).reshape(B, R, 4)
    x1 = bx[..., 0] * 0.8
    y1 = bx[..., 1] * 0.8
    w = 0.05 + bx[..., 2] * 0.15
    h = 0.05 + bx[..., 3] * 0.15
    if num_locs == 7:
        spatials = torch.stack([x1, y1, x1 + w, y1 + h, w, h, w * h], dim=-1)
    else:
        spatials = torch.stack([x1, y1, x1 + w, y1 + h, w * h], dim=-1)
        if normalize:
            spatials = spatials / spatials.norm(dim=-1, keepdim=True).clamp_min(1e-12)

    if ragged_regions:
        n_img = hash_randint(seed + 5, B, 10, R + 1)
    else:
        n_img = torch.full((B,), R, dtype=torch.int64)
    image_mask = (torch.arange(R).unsqueeze(0) < n_img.unsqueeze(1)).long()
    if ragged_regions:  # the loader zero-pads missing regions (gqa_dataset_semantic_code_mix.py:L583-611)
        feats = feats * image_mask.unsqueeze(-1).to(feats.dtype)
        spatials = spatials * image_mask.unsqueeze(-1).to(spatials.dtype)

    labels = hash_randint(seed + 6, B, 0, num_labels)
    target = torch.zeros(B, num_labels, dtype=torch.float32)
    target[torch.arange(B), labels] = 1.0
    distances = label_distance_rows(labels, num_labels)
    question_id = torch.arange(B, dtype=torch.int64) + 1000 * seed
    return dict(features=feats.contiguous(), spatials=spatials.contiguous(), image_mask=image_mask,
                question=ids.contiguous(), target=target, input_mask=input_mask,
                segment_ids=segment_ids, question_id=question_id, labels=labels,
                distances=distances)


def label_distance_rows(labels: torch.Tensor, num_labels: int, seed: int = 77) -> torch.Tensor:
    """Rows of a fixed symmetric table D ~ U(0, 1.2), D[i,i] = 0 (mimics 1-cos of the
    reference's embedding_distance.pkl; gqa_dataset_semantic_code_mix.py:L371-381).
    D[i,j] = h(min(i,j), max(i,j)) is evaluated lazily for the requested rows only."""
    B = labels.numel()
    j = torch.arange(num_labels, dtype=torch.int64).unsqueeze(0).expand(B, -1)
    i = labels.reshape(B, 1).expand(-1, num_labels)
    lo = torch.minimum(i, j)
    hi = torch.maximum(i, j)
    key = (lo * num_labels + hi).reshape(-1)
    z = key * _GOLD + _i64(seed * 0xD1B54A32D192ED03 + 0x632BE59BD9B4E019)
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    z = z ^ _lsr(z, 31)
    u = _lsr(z, 40).to(torch.float32) * (1.0 / 16777216.0)
    d = (u * 1.2).reshape(B, num_labels)
    d[i == j] = 0.0
    return d


def batch_tuple(batch: dict):
    """The 10-tuple order of the reference loader (task_utils.py:L315)."""
    return (batch["features"], batch["spatials"], batch["image_mask"], batch["question"], batch["target"],
            batch["input_mask"], batch["segment_ids"], batch["question_id"],
            torch.zeros_like(batch["question_id"]), batch["distances"])
