"""M3P schedule of the step engine (M3PForVLTasks).

Differences from UC2 that matter (m3p_transformer.py:L877-964, L149-210, L247-269; encoders.py:L1033-1042):
sequence order [regions ; text], pooled token = first region, PREFIX validity mask (-inf on invalid
keys, hidden rows re-zeroed before the embedding LN and after every layer), one LayerNorm on the
region path, absolute positions 0..S-1, tanh pooler, LN eps 1e-12, classifier hidden 1536.
q / sqrt(d) before QK^T is the same arithmetic as scaling the scores (power of two).
Parameters outside `jointfwd` never receive gradients and are skipped by the fused optimiser.
"""
from __future__ import annotations

import torch

from . import m3p_layout, ops
from .ops import rowmap


def collect(eng):
    enc = eng.model.bert.encoder
    eng.enc = enc
    eng.pooler = enc.pooled_layer
    eng.pool_act = ops.ACT_TANH
    eng.layer_mods = None
    live = tuple("bert.encoder." + p for p in m3p_layout.live_prefixes(eng.cfg)) + ("clfs_dict.",)
    eng.live_name = lambda n: n.startswith(live) and not n.startswith("bert.encoder.embeddings_")  # noqa: E731


def mark_dead_slots(eng):
    """slots whose parameters are never touched by jointfwd: no gradient, skipped by AdamW
    (pytorch-transformers skips p.grad is None)"""
    for s in eng.slots:
        s.has_grad = all(eng.live_name(n) for n in s.names)


def _p(cfg):
    return float(getattr(cfg, "dropout", 0.1)), float(getattr(cfg, "attention_dropout", 0.1))


def forward(eng, ids, feats, locs, task_id, seg, tmask, imask, training, save):
    cfg, d, F, H, Ln = eng.cfg, eng.d, eng.F, eng.H, eng.L
    enc = eng.enc
    B, T = ids.shape
    R = feats.shape[1]
    S = T + R
    M, Mt, Mr = B * S, B * T, B * R
    ws = eng._workspace(B, T, R, save)
    lp = ws["lp"]
    bf = lp == torch.bfloat16
    eps = 1e-12
    eng.step_counter += 1
    eng._seed = eng.step_seed()
    eng._ticket += 1
    p_h, p_a = _p(cfg)
    IE = enc.image_embeddings
    if "e_joint" not in ws:
        ws["e_joint"] = torch.empty(M, d, dtype=torch.float32, device=eng.device)
        ws["m_e"] = torch.empty(M, dtype=torch.float32, device=eng.device)
        ws["r_e"] = torch.empty(M, dtype=torch.float32, device=eng.device)
        ws["arange_r"] = torch.arange(Mr, dtype=torch.int64, device=eng.device).view(B, R)
    ops.build_key_bias(tmask, imask, ws["kb"], ws["row_valid"], m3p_prefix=True)
    if save and eng.dp is not None and eng.dp.rows_mode():
        eng.dp.plan_rows(ids, int(getattr(cfg, "pad_index", 1)))
    feats2 = feats.view(Mr, -1)
    locs2 = locs.view(Mr, -1)
    if bf:
        if feats2.dtype == torch.bfloat16:
            fa = feats2                                  # delivered in the operand dtype
        else:
            ops.cast_f32_to_bf16(feats2, ws["feat_lp"])
            fa = ws["feat_lp"]
        ops.cast_f32_to_bf16(locs2, ws["loc_lp"])
        la = ws["loc_lp"]
    else:
        fa, la = feats2, locs2
    nl = locs2.shape[1]
    ops.gemm(la, eng._w(IE.image_location_embeddings, d, nl), ws["l_loc"], bias=eng._bias(IE.image_location_embeddings, d))
    ops.gemm(fa, eng._w(IE.image_embeddings, d, feats2.shape[1]), ws["g_img"], bias=eng._bias(IE.image_embeddings, d),
             residual=ws["l_loc"])
    ops.layernorm_fwd(ws["g_img"], IE.LayerNorm.weight.data, IE.LayerNorm.bias.data, eps, y=ws["t1"], mean=ws["m_img"],
                      rstd=ws["r_img"], dropout=eng._site(1, p_h, training))
    pos = enc.position_embeddings.weight.data
    ops.embed_text_fwd(ws["arange_r"], None, ws["t1"], pos, None, ws["e_joint"], pad_id=-1, out_map=rowmap(R, 0, S),
                       pos_mode=1, pos_base=0)
    ops.embed_text_fwd(ids, None, enc.embeddings.weight.data, pos, None, ws["e_joint"], pad_id=-1, out_map=rowmap(T, R, S),
                       pos_mode=1, pos_base=R)
    ln = enc.layer_norm_emb
    ops.layernorm_fwd(ws["e_joint"], ln.weight.data, ln.bias.data, eps, y=ws["x0"], y_lowp=ws["x0_lp"] if bf else None,
                      mean=ws["m_e"], rstd=ws["r_e"], row_valid=ws["row_valid"], valid_pre=True,
                      dropout=eng._site(0, p_h, training))
    x, x_lp, kb = ws["x0"], ws["x0_lp"], ws["kb"]
    eng._probe_capture("emb", x, B, S, None)
    # Packing (forward-only runs: the S <= 160 pipelined attention forward reads the packed layout, the wide backward does
    # not yet): the prefix mask keeps the first n = n_img + n_txt rows of [regions ; text] (get_masks, m3p_transformer.py
    # L59-78); rows past it are masked keys whose hidden state is re-zeroed after every layer - never observable, since
    # the pooler reads row 0.  Dropping them is exact, as for UC2 (engine._forward).
    pack = None
    row_valid = ws["row_valid"]
    if bf and eng.pack_enabled and not save and Ln > 0 and S <= 160:
        n_valid = tmask.sum(1) + imask.sum(1)
        prefix = (torch.arange(S, device=eng.device).unsqueeze(0) < n_valid.unsqueeze(1)).to(torch.int64)
        ops.pack_index(prefix[:, :R].contiguous(), prefix[:, R:].contiguous(), ws["cu"], ws["src"])
        Mp = int(ws["cu"][B])
        if Mp < M:
            ops.rows_gather(x, ws["src"], ws["xp"], n=Mp)
            ops.rows_gather(x_lp, ws["src"], ws["xp_lp"], n=Mp)
            ops.rows_gather(kb.view(M, 1), ws["src"], ws["kbp"], n=Mp)
            x, x_lp, kb = ws["xp"], ws["xp_lp"], ws["kbp"]
            pack = (Mp, ws["cu"], ws["src"])
            row_valid = None                      # every kept row is inside the prefix
    rows = pack[0] if pack else None
    eng.last_rows = rows if pack else M
    cu = pack[1] if pack else None
    nslots = len(ws["layers"])
    for l in range(Ln):
        a = enc.attentions[l]
        lw = ws["layers"][l % nslots]
        eng._probe_inject(l, x, x_lp, B, S, pack)
        eng._planned(ws, ("f", l, pack is not None, training, save), rows if rows is not None else M,
                     lambda a=a, lw=lw, x=x, x_lp=x_lp, l=l: eng._layer_fwd(
                         lw, x, x_lp, kb, B, S, a.q_lin, a.out_lin, enc.layer_norm1[l], enc.ffns[l].lin1, enc.ffns[l].lin2,
                         enc.layer_norm2[l], eps, 0.125, eng._site(2 + 4 * l, p_a, training), eng._site(3 + 4 * l, p_h, training),
                         eng._site(4 + 4 * l, p_h, training), bf, row_valid=row_valid, rows=rows, cu=cu, save=save))
        x, x_lp = lw["x2"], lw["x2_lp"]
        eng._probe_capture(l, x, B, S, pack)
    eng._probe_inject(Ln, x, x_lp, B, S, pack)
    logits = eng._head_fwd(ws, x_lp, B, S, task_id, eng.pooler.dense, eng.pool_act,
                           eng._site(2 + 4 * Ln, float(eng.model.dropout.p), training),
                           float(getattr(cfg, "layer_norm_eps", 1e-12)), bf, pack=pack)
    if save:
        eng._saved = dict(ws=ws, ids=ids, seg=seg, fa=fa, la=la, task=task_id, training=training, ticket=eng._ticket,
                          B=B, T=T, R=R, feats=feats, locs=locs)
    return logits


def backward(eng, dl, sv):
    cfg, d, Ln = eng.cfg, eng.d, eng.L
    enc = eng.enc
    ws = sv["ws"]
    B, T, R = sv["B"], sv["T"], sv["R"]
    S = T + R
    training = sv["training"]
    p_h, p_a = _p(cfg)
    last = ws["layers"][(Ln - 1) % len(ws["layers"])] if Ln > 0 else None
    x_last_lp = last["x2_lp"] if last is not None else ws["x0_lp"]
    eng._head_bwd(ws, sv, dl, x_last_lp, eng.pooler.dense, ops.ACT_DTANH,
                  eng._site(2 + 4 * Ln, float(eng.model.dropout.p), training))
    eng._probe_inject_grad(Ln, ws["dx"], B, S, None)
    for l in reversed(range(Ln)):
        a = enc.attentions[l]
        lw = ws["layers"][l]
        if l == 0:
            x_in, x_in_lp = ws["x0"], ws["x0_lp"]
        else:
            x_in, x_in_lp = ws["layers"][l - 1]["x2"], ws["layers"][l - 1]["x2_lp"]
        eng._planned(ws, ("b", l, False, training, eng.group_wgrad), B * S,
                     lambda a=a, lw=lw, x_in=x_in, x_in_lp=x_in_lp, l=l: eng._layer_bwd(
                         ws, lw, x_in, x_in_lp, ws["kb"], B, S, a.q_lin, a.out_lin, enc.layer_norm1[l], enc.ffns[l].lin1,
                         enc.ffns[l].lin2, enc.layer_norm2[l], 0.125, eng._site(2 + 4 * l, p_a, training),
                         eng._site(3 + 4 * l, p_h, training), eng._site(4 + 4 * l, p_h, training), row_valid=ws["row_valid"]))
        eng._probe_inject_grad(l, ws["dx"], B, S, None)
        if eng.dp is not None:
            eng.dp.layer_done(eng, l)
    dx = ws["dx"]
    ln = enc.layer_norm_emb
    if "de_joint" not in ws:
        ws["de_joint"] = torch.empty(B * S, d, dtype=torch.float32, device=eng.device)
    de = ws["de_joint"]
    ops.layernorm_bwd(dx, ws["e_joint"], ln.weight.data, ws["m_e"], ws["r_e"], dx=de, dgamma=eng.grad_view(ln.weight),
                      dbeta=eng.grad_view(ln.bias), row_valid=ws["row_valid"], valid_pre=True,
                      dropout_out=eng._site(0, p_h, training))
    gpos = eng.grad_view(enc.position_embeddings.weight)
    # text rows: word table (padding row skipped: nn.Embedding padding_idx = pad_index) + positions R..S-1
    rows_mode = eng.dp is not None and eng.dp.rows_mode()      # data parallel: word rows are reduced after the all-gather
    ops.embed_text_bwd(sv["ids"], None, de, None if rows_mode else eng.grad_view(enc.embeddings.weight), gpos, None,
                       pad_id=int(getattr(cfg, "pad_index", 1)), dout_map=rowmap(T, R, S), pos_mode=1, pos_base=R)
    if rows_mode:
        eng.dp.add_rows(sv["ids"], de.view(B, S, d)[:, R:, :].reshape(B * T, d))
        eng.dp.rows_ready(eng)
    else:
        eng.mark_word_rows(sv["ids"])
    # region rows: positions 0..R-1
    ops.embed_text_bwd(ws["arange_r"], None, de, None, gpos, None, pad_id=-1, dout_map=rowmap(R, 0, S), pos_mode=1,
                       pos_base=0)
    IE = enc.image_embeddings
    ops.layernorm_bwd(de, ws["g_img"], IE.LayerNorm.weight.data, ws["m_img"], ws["r_img"], dx=ws["ds"], dx_lowp=ws["dg"],
                      dgamma=eng.grad_view(IE.LayerNorm.weight), dbeta=eng.grad_view(IE.LayerNorm.bias),
                      dbias=eng._gvec(eng.b_slot[id(IE.image_embeddings)], d), dy_map=rowmap(R, 0, S),
                      dropout_out=eng._site(1, p_h, training))
    eng._wgrad(ws["dg"], sv["fa"], IE.image_embeddings, d, sv["fa"].shape[1])
    # the location projection was added to the feature projection before the LN: same gradient
    ops.colsum(ws["ds"], eng._gvec(eng.b_slot[id(IE.image_location_embeddings)], d))
    eng._wgrad(ws["dg"], sv["la"], IE.image_location_embeddings, d, sv["la"].shape[1])
