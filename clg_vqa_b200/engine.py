"""Step engine: owns the flat parameter / gradient arenas, the bf16 weight shadows, the packed
SFT masks and the activation workspace, and runs the hand-scheduled forward and backward of the
single-stream encoder as a sequence of C-ABI kernel calls.

Data layout in HBM
  p_flat / g_flat   fp32 arenas; every parameter (or fused group: q|k|v weights, q|k|v biases of a
                    layer) starts on a 2048-element boundary; `param.data` / `param.grad` are views.
  shadow            bf16 copies of the GEMM weights (W (.) mask), refreshed by the fused optimiser or
                    when a parameter's version counter moved (foreign optimiser, load_state_dict).
  mask_bits         1 bit per GEMM-weight element (torch.nn.utils.prune `weight_mask`, packed).
  workspace         activations for one forward: fp32 residual stream [B*S, 768] + bf16 operand
                    copies, qkv / ctx / FFN hidden in the operand dtype, LN statistics, log-sum-exp.
The residual stream and every LayerNorm / softmax / loss stay fp32; only GEMM and attention operands
are bf16 - the rounding points of the bf16 oracle.
"""
from __future__ import annotations

import os

import math
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import config as cfgmod
from . import ops
from .ops import DropSite, rowmap

CHUNK = L.CHUNK


def _pad(n: int, m: int = CHUNK) -> int:
    return (n + m - 1) // m * m


class _Slot:
    """one optimiser tensor of the arena: a parameter or a fused group of parameters"""
    __slots__ = ("names", "params", "start", "numel", "shadow_off", "mask_off", "mask_mods", "has_grad", "index")

    def __init__(self, names, params):
        self.names, self.params = names, params
        self.start = 0
        self.numel = sum(p.numel() for p in params)
        self.shadow_off = -1
        self.mask_off = -1
        self.mask_mods = None
        self.has_grad = True
        self.index = -1


class _EngineFn(torch.autograd.Function):
    """Bridges the engine into autograd: forward hands out the logits, backward receives
    d(loss)/d(logits) and runs the engine's backward, which writes parameter gradients straight into
    the gradient arena (`param.grad` views)."""

    @staticmethod
    def forward(ctx, anchor, engine, logits, ticket):
        ctx.engine = engine
        ctx.ticket = ticket
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        ctx.engine.backward(dlogits, ctx.ticket)
        return None, None, None, None


class Engine:
    def __init__(self, model, family: str = "uc2"):
        self.model = model
        self.family = family
        self.cfg = model.config
        self.precision = model.__dict__.get("_precision", "bf16")
        self.device: Optional[torch.device] = None
        self.slots: List[_Slot] = []
        self.p_flat = self.g_flat = None
        self.shadow = None
        self.mask_bits = None
        self._ws: Dict = {}
        self._anchor = None
        self._ticket = 0
        self._saved = None
        self._shadow_versions = None
        self._mask_versions = None
        self._struct_sig = None
        # dropout stream: keyed on torch's seed (torch.manual_seed / args.seed picks the stream, train_task.py:L189-191),
        # the rank (parallel.attach folds it in: independent masks per process, as with per-process RNGs) and the step
        # counter (saved / restored by checkpoint.save / resume)
        self.seed_root = (int(torch.initial_seed()) ^ 0x5EED) & 0x7FFFFFFFFFFFFFFF    # rank-free part: what a checkpoint stores
        self.rank_salt = 0                                                              # set by parallel.attach
        self.step_counter = 0
        self.grads_consumed = True       # True: next backward zeroes the gradient arena first
        self.pack_enabled = os.environ.get("CLG_PACK", "1") != "0"   # drop masked positions before the encoder layers
        self.group_wgrad = os.environ.get("CLG_GROUP_WGRAD", "1") != "0"   # QKV + attention-output weight gradients in one launch
        self.use_plans = os.environ.get("CLG_PLANS", "1") != "0"           # replay the per-layer launch lists natively (ops.run_plan)
        self.dp = None                   # optional parallel.GradExchange
        self.grad_scale = 1.0            # 1/world under data parallelism (parallel.attach): the arena holds the SUM over ranks
        self.probe = None                # parity-test hook: {"inject": {l: [B,S,d] fp32}, "capture": {}} (see _probe_*)
        self.grads_zeroed = False        # the fused optimiser step left the gradient arena all zero
        self.L = cfgmod.num_encoder_layers(self.cfg)
        self.d = self.cfg.hidden_size
        self.H = self.d // 64
        self.F = self.cfg.intermediate_size

    # ------------------------------------------------------------------------------------------
    # structure discovery
    # ------------------------------------------------------------------------------------------
    def _wb(self, mod, which="weight"):
        """(parameter, mask-or-None) honouring torch.nn.utils.prune's re-parametrisation
        (train_task_sft.py:L122-132: `weight_orig` Parameter + `weight_mask` buffer)."""
        if which == "weight" and "weight_orig" in mod._parameters:
            return mod._parameters["weight_orig"], mod._buffers.get("weight_mask")
        return mod._parameters[which], None

    def _structure_signature(self):
        """(name, storage pointer) of every parameter: changes when parameters are re-created / moved / renamed (model.to(),
        prune re-parametrisation).  Evaluated on every forward, so it walks a cached module list and the modules'
        own `_parameters` dicts instead of the recursive named_parameters() generators (3 x cheaper)."""
        mods = self.__dict__.get("_sig_modules")
        if mods is None:
            mods = [(n, m) for n, m in self.model.named_modules()]
            self.__dict__["_sig_modules"] = mods
        sig = []
        seen = set()
        for n, m in mods:
            for k, p in m._parameters.items():
                if p is None or id(p) in seen:
                    continue
                seen.add(id(p))
                sig.append((n, k, p.data_ptr() if p.is_cuda else -1))
        return tuple(sig)

    def _collect_uc2(self):
        m = self.model
        E = m.bert.embeddings
        g = {}   # logical name -> module
        self.mods = g
        g["emb"] = E
        layers = []
        for l in range(self.L):
            a = m.bert.encoder.layer[2 * l]
            f = m.bert.encoder.layer[2 * l + 1]
            layers.append((a.attention_self, a.attention_output, f.intermediate, f.output))
        self.layer_mods = layers
        self.pooler = m.bert.t_pooler
        self.pool_act = ops.ACT_RELU if getattr(self.cfg, "fusion_act", "relu") == "relu" else ops.ACT_TANH

    # ------------------------------------------------------------------------------------------
    # arenas
    # ------------------------------------------------------------------------------------------
    def _gemm_weight_modules(self):
        """[(list of modules fused along N, needs_shadow)] in arena order"""
        out = []
        if self.family == "uc2":
            E = self.model.bert.embeddings
            out.append([E.image_embeddings])
            out.append([E.image_location_embeddings])
            for (sa, so, fi, fo) in self.layer_mods:
                out.append([sa.query, sa.key, sa.value])
                out.append([so.dense])
                out.append([fi.dense])
                out.append([fo.dense])
            out.append([self.pooler.dense])
        else:
            enc = self.model.bert.encoder
            out.append([enc.image_embeddings.image_embeddings])
            out.append([enc.image_embeddings.image_location_embeddings])
            for i in range(self.L):
                a = enc.attentions[i]
                out.append([a.q_lin, a.k_lin, a.v_lin])
                out.append([a.out_lin])
                out.append([enc.ffns[i].lin1])
                out.append([enc.ffns[i].lin2])
            out.append([enc.pooled_layer.dense])
        for task, clf in self.model.clfs_dict.items():
            out.append([clf.logit_fc[0]])
            out.append([clf.logit_fc[3]])
        return out

    def adopt(self, device):
        """(Re)build the arenas on `device` and rebind every parameter to its arena view."""
        ops.device_check(device.index if device.index is not None else torch.cuda.current_device())
        self.device = device
        if self.family == "uc2":
            self._collect_uc2()
        else:
            self._collect_m3p()
        name_of = {}
        for n, p in self.model.named_parameters():
            name_of.setdefault(id(p), n)
        slots: List[_Slot] = []
        taken = set()
        self.w_slot = {}   # id(first module of a GEMM group) -> slot (weights)
        self.b_slot = {}
        for group in self._gemm_weight_modules():
            ws = [self._wb(mm)[0] for mm in group]
            s = _Slot([name_of[id(p)] for p in ws], ws)
            s.mask_mods = group
            slots.append(s)
            self.w_slot[id(group[0])] = s
            taken.update(id(p) for p in ws)
            bs = [mm._parameters["bias"] for mm in group]
            sb = _Slot([name_of[id(p)] for p in bs], bs)
            slots.append(sb)
            self.b_slot[id(group[0])] = sb
            taken.update(id(p) for p in bs)
        for n, p in self.model.named_parameters():
            if id(p) in taken:
                continue
            taken.add(id(p))
            slots.append(_Slot([n], [p]))
        # members of a fused group must be laid out back to back: every member size is a multiple of 8
        off = 0
        for i, s in enumerate(slots):
            s.index = i
            s.start = off
            off += _pad(s.numel)
        total = off
        p_flat = torch.zeros(total, dtype=torch.float32, device=device)
        g_flat = torch.zeros(total, dtype=torch.float32, device=device)
        shadow_off = 0
        mask_off = 0
        for s in slots:
            o = s.start
            for p in s.params:
                n = p.numel()
                view = p_flat[o:o + n].view(p.shape)
                view.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = view
                p.grad = None
                o += n
            if s.mask_mods is not None:
                s.shadow_off = shadow_off
                s.mask_off = mask_off
                shadow_off += _pad(s.numel)
                mask_off += _pad(s.numel)
        self.slots = slots
        if self.family == "m3p":
            from . import engine_m3p
            engine_m3p.mark_dead_slots(self)
        self.p_flat, self.g_flat = p_flat, g_flat
        self.total = total
        self.shadow = torch.zeros(max(shadow_off, CHUNK), dtype=torch.bfloat16, device=device)
        self.mask_bits = torch.full((max(mask_off, CHUNK) // 32,), -1, dtype=torch.int32, device=device)
        self.has_masks = False
        self.w32 = None                      # fp32 masked copy (fp32 mode with SFT masks only)
        self._slot_of_param = {}
        for s in slots:
            for p in s.params:
                self._slot_of_param[id(p)] = s
        self._build_tables()
        self._shadow_versions = None
        self._mask_versions = None
        self._struct_sig = self._structure_signature()
        self._anchor = torch.zeros(1, device=device, requires_grad=True)
        self._ws.clear()
        self.grads_consumed = True
        self.grads_zeroed = False
        self._partials = torch.zeros(4096, dtype=torch.float32, device=device)
        from . import optim as _optim
        _optim.register_engine(self)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=device)
        # one byte per 256 arena elements: 1 = the block has received a non-zero gradient at some point (optim.py)
        self._hot = torch.zeros(self.p_flat.numel() // 256, dtype=torch.uint8, device=device)
        self._hot_valid = False
        # per-backward marks of the word-table blocks that received gradient rows (ops.mark_rows): the norm pass reads only
        # those blocks of the 768 MB table gradient (optim.clip_grad_norm_); cleared whenever the arena is cleared
        wslot = self._slot_of_param[id(self.word_table_param())]
        self._table_blocks = (wslot.start // 256, (wslot.start + wslot.numel + 255) // 256)
        self._touched = torch.zeros(self.p_flat.numel() // 256, dtype=torch.uint8, device=device)
        self._touch_valid = False

    def _build_tables(self, lr_wd=None):
        """chunk -> tensor table and the clgTensorMeta array (lr / wd filled by the optimiser)."""
        n_chunks = self.total // CHUNK
        chunk_tensor = torch.full((n_chunks,), -1, dtype=torch.int32)
        metas = (L.TensorMeta * len(self.slots))()
        for i, s in enumerate(self.slots):
            c0 = s.start // CHUNK
            c1 = (s.start + _pad(s.numel)) // CHUNK
            chunk_tensor[c0:c1] = i
            lr, wd = (0.0, 0.0) if lr_wd is None else lr_wd[i]
            metas[i] = L.TensorMeta(lr, wd, s.shadow_off, s.mask_off if (s.mask_mods is not None and self.has_masks) else -1,
                                    s.start, s.numel)
        self.chunk_tensor = chunk_tensor.to(self.device)
        self.metas_host = metas
        self.tensors_dev = torch.frombuffer(bytearray(bytes(metas)), dtype=torch.uint8).to(self.device)

    def set_precision(self, mode):
        if mode != self.precision:
            self.precision = mode
            self._ws.clear()
            self._shadow_versions = None

    # views -------------------------------------------------------------------------------------
    def grad_view(self, p):
        s = self._slot_of_param[id(p)]
        o = s.start
        for q in s.params:
            if q is p:
                return self.g_flat[o:o + p.numel()].view(p.shape)
            o += q.numel()
        raise KeyError

    def _g(self, slot: _Slot, rows: int, cols: int):
        return self.g_flat[slot.start:slot.start + rows * cols].view(rows, cols)

    def _gvec(self, slot: _Slot, n: Optional[int] = None):
        n = slot.numel if n is None else n
        return self.g_flat[slot.start:slot.start + n]

    def _w(self, mod0, rows, cols):
        """GEMM weight operand [rows, cols] of the group starting at module mod0 (operand dtype)."""
        s = self.w_slot[id(mod0)]
        if self.precision == "bf16":
            return self.shadow[s.shadow_off:s.shadow_off + rows * cols].view(rows, cols)
        if self.has_masks:
            return self.w32[s.shadow_off:s.shadow_off + rows * cols].view(rows, cols)
        return self.p_flat[s.start:s.start + rows * cols].view(rows, cols)

    def _bias(self, mod0, n):
        s = self.b_slot[id(mod0)]
        return self.p_flat[s.start:s.start + n]

    def _wmask(self, mod0):
        if not self.has_masks:
            return None
        s = self.w_slot[id(mod0)]
        if not any("weight_mask" in mm._buffers for mm in s.mask_mods):
            return None
        return self.mask_bits[s.mask_off // 32:]

    # ------------------------------------------------------------------------------------------
    # per-forward refresh of masks and shadows
    # ------------------------------------------------------------------------------------------
    def _refresh(self):
        sig = self._structure_signature()
        if sig != self._struct_sig:
            # parameters were re-created / moved (model.to(), prune re-parametrisation renames keep the
            # same tensors but change names) -> rebuild the arenas from the current values
            same_ptrs = [x[-1] for x in sig] == [x[-1] for x in self._struct_sig]
            if same_ptrs:
                self._struct_sig = sig
                self._shadow_versions = None
                self._mask_versions = None
            else:
                self.adopt(self.device)
        # masks
        mv = []
        any_mask = False
        for s in self.slots:
            if s.mask_mods is None:
                continue
            for mm in s.mask_mods:
                mk = mm._buffers.get("weight_mask")
                if mk is not None:
                    any_mask = True
                    mv.append((id(mm), mk.data_ptr(), mk._version))
                else:
                    mv.append((id(mm), 0, 0))
        mv = tuple(mv)
        if mv != self._mask_versions:
            self._mask_versions = mv
            had = self.has_masks
            self.has_masks = any_mask
            self.mask_bits.fill_(-1)
            if any_mask:
                for s in self.slots:
                    if s.mask_mods is None:
                        continue
                    o = s.mask_off
                    for mm in s.mask_mods:
                        mk = mm._buffers.get("weight_mask")
                        n = self._wb(mm)[0].numel()
                        if mk is not None:
                            mk = mk.to(device=self.device, dtype=torch.float32).contiguous()
                            ops.pack_mask(mk, self.mask_bits[o // 32:], n)
                        o += n
            if had != any_mask:
                self._rebuild_meta_masks()
            self._shadow_versions = None
            self._drop_plans()              # recorded launches hold the mask pointers of the old state
        # shadows
        need_shadow = self.precision == "bf16" or self.has_masks
        if need_shadow:
            sv = tuple(p._version for s in self.slots if s.mask_mods is not None for p in s.params)
            if sv != self._shadow_versions:
                self._shadow_versions = sv
                if self.precision == "bf16":
                    ops.cast_shadow(self.p_flat, self.chunk_tensor, self.tensors_dev, self.shadow,
                                    self.mask_bits if self.has_masks else None)
                else:
                    if self.w32 is None or self.w32.numel() != self.shadow.numel():
                        self.w32 = torch.zeros(self.shadow.numel(), dtype=torch.float32, device=self.device)
                    for s in self.slots:
                        if s.mask_mods is None:
                            continue
                        dst = self.w32[s.shadow_off:s.shadow_off + s.numel]
                        dst.copy_(self.p_flat[s.start:s.start + s.numel])
                        ops.apply_mask(dst, self.mask_bits[s.mask_off // 32:], s.numel)

    def _rebuild_meta_masks(self):
        for i, s in enumerate(self.slots):
            self.metas_host[i].mask_off = s.mask_off if (s.mask_mods is not None and self.has_masks) else -1
        self.tensors_dev = torch.frombuffer(bytearray(bytes(self.metas_host)), dtype=torch.uint8).to(self.device)

    def invalidate_masks(self):
        """masks were rewritten behind torch's back (pruning.global_magnitude_prune): repack the mask bits and re-cast
        the masked weight shadow at the next forward"""
        self._mask_versions = None
        self._shadow_versions = None

    def mark_shadow_fresh(self):
        """called by the fused optimiser after it rewrote parameters AND shadows in one kernel"""
        self._shadow_versions = tuple(p._version for s in self.slots if s.mask_mods is not None for p in s.params)

    # ------------------------------------------------------------------------------------------
    # workspace
    # ------------------------------------------------------------------------------------------
    def _workspace(self, B, T, R, training):
        key = (B, T, R, self.precision, bool(training))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        self._ws.clear()                    # one shape at a time keeps the footprint bounded
        dev, d, F, H = self.device, self.d, self.F, self.H
        lp = torch.bfloat16 if self.precision == "bf16" else torch.float32
        S = T + R
        M, Mt, Mr = B * S, B * T, B * R
        f32 = torch.float32

        def E(*shape, dtype=f32):
            return torch.empty(*shape, dtype=dtype, device=dev)

        ws = {"B": B, "T": T, "R": R, "S": S, "lp": lp}
        ws["kb"] = E(B, S)
        ws["row_valid"] = E(B * S, dtype=torch.uint8)
        ws["e_txt"] = E(Mt, d); ws["m_txt"] = E(Mt); ws["r_txt"] = E(Mt)
        fdim = self.cfg.v_feature_size
        nl = self.cfg.num_locs
        ws["feat_lp"] = E(Mr, fdim, dtype=lp) if lp != f32 else None
        ws["loc_lp"] = E(Mr, nl, dtype=lp) if lp != f32 else None
        ws["g_img"] = E(Mr, d); ws["l_loc"] = E(Mr, d); ws["t1"] = E(Mr, d); ws["t2"] = E(Mr, d)
        for k in ("m_loc", "r_loc", "m_img", "r_img", "m_v", "r_v"):
            ws[k] = E(Mr)
        ws["x0"] = E(M, d)
        ws["x0_lp"] = E(M, d, dtype=lp) if lp != f32 else ws["x0"]
        if lp != f32:
            # packed joint sequence (masked positions dropped after the embeddings, see _forward)
            ws["cu"] = torch.zeros(B + 1, dtype=torch.int32, device=dev)
            ws["src"] = torch.zeros(M, dtype=torch.int32, device=dev)
            ws["xp"] = E(M, d); ws["xp_lp"] = E(M, d, dtype=lp); ws["kbp"] = E(M, 1)
            ws["xfirst"] = E(B, d, dtype=lp)
        nslots = self.L if training else 1
        layers = []
        for _ in range(nslots):
            lw = {"qkv": E(M, 3 * d, dtype=lp), "lse": E(B, H, S), "ctx": E(M, d, dtype=lp),
                  "y1": E(M, d), "m1": E(M), "r1": E(M), "x1": E(M, d),
                  "u": E(M, F, dtype=lp), "h": E(M, F, dtype=lp),
                  "y2": E(M, d), "m2": E(M), "r2": E(M), "x2": E(M, d)}
            lw["x1_lp"] = E(M, d, dtype=lp) if lp != f32 else lw["x1"]
            lw["x2_lp"] = E(M, d, dtype=lp) if lp != f32 else lw["x2"]
            layers.append(lw)
        ws["layers"] = layers
        P = self.cfg.pooler_size
        ws["pact"] = E(B, P)                       # pooler activation output, before dropout
        ws["pooled32"] = E(B, P)
        ws["pooled"] = E(B, P, dtype=lp) if lp != f32 else ws["pooled32"]
        heads = {}
        for task, clf in self.model.clfs_dict.items():
            Cc = self._wb(clf.logit_fc[0])[0].shape[0]
            A = clf.logit_fc[3].bias.shape[0]
            heads[task] = {"C": Cc, "A": A, "uc": E(B, Cc), "hc": E(B, Cc), "mc": E(B), "rc": E(B),
                           # logits rows padded to a 16 B pitch (1842 -> 1844 fp32): the GEMM epilogue then stores them with
                           # vector instructions instead of one column per lane; callers see the [B, A] view
                           "hln": E(B, Cc, dtype=lp), "logits": E(B, (A + 3) // 4 * 4)[:, :A]}
        ws["heads"] = heads
        if training:
            ws["dx"] = E(M, d); ws["dy"] = E(M, d)
            ws["dgo"] = E(M, d, dtype=lp)
            ws["du"] = E(M, F, dtype=lp)
            ws["dctx"] = E(M, d, dtype=lp)
            ws["dqkv"] = E(M, 3 * d, dtype=lp)
            ws["de_txt"] = E(Mt, d)
            ws["ds"] = E(Mr, d); ws["dg"] = E(Mr, d, dtype=lp); ws["dl"] = E(Mr, d, dtype=lp)
            ws["dpre"] = E(B, P)
            ws["dpre_lp"] = E(B, P, dtype=lp) if lp != f32 else ws["dpre"]
            if lp != f32:
                ws["dxfirst"] = E(B, d); ws["dxd"] = E(M, d)
            for task, hw in heads.items():
                hw["dhln"] = E(B, hw["C"]); hw["dhc"] = E(B, hw["C"])
                if lp != f32:
                    # bf16 GEMM operands of the head backward (tensor-core path); the logit gradient gets a
                    # 16 B-aligned row pitch (1842 -> 1848) so TMA can fetch it
                    a_pad = (hw["A"] + 7) // 8 * 8
                    hw["dl_lp"] = E(B, a_pad, dtype=lp)[:, :hw["A"]]
                    hw["dhc_lp"] = E(B, hw["C"], dtype=lp)
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------
    # dropout sites
    # ------------------------------------------------------------------------------------------
    def step_seed(self) -> int:
        return (self.base_seed * 0x9E3779B1 + self.step_counter) & 0x7FFFFFFFFFFFFFFF

    @property
    def base_seed(self) -> int:
        return (self.seed_root ^ self.rank_salt) & 0x7FFFFFFFFFFFFFFF

    def set_rank(self, rank: int):
        """independent dropout masks per data-parallel rank (the reference's processes have independent RNG streams)"""
        self.rank_salt = ((int(rank) + 1) * 0x9E3779B97F4A7C15) & 0x7FFFFFFFFFFFFFFF

    def rng_state(self) -> dict:
        """the rank-free seed and the step counter: rank 0 writes the checkpoint, every rank restores it (before or after
        parallel.attach) and still draws its own masks"""
        return {"base_seed": int(self.seed_root), "step_counter": int(self.step_counter)}

    def set_rng_state(self, st: dict):
        self.seed_root = int(st["base_seed"]) & 0x7FFFFFFFFFFFFFFF
        self.step_counter = int(st["step_counter"])

    def _site(self, idx: int, p: float, training: bool) -> DropSite:
        if not training or p <= 0.0:
            return ops.NO_DROP
        return DropSite(p, self._seed, idx << 40)

    # ------------------------------------------------------------------------------------------
    # parity-test hook (tests/test_layerwise_gpu.py): teacher-forced per-layer comparison.  `probe["inject"][l]` replaces
    # the input of encoder layer l (l = L: the input of the head) by a caller-supplied [B, S, d] fp32 hidden state (and
    # its bf16 rounding as the GEMM operand); `probe["capture"]` receives the embedding output ("emb") and every layer's
    # output as [B, S, d] fp32 (packed rows scattered back; dropped rows read 0).  Off (None) on the product path.
    # ------------------------------------------------------------------------------------------
    def _probe_inject(self, l, x, x_lp, B, S, pack):
        inj = self.probe.get("inject", {}).get(l) if self.probe is not None else None
        if inj is None:
            return
        v = inj.to(device=self.device, dtype=torch.float32).reshape(B * S, self.d)
        n = B * S
        if pack is not None:
            n = pack[0]
            v = v[pack[2][:n].long()]
        x[:n].copy_(v)
        if x_lp is not x:
            x_lp[:n].copy_(v.to(x_lp.dtype))

    def _probe_inject_grad(self, l, dx, B, S, pack):
        """backward twin of _probe_inject: `probe["inject_grad"][l]` replaces d(loss)/d(hidden state l) ([B, S, d] fp32),
        so that every layer's parameter gradients can be compared with the oracle's without the layers above it"""
        inj = self.probe.get("inject_grad", {}).get(l) if self.probe is not None else None
        if inj is None:
            return
        v = inj.to(device=self.device, dtype=torch.float32).reshape(B * S, self.d)
        n = B * S
        if pack is not None:
            n = pack[0]
            v = v[pack[2][:n].long()]
        dx[:n].copy_(v)

    def _probe_capture(self, key, x, B, S, pack):
        if self.probe is None:
            return
        if pack is None:
            out = x[:B * S].reshape(B, S, self.d).clone()
        else:
            out = torch.zeros(B * S, self.d, dtype=x.dtype, device=x.device)
            out[pack[2][:pack[0]].long()] = x[:pack[0]]
            out = out.view(B, S, self.d)
        self.probe.setdefault("capture", {})[key] = out

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def run(self, input_txt, input_imgs, image_loc, task_id, token_type_ids, attention_mask, image_attention_mask):
        if not input_imgs.is_cuda:
            raise ops.ClgError("clg_vqa_b200 runs on a B200 GPU only: inputs are on %s and there is no CPU fallback"
                               % input_imgs.device)
        dev = input_imgs.device
        if self.device is None or self.device != dev or self.p_flat is None:
            self.adopt(dev)
        want_grad = torch.is_grad_enabled()
        with torch.no_grad(), ops.nvtx_range("clg.forward"):
            self._refresh()
            training = bool(self.model.training)
            feats = input_imgs.contiguous()
            if not (feats.dtype == torch.bfloat16 and self.precision == "bf16"):
                feats = feats.float()     # bf16 features are the GEMM operand as delivered: same rounding point
            logits = self._forward(input_txt.contiguous(), feats,
                                   image_loc.contiguous().float(), task_id, token_type_ids.contiguous(),
                                   attention_mask.contiguous(), image_attention_mask.contiguous(),
                                   training, save=want_grad)
        if want_grad:
            return _EngineFn.apply(self._anchor, self, logits, self._ticket)
        return logits.clone()

    def _forward(self, ids, feats, locs, task_id, seg, tmask, imask, training, save):
        if self.family == "m3p":
            return self._forward_m3p(ids, feats, locs, task_id, seg, tmask, imask, training, save)
        cfg, d, F, H, Ln = self.cfg, self.d, self.F, self.H, self.L
        B, T = ids.shape
        R = feats.shape[1]
        S = T + R
        M, Mt, Mr = B * S, B * T, B * R
        ws = self._workspace(B, T, R, save)
        lp = ws["lp"]
        bf = lp == torch.bfloat16
        eps = float(cfg.layer_norm_eps)
        self.step_counter += 1
        self._seed = self.step_seed()
        self._ticket += 1
        E = self.model.bert.embeddings
        p_h = float(cfg.hidden_dropout_prob)
        p_a = float(cfg.attention_probs_dropout_prob)

        ops.build_key_bias(tmask, imask, ws["kb"])
        if save and self.dp is not None and self.dp.rows_mode():
            self.dp.plan_rows(ids, int(cfg.pad_token_id))       # id-only half of the word-row exchange, off the backward's tail
        # ---- text embeddings (embeddings.py:L640-654)
        ops.embed_text_fwd(ids, seg, E.word_embeddings.weight.data, E.position_embeddings.weight.data,
                           E.new_token_type_embeddings.weight.data, ws["e_txt"], pad_id=cfg.pad_token_id)
        ops.layernorm_fwd(ws["e_txt"], E.LayerNorm.weight.data, E.LayerNorm.bias.data, eps, y=ws["x0"],
                          y_lowp=ws["x0_lp"] if bf else None, mean=ws["m_txt"], rstd=ws["r_txt"],
                          out_map=rowmap(T, 0, S), dropout=self._site(0, p_h, training))
        # ---- region embeddings (embeddings.py:L661-667)
        feats2 = feats.view(Mr, -1)
        locs2 = locs.view(Mr, -1)
        if bf:
            if feats2.dtype == torch.bfloat16:
                fa = feats2                                  # delivered in the operand dtype: no cast pass, half the H2D bytes
            else:
                ops.cast_f32_to_bf16(feats2, ws["feat_lp"])
                fa = ws["feat_lp"]
            ops.cast_f32_to_bf16(locs2, ws["loc_lp"])
            la = ws["loc_lp"]
        else:
            fa, la = feats2, locs2
        nl = locs2.shape[1]
        ops.gemm(fa, self._w(E.image_embeddings, d, feats2.shape[1]), ws["g_img"], bias=self._bias(E.image_embeddings, d))
        ops.gemm(la, self._w(E.image_location_embeddings, d, nl), ws["l_loc"], bias=self._bias(E.image_location_embeddings, d))
        ops.layernorm_fwd(ws["l_loc"], E.image_location_layer_norm.weight.data, E.image_location_layer_norm.bias.data, eps,
                          y=ws["t1"], mean=ws["m_loc"], rstd=ws["r_loc"])
        ops.layernorm_fwd(ws["g_img"], E.image_layer_norm.weight.data, E.image_layer_norm.bias.data, eps, y=ws["t2"],
                          mean=ws["m_img"], rstd=ws["r_img"], add_mat=ws["t1"],
                          add_vec=E.new_token_type_embeddings.weight.data[1])
        ops.layernorm_fwd(ws["t2"], E.v_LayerNorm.weight.data, E.v_LayerNorm.bias.data, eps, y=ws["x0"],
                          y_lowp=ws["x0_lp"] if bf else None, mean=ws["m_v"], rstd=ws["r_v"],
                          out_map=rowmap(R, T, S), dropout=self._site(1, p_h, training))
        # ---- encoder layers
        x, x_lp, kb = ws["x0"], ws["x0_lp"], ws["kb"]
        self._probe_capture("emb", x, B, S, None)
        # Packing: masked positions (text padding, missing regions) are dropped before the encoder.  Exact: their key
        # bias (-10000) makes exp() underflow to 0 for every kept query, nothing downstream reads their outputs (the
        # pooler takes token 0, which is always kept) and hence their gradients vanish identically.  The layers then run
        # on Mp = sum(kept) rows instead of B*S; the pipelined attention kernels read the per-sample row ranges.
        pack = None
        if bf and self.pack_enabled and Ln > 0 and S <= (96 if save else 160):   # pipelined attention: bwd S <= 96, fwd S <= 160
            ops.pack_index(tmask, imask, ws["cu"], ws["src"])
            Mp = int(ws["cu"][B])                     # the one host read of the step (launch shapes depend on it)
            if save and self.dp is not None:
                self.dp.finalize_plan()              # the row count of the word-row exchange is on the host by now as well
            if Mp < M:
                ops.rows_gather(x, ws["src"], ws["xp"], n=Mp)
                ops.rows_gather(x_lp, ws["src"], ws["xp_lp"], n=Mp)
                ops.rows_gather(kb.view(M, 1), ws["src"], ws["kbp"], n=Mp)
                x, x_lp, kb = ws["xp"], ws["xp_lp"], ws["kbp"]
                pack = (Mp, ws["cu"], ws["src"])
        rows = pack[0] if pack else M
        self.last_rows = rows
        cu = pack[1] if pack else None
        nslots = len(ws["layers"])
        for l in range(Ln):
            sa, so, fi, fo = self.layer_mods[l]
            lw = ws["layers"][l % nslots]
            self._probe_inject(l, x, x_lp, B, S, pack)
            xin, xin_lp = x, x_lp
            self._planned(ws, ("f", l, pack is not None, training, save), rows,
                          lambda: self._layer_fwd(lw, xin, xin_lp, kb, B, S, sa.query, so.dense, so.LayerNorm, fi.dense, fo.dense,
                                                  fo.LayerNorm, eps, 0.125, self._site(2 + 4 * l, p_a, training),
                                                  self._site(3 + 4 * l, p_h, training), self._site(4 + 4 * l, p_h, training), bf,
                                                  rows=rows, cu=cu, save=save))
            x, x_lp = lw["x2"], lw["x2_lp"]
            self._probe_capture(l, x, B, S, pack)
        self._probe_inject(Ln, x, x_lp, B, S, pack)
        logits = self._head_fwd(ws, x_lp, B, S, task_id, self.pooler.dense, self.pool_act,
                                self._site(2 + 4 * Ln, float(self.model.dropout.p), training), eps, bf, pack=pack)
        if save:
            self._saved = dict(ws=ws, ids=ids, seg=seg, fa=fa, la=la, task=task_id, training=training, ticket=self._ticket,
                               B=B, T=T, R=R, feats=feats, locs=locs, pack=pack, kb=kb)
        return logits

    def _planned(self, ws, key, rows, fn):
        """run the launches `fn` issues - recorded on first use, replayed natively afterwards (only the packed row count and
        the dropout seed differ between steps).  Falls back to issuing them one by one when a test hook or the per-GEMM
        timer needs to see the individual launches."""
        if ops.NVTX:
            with ops.nvtx_range("clg.layer%d.%s" % (key[1], "fwd" if key[0] == "f" else "bwd")):
                return self._planned_run(ws, key, rows, fn)
        return self._planned_run(ws, key, rows, fn)

    def _planned_run(self, ws, key, rows, fn):
        if not self.use_plans or self.probe is not None or ops._gemm_timing is not None or ops._plan_rec is not None:
            return fn()
        plans = ws.setdefault("plans", {})
        plan = plans.get(key)
        if plan is None:
            ops.begin_plan()
            try:
                fn()
            finally:
                plan = ops.end_plan()
            plans[key] = plan
        ops.run_plan(plan, rows, self._seed)

    def _drop_plans(self):
        for ws in self._ws.values():
            ws.pop("plans", None)

    def _layer_fwd(self, lw, x, x_lp, kb, B, S, qmod, omod, ln1, f1mod, f2mod, ln2, eps, scale, site_p, site_o, site_f, bf,
                   row_valid=None, rows=None, cu=None, save=True):
        """rows: token rows actually present (B*S, or the packed count with cu = per-sample row ranges)"""
        d, F, H = self.d, self.F, self.H
        M = B * S if rows is None else rows
        ops.gemm(x_lp, self._w(qmod, 3 * d, d), lw["qkv"], bias=self._bias(qmod, 3 * d), M=M)
        ops.attention_fwd(lw["qkv"], kb, lw["ctx"], lw["lse"], B, S, H, scale, site_p, cu_rows=cu, total_rows=M if cu is not None else 0)
        ops.gemm(lw["ctx"], self._w(omod, d, d), lw["y1"], bias=self._bias(omod, d), residual=x, dropout=site_o, M=M)
        ops.layernorm_fwd(lw["y1"], ln1.weight.data, ln1.bias.data, eps, y=lw["x1"], y_lowp=lw["x1_lp"] if bf else None,
                          mean=lw["m1"], rstd=lw["r1"], rows=M)
        # bf16 mode saves gelu'(u) (one shared exponential in the epilogue) so the backward epilogue is a multiply;
        # the fp32 correctness mode saves u and evaluates the exact erf / exp derivative in the backward
        # (forward-only runs do not store gelu': same epilogue, one output)
        ops.gemm(lw["x1_lp"], self._w(f1mod, F, d), lw["h"], bias=self._bias(f1mod, F),
                 act=ops.ACT_GELU_GRAD if bf else ops.ACT_GELU, aux_out=lw["u"] if (save or not bf) else None, M=M)
        ops.gemm(lw["h"], self._w(f2mod, d, F), lw["y2"], bias=self._bias(f2mod, d), residual=lw["x1"], dropout=site_f, M=M)
        ops.layernorm_fwd(lw["y2"], ln2.weight.data, ln2.bias.data, eps, y=lw["x2"], y_lowp=lw["x2_lp"] if bf else None,
                          mean=lw["m2"], rstd=lw["r2"], row_valid=row_valid, valid_post=row_valid is not None, rows=M)
        return lw["x2"], lw["x2_lp"]

    def _first_rows(self, ws, x_lp, B, S, pack):
        """token 0 of every sample as a [B, d] operand: a strided view, or gathered from the packed rows"""
        if pack is None:
            return x_lp.view(B, S * self.d)[:, :self.d]
        ops.rows_gather(x_lp, pack[1], ws["xfirst"], n=B)
        return ws["xfirst"]

    def _head_fwd(self, ws, x_lp, B, S, task_id, pmod, pool_act, site_pool, eps_clf, bf, pack=None):
        d = self.d
        hw = ws["heads"][task_id]
        clf = self.model.clfs_dict[task_id].logit_fc
        P = self.cfg.pooler_size
        x_first = self._first_rows(ws, x_lp, B, S, pack)
        ops.gemm(x_first, self._w(pmod, P, d), ws["pact"], bias=self._bias(pmod, P), act=pool_act)
        ws["pooled32"].copy_(ws["pact"])
        ops.dropout_inplace(ws["pooled32"], site_pool)            # encoders.py:L1238
        if bf:
            ops.cast_f32_to_bf16(ws["pooled32"], ws["pooled"])
        ops.gemm(ws["pooled"], self._w(clf[0], hw["C"], P), hw["hc"], bias=self._bias(clf[0], hw["C"]), act=ops.ACT_GELU,
                 aux_out=hw["uc"])
        ln = clf[2]
        ops.layernorm_fwd(hw["hc"], ln.weight.data, ln.bias.data, float(getattr(ln, "eps", eps_clf)), y_lowp=hw["hln"],
                          mean=hw["mc"], rstd=hw["rc"])
        ops.gemm(hw["hln"], self._w(clf[3], hw["A"], hw["C"]), hw["logits"], bias=self._bias(clf[3], hw["A"]))
        return hw["logits"]

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def _attach_grads(self):
        """Gradient-arena policy: if the caller dropped the gradients (zero_grad(set_to_none=True)) or
        the fused optimiser consumed them, zero the arena and (re)attach the views; otherwise the
        existing .grad views keep accumulating (gradient accumulation, train_task.py:L322-338)."""
        first = self.slots[0].params[0]
        dropped = first.grad is None or first.grad.data_ptr() != self.grad_view(first).data_ptr()
        if dropped or self.grads_consumed:
            if not self.grads_zeroed:
                self.g_flat.zero_()
            lo, hi = self._table_blocks
            self._touched[lo:hi].zero_()
            self._touch_valid = True        # from here on every write into the table gradient is marked (mark_word_rows)
            if dropped:                     # (views that are still attached stay: the step only cleared their contents)
                for s in self.slots:
                    o = s.start
                    for p in s.params:
                        n = p.numel()
                        if p.requires_grad and s.has_grad:
                            p.grad = self.g_flat[o:o + n].view(p.shape)
                        o += n
            self.grads_consumed = False
        self.grads_zeroed = False       # this backward writes the arena

    def backward(self, dlogits, ticket):
        sv = self._saved
        if sv is None or sv["ticket"] != ticket:
            raise ops.ClgError("backward called for a forward whose activations were overwritten by a later forward")
        with torch.no_grad(), ops.nvtx_range("clg.backward"):
            self._attach_grads()
            self._hot_valid = False          # new gradients: the block marks are stale until the next norm pass
            dl = dlogits.contiguous().float()
            if self.family == "m3p":
                self._backward_m3p(dl, sv)
            else:
                self._backward_uc2(dl, sv)
            if self.dp is not None:
                with ops.nvtx_range("clg.exchange"):
                    self.dp.after_backward(self)
        self._saved = None

    def _wgrad(self, a, b, mod0, rows, cols, tokens=None):
        """dW[rows, cols] += a^T b  (a: [tokens, rows], b: [tokens, cols]); SFT mask in the epilogue."""
        s = self.w_slot[id(mod0)]
        ops.gemm(a, b, self._g(s, rows, cols), a_kmajor=False, b_kmajor=False, accumulate=True,
                 mask_bits=self._wmask(mod0), K=tokens)

    def _wgrad2(self, a, b, mod0, rows, cols, a2, b2, mod2, rows2, tokens=None):
        """Two weight gradients with the same input width in one launch (rows % 256 == 0): together the QKV and
        attention-output gradients have the tile count of one FFN gradient, which fills the machine."""
        ops.gemm(a, b, self._g(self.w_slot[id(mod0)], rows, cols), a_kmajor=False, b_kmajor=False, accumulate=True,
                 mask_bits=self._wmask(mod0), K=tokens,
                 group2=(a2, b2, self._g(self.w_slot[id(mod2)], rows2, cols), self._wmask(mod2)))

    def _head_bwd(self, ws, sv, dl, x_lp, pmod, pool_dact, site_pool):
        d = self.d
        B, S = sv["B"], sv["T"] + sv["R"]
        task = sv["task"]
        hw = ws["heads"][task]
        clf = self.model.clfs_dict[task].logit_fc
        P = self.cfg.pooler_size
        C, A = hw["C"], hw["A"]
        bf = ws["lp"] == torch.bfloat16
        ops.colsum(dl, self._gvec(self.b_slot[id(clf[3])], A))
        if bf:      # bf16 operand copies: every GEMM of the head backward runs on the tcgen05 path
            ops.cast_f32_to_bf16_2d(dl, hw["dl_lp"])
            dl = hw["dl_lp"]
        self._wgrad(dl, hw["hln"], clf[3], A, C)
        ops.gemm(dl, self._w(clf[3], A, C), hw["dhln"], b_kmajor=False)
        ln = clf[2]
        ops.layernorm_bwd(hw["dhln"], hw["hc"], ln.weight.data, hw["mc"], hw["rc"], dx=hw["dhc"],
                          dgamma=self.grad_view(ln.weight), dbeta=self.grad_view(ln.bias))
        ops.act_bwd(hw["dhc"], hw["uc"], hw["dhc"], ops.ACT_DGELU)
        ops.colsum(hw["dhc"], self._gvec(self.b_slot[id(clf[0])], C))
        dhc = hw["dhc"]
        if bf:
            ops.cast_f32_to_bf16(hw["dhc"], hw["dhc_lp"])
            dhc = hw["dhc_lp"]
        self._wgrad(dhc, ws["pooled"], clf[0], C, P)
        ops.gemm(dhc, self._w(clf[0], C, P), ws["dpre"], b_kmajor=False, dropout=site_pool, drop_before_act=True,
                 act=pool_dact, aux_in=ws["pact"])
        ops.colsum(ws["dpre"], self._gvec(self.b_slot[id(pmod)], P))
        dpre = ws["dpre"]
        if bf:
            ops.cast_f32_to_bf16(ws["dpre"], ws["dpre_lp"])
            dpre = ws["dpre_lp"]
        pack = sv.get("pack")
        x_first = x_lp.view(B, S * d)[:, :d] if pack is None else ws["xfirst"]
        self._wgrad(dpre, x_first, pmod, P, d)
        ws["dx"].zero_()
        if pack is None:
            ops.gemm(dpre, self._w(pmod, P, d), ws["dx"].view(B, S * d)[:, :d], b_kmajor=False)
        else:
            ops.gemm(dpre, self._w(pmod, P, d), ws["dxfirst"], b_kmajor=False)
            ops.rows_scatter(ws["dxfirst"], pack[1], ws["dx"], n=B)

    def _layer_bwd(self, ws, lw, x_in, x_in_lp, kb, B, S, qmod, omod, ln1, f1mod, f2mod, ln2, scale, site_p, site_o, site_f,
                   row_valid=None, rows=None, cu=None):
        """dx (ws['dx']) holds d(loss)/d(layer output) on entry and d(loss)/d(layer input) on exit."""
        d, F, H = self.d, self.F, self.H
        M = B * S if rows is None else rows
        dx, dy, dgo = ws["dx"], ws["dy"], ws["dgo"]
        # LN2 -> residual grad dy, dropout-masked operand copy dgo, LN and FFN2-bias grads
        ops.layernorm_bwd(dx, lw["y2"], ln2.weight.data, lw["m2"], lw["r2"], dx=dy, dx_lowp=dgo,
                          dgamma=self.grad_view(ln2.weight), dbeta=self.grad_view(ln2.bias),
                          dbias=self._gvec(self.b_slot[id(f2mod)], d), dropout_in=site_f,
                          row_valid=row_valid, valid_post=row_valid is not None, rows=M)
        self._wgrad(dgo, lw["h"], f2mod, d, F, tokens=M)
        if ws["lp"] == torch.bfloat16:      # FFN1 bias gradient = column sums of dU, reduced inside the GEMM epilogue
            ops.gemm(dgo, self._w(f2mod, d, F), ws["du"], b_kmajor=False, act=ops.ACT_MUL_AUX, aux_in=lw["u"],
                     colsum=self._gvec(self.b_slot[id(f1mod)], F), M=M)
        else:
            ops.gemm(dgo, self._w(f2mod, d, F), ws["du"], b_kmajor=False, act=ops.ACT_DGELU, aux_in=lw["u"], M=M)
            ops.colsum(ws["du"], self._gvec(self.b_slot[id(f1mod)], F), M=M)
        self._wgrad(ws["du"], lw["x1_lp"], f1mod, F, d, tokens=M)
        ops.gemm(ws["du"], self._w(f1mod, F, d), dx, b_kmajor=False, residual=dy, M=M)
        # LN1
        ops.layernorm_bwd(dx, lw["y1"], ln1.weight.data, lw["m1"], lw["r1"], dx=dy, dx_lowp=dgo,
                          dgamma=self.grad_view(ln1.weight), dbeta=self.grad_view(ln1.bias),
                          dbias=self._gvec(self.b_slot[id(omod)], d), dropout_in=site_o, rows=M)
        grouped = self.group_wgrad and ws["lp"] == torch.bfloat16 and (3 * d) % 256 == 0
        if not grouped:
            self._wgrad(dgo, lw["ctx"], omod, d, d, tokens=M)
        ops.gemm(dgo, self._w(omod, d, d), ws["dctx"], b_kmajor=False, M=M)
        ops.attention_bwd(lw["qkv"], kb, lw["ctx"], ws["dctx"], lw["lse"], ws["dqkv"], B, S, H, scale, site_p,
                          cu_rows=cu, total_rows=M if cu is not None else 0)
        ops.colsum(ws["dqkv"], self._gvec(self.b_slot[id(qmod)], 3 * d), M=M)
        if grouped:     # dgo still holds the attention-output operand: nothing between LN1's backward and here writes it
            self._wgrad2(ws["dqkv"], x_in_lp, qmod, 3 * d, d, dgo, lw["ctx"], omod, d, tokens=M)
        else:
            self._wgrad(ws["dqkv"], x_in_lp, qmod, 3 * d, d, tokens=M)
        ops.gemm(ws["dqkv"], self._w(qmod, 3 * d, d), dx, b_kmajor=False, residual=dy, M=M)

    def _backward_uc2(self, dl, sv):
        cfg, d, Ln = self.cfg, self.d, self.L
        ws = sv["ws"]
        B, T, R = sv["B"], sv["T"], sv["R"]
        S = T + R
        training = sv["training"]
        E = self.model.bert.embeddings
        p_h = float(cfg.hidden_dropout_prob)
        p_a = float(cfg.attention_probs_dropout_prob)
        bf = ws["lp"] == torch.bfloat16
        last = ws["layers"][(Ln - 1) % len(ws["layers"])] if Ln > 0 else None
        x_last_lp = last["x2_lp"] if last is not None else ws["x0_lp"]
        dact = ops.ACT_DRELU if self.pool_act == ops.ACT_RELU else ops.ACT_DTANH
        self._head_bwd(ws, sv, dl, x_last_lp, self.pooler.dense, dact,
                       self._site(2 + 4 * Ln, float(self.model.dropout.p), training))
        pack = sv.get("pack")
        rows = pack[0] if pack else None
        cu = pack[1] if pack else None
        kb = sv.get("kb", ws["kb"])
        self._probe_inject_grad(Ln, ws["dx"], B, S, pack)
        for l in reversed(range(Ln)):
            sa, so, fi, fo = self.layer_mods[l]
            lw = ws["layers"][l]
            if l == 0:
                x_in, x_in_lp = (ws["xp"], ws["xp_lp"]) if pack else (ws["x0"], ws["x0_lp"])
            else:
                x_in, x_in_lp = ws["layers"][l - 1]["x2"], ws["layers"][l - 1]["x2_lp"]
            self._planned(ws, ("b", l, pack is not None, training, self.group_wgrad), rows if rows is not None else B * S,
                          lambda: self._layer_bwd(ws, lw, x_in, x_in_lp, kb, B, S, sa.query, so.dense, so.LayerNorm, fi.dense,
                                                  fo.dense, fo.LayerNorm, 0.125, self._site(2 + 4 * l, p_a, training),
                                                  self._site(3 + 4 * l, p_h, training), self._site(4 + 4 * l, p_h, training),
                                                  rows=rows, cu=cu))
            self._probe_inject_grad(l, ws["dx"], B, S, pack)
            if self.dp is not None:
                self.dp.layer_done(self, l)
        dx = ws["dx"]
        if pack:                                 # gradients of the kept rows back to their [B, S] positions; masked rows: 0
            ws["dxd"].zero_()
            ops.rows_scatter(dx, pack[2], ws["dxd"], n=pack[0])
            dx = ws["dxd"]
        # ---- text embeddings
        ttab = E.new_token_type_embeddings.weight
        # the LN backward also emits the column sums of de_txt over all tokens into row 0 of the token-type gradient
        # (every GQA token has segment 0); the scatter kernel only corrects for segment-1 tokens
        ops.layernorm_bwd(dx, ws["e_txt"], E.LayerNorm.weight.data, ws["m_txt"], ws["r_txt"], dx=ws["de_txt"],
                          dgamma=self.grad_view(E.LayerNorm.weight), dbeta=self.grad_view(E.LayerNorm.bias),
                          dbias=self.grad_view(ttab)[0], dy_map=rowmap(T, 0, S), dropout_out=self._site(0, p_h, training))
        rows_mode = self.dp is not None and self.dp.rows_mode()     # data parallel: the word rows are reduced after the all-gather
        ops.embed_text_bwd(sv["ids"], sv["seg"], ws["de_txt"], None if rows_mode else self.grad_view(E.word_embeddings.weight),
                           self.grad_view(E.position_embeddings.weight), self.grad_view(ttab), pad_id=cfg.pad_token_id,
                           type_row0_has_total=True)
        if rows_mode:
            self.dp.add_rows(sv["ids"], ws["de_txt"])
            self.dp.rows_ready(self)                # pre-reduction + all-gather start here, under the region half below
        else:
            self.mark_word_rows(sv["ids"])
        # ---- region embeddings
        ops.layernorm_bwd(dx, ws["t2"], E.v_LayerNorm.weight.data, ws["m_v"], ws["r_v"], dx=ws["ds"],
                          dgamma=self.grad_view(E.v_LayerNorm.weight), dbeta=self.grad_view(E.v_LayerNorm.bias),
                          dbias=self.grad_view(ttab)[1], dy_map=rowmap(R, T, S), dropout_out=self._site(1, p_h, training))
        ops.layernorm_bwd(ws["ds"], ws["g_img"], E.image_layer_norm.weight.data, ws["m_img"], ws["r_img"], dx_lowp=ws["dg"],
                          dgamma=self.grad_view(E.image_layer_norm.weight), dbeta=self.grad_view(E.image_layer_norm.bias),
                          dbias=self._gvec(self.b_slot[id(E.image_embeddings)], d))
        self._wgrad(ws["dg"], sv["fa"], E.image_embeddings, d, sv["fa"].shape[1])
        ops.layernorm_bwd(ws["ds"], ws["l_loc"], E.image_location_layer_norm.weight.data, ws["m_loc"], ws["r_loc"],
                          dx_lowp=ws["dl"], dgamma=self.grad_view(E.image_location_layer_norm.weight),
                          dbeta=self.grad_view(E.image_location_layer_norm.bias),
                          dbias=self._gvec(self.b_slot[id(E.image_location_embeddings)], d))
        self._wgrad(ws["dl"], sv["la"], E.image_location_embeddings, d, sv["la"].shape[1])

    # ------------------------------------------------------------------------------------------
    # hooks for the data-parallel exchange (parallel.GradExchange)
    # ------------------------------------------------------------------------------------------
    def word_table_param(self):
        if self.family == "uc2":
            return self.model.bert.embeddings.word_embeddings.weight
        return self.model.bert.encoder.embeddings.weight

    def mark_word_rows(self, ids: torch.Tensor):
        """record which blocks of the word-table gradient the rows of `ids` touch (see adopt / optim.clip_grad_norm_)"""
        word = self.word_table_param()
        wslot = self._slot_of_param[id(word)]
        pad = int(self.cfg.pad_token_id) if self.family == "uc2" else int(getattr(self.cfg, "pad_index", 1))
        ops.mark_rows(ids, word.shape[1], wslot.start, word.shape[0], pad, self._touched)

    def reduce_word_rows(self, ids: torch.Tensor, rows: torch.Tensor):
        """word-table gradient += the (token id, gradient row) list, reduced deterministically: sorted by id (stable), one
        warp per distinct id summing its rows in list order.  Called by parallel.GradExchange with the all-gathered list."""
        gword = self.grad_view(self.word_table_param())
        pad = int(self.cfg.pad_token_id) if self.family == "uc2" else int(getattr(self.cfg, "pad_index", 1))
        sorted_ids, order = torch.sort(ids, stable=True)           # index bookkeeping only; no host sync
        ops.segment_sum_rows(rows, order, sorted_ids, gword, pad)
        self.mark_word_rows(ids)

    def prereduce_word_rows(self, ids: torch.Tensor, rows: torch.Tensor):
        """local step before the all-gather: duplicate token ids of this rank are summed into one row (in place) and the
        other copies are renamed to the padding id, so the gathered list has at most world-size rows per id
        (<s> / </s> occur once per sample: 2048-long runs on 8 GPUs otherwise).  Returns the ids to send."""
        pad = int(self.cfg.pad_token_id) if self.family == "uc2" else int(getattr(self.cfg, "pad_index", 1))
        ids_out = ids.clone()
        sorted_ids, order = torch.sort(ids, stable=True)
        ops.segment_sum_rows(rows, order, sorted_ids, None, pad, ids_out=ids_out)
        return ids_out

    # ------------------------------------------------------------------------------------------
    # M3P (filled in by engine_m3p.py)
    # ------------------------------------------------------------------------------------------
    def _collect_m3p(self):
        from . import engine_m3p
        engine_m3p.collect(self)

    def _forward_m3p(self, *a, **k):
        from . import engine_m3p
        return engine_m3p.forward(self, *a, **k)

    def _backward_m3p(self, dl, sv):
        from . import engine_m3p
        return engine_m3p.backward(self, dl, sv)
