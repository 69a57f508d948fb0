"""Golden vectors for the optimiser step from code that lives in the reference tree.  TEST INFRASTRUCTURE ONLY.
Run in the build container:  python -m oracle.make_golden_adamw

The fine-tune loop takes `AdamW` from the un-vendored pip dependency pytorch-transformers (train_task.py:L23, L263-267).
The reference tree nevertheless contains that optimiser's step body verbatim:
/root/reference/features_extraction/mmf/mmf/modules/optimizers.py:L15-84 (`AdamWSkipParamsWithZeroGrad.step`, "modified
from huggingface/transformers ... optimization.py#L259-L318" - the HuggingFace AdamW, whose algorithm is the
pytorch-transformers 1.x one unchanged: eps added before the bias correction, decoupled decay after the update; the only
modification is that parameters with an all-zero gradient are skipped).  That method is taken out of the file with `ast`
and executed UNMODIFIED (bound to a plain torch Optimizer that supplies the param-group defaults AdamW.__init__ would) on
seeded parameters and gradients for several steps, with and without bias correction / weight decay; the trajectories are
stored in tests/golden/adamw_ref_step.npz.  tests/test_oracle_golden.py pins oracle/adamw_oracle.py and
clg_vqa_b200.optim.AdamW (per-parameter path) against them; the GPU tests pin the fused kernel against the same file.
"""
from __future__ import annotations

import ast
import math
import os
import sys
from typing import Callable

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = "/root/reference/features_extraction/mmf/mmf/modules/optimizers.py"
OUT = os.path.join(ROOT, "tests", "golden", "adamw_ref_step.npz")


def reference_step_function():
    """the `step` method of AdamWSkipParamsWithZeroGrad, compiled from the reference file as it stands"""
    tree = ast.parse(open(SRC).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "AdamWSkipParamsWithZeroGrad")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "step")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"math": math, "torch": torch, "Callable": Callable}
    exec(compile(mod, SRC, "exec"), ns)
    return ns["step"]


class _Host(torch.optim.Optimizer):
    """param groups with the defaults of the HuggingFace / pytorch-transformers AdamW constructor"""

    def __init__(self, params, lr, betas, eps, weight_decay, correct_bias):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))


CASES = {
    # name: (lr, betas, eps, weight_decay, correct_bias)
    "default_correct_bias": (4e-5, (0.9, 0.999), 1e-6, 1e-4, True),
    "no_bias_correction": (4e-5, (0.9, 0.98), 1e-6, 1e-2, False),       # --adam_correct_bias off (the scripts' default)
    "no_decay": (1e-4, (0.9, 0.999), 1e-6, 0.0, True),
    "big_eps": (1e-3, (0.8, 0.99), 1e-3, 5e-2, True),
}
STEPS = 6
SHAPE = (37, 19)


def main():
    step = reference_step_function()
    out = {}
    for name, (lr, betas, eps, wd, cb) in CASES.items():
        g = torch.Generator().manual_seed(1234)
        p = torch.randn(SHAPE, generator=g) * 0.05
        p = p.clone().requires_grad_(True)
        host = _Host([p], lr, betas, eps, wd, cb)
        grads, traj = [], []
        for t in range(STEPS):
            gr = torch.randn(SHAPE, generator=g) * (0.02 if t != 3 else 2e-5)      # one step with tiny gradients (eps matters)
            p.grad = gr.clone()
            step(host)
            grads.append(gr.numpy().copy())
            traj.append(p.detach().numpy().copy())
        g0 = torch.Generator().manual_seed(1234)
        out[name + "::p0"] = (torch.randn(SHAPE, generator=g0) * 0.05).numpy()
        out[name + "::grads"] = np.stack(grads)
        out[name + "::params"] = np.stack(traj)
        out[name + "::hyper"] = np.array([lr, betas[0], betas[1], eps, wd, float(cb)], dtype=np.float64)
        st = host.state[p]
        out[name + "::exp_avg"] = st["exp_avg"].numpy().copy()
        out[name + "::exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    out["source"] = np.array(SRC + ":L15-84 (AdamWSkipParamsWithZeroGrad.step, executed unmodified via ast)")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 0})


if __name__ == "__main__":
    sys.exit(main())
