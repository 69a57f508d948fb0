"""CPU oracle for the CLG-VQA hot path.  TEST INFRASTRUCTURE ONLY.

Plain-PyTorch (CPU, fp32 / fp64) restatement of the reference's algorithm for the VOLTA
single-stream encoder step: UC2 (`uc2_forward`), M3P (`m3p_forward`), the GQA label-prior
loss (`gqa_prior_loss`) and the pytorch-transformers AdamW / warm-up schedule / clip
(`adamw`).  Every function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import this package, and only as the checker or the timed CPU arm - never from the
product path (`clg_vqa_b200/`), which fails loudly when its CUDA library is missing.

Pinning status
  * uc2_forward / m3p_forward / gqa_prior_loss: PINNED against the unmodified reference
    (`volta.encoders.BertForVLTasks`, `M3PForVLTasks`, `volta.task_utils.ForwardModelsTrain`)
    imported from /root/reference in the build container by `oracle/make_golden.py`; the
    outputs of the reference itself are committed as fixtures under `tests/golden/` and
    `tests/test_oracle_golden.py` re-checks the oracle against them on every run.
    The reference ships no golden vectors or tests of its own for this path.
  * adamw (optimiser / schedule): PINNED (round 2).  The drivers import the optimiser from the un-vendored
    pip dependency `pytorch-transformers>=1.1.0` (volta/requirements.txt:L51; call sites
    train_task.py:L263-276, L326-338), but the reference tree carries that optimiser's step body verbatim
    (features_extraction/mmf/mmf/modules/optimizers.py:L15-84).  `oracle/make_golden_adamw.py` executes it
    unmodified on seeded parameters / gradients and commits the trajectories
    (`tests/golden/adamw_ref_step.npz`); the oracle and `clg_vqa_b200.optim.AdamW` reproduce them bit for
    bit on the CPU, the fused kernel within 2e-6 on the GPU.  The warm-up schedule is pinned against the
    installed `transformers.get_linear_schedule_with_warmup` (the same lambda).
"""
