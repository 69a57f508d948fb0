"""Import the UNMODIFIED reference (volta.encoders / volta.task_utils) from /root/reference.
TEST INFRASTRUCTURE ONLY; usable only in the build container (the GPU box has no /root/reference).

Recipe = SURVEY.md Appendix D.1-2: stub the unused third-party imports (boto3, botocore,
lmdb, tensorpack, msgpack_numpy, tensorboardX, easydict), pre-register an empty `volta.datasets`
package because the shipped volta/volta/datasets/__init__.py raises NameError, and neutralise the
hard-coded `.cuda()` calls of task_utils.py:L309, L708 for the CPU run.  Nothing is copied.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CLG_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "volta", "volta"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_imported = {}


def import_reference():
    """Returns a namespace with BertForVLTasks, M3PForVLTasks, BertConfig, M3PConfig, task_utils."""
    if _imported:
        return types.SimpleNamespace(**_imported)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("boto3")
    _stub("botocore")
    _stub("botocore.exceptions", ClientError=type("ClientError", (Exception,), {}))
    _stub("lmdb")
    tp = _stub("tensorpack")
    tp.dataflow = _stub("tensorpack.dataflow", RNGDataFlow=object, PrefetchDataZMQ=object,
                        LMDBSerializer=object, BatchData=object)
    _stub("msgpack_numpy", patch=lambda: None)
    _stub("tensorboardX", SummaryWriter=object)

    class _EasyDict(dict):
        __getattr__ = dict.get
        __setattr__ = dict.__setitem__
    _stub("easydict", EasyDict=_EasyDict)

    vroot = os.path.join(REFERENCE_ROOT, "volta")
    if vroot not in sys.path:
        sys.path.insert(0, vroot)
    import volta  # noqa: F401  (the reference package)
    if "volta.datasets" not in sys.modules or not hasattr(sys.modules["volta.datasets"], "DatasetMapTrain"):
        pkg = types.ModuleType("volta.datasets")
        pkg.__path__ = [os.path.join(vroot, "volta", "datasets")]
        pkg.DatasetMapTrain = {}
        pkg.DatasetMapEval = {}
        sys.modules["volta.datasets"] = pkg
        lb = types.ModuleType("volta.datasets._image_features_reader")
        lb.ImageFeaturesH5Reader = object
        sys.modules["volta.datasets._image_features_reader"] = lb
    from volta.config import BertConfig, M3PConfig
    from volta.encoders import BertForVLTasks, M3PForVLTasks
    try:
        import volta.task_utils as task_utils
    except Exception as e:  # pragma: no cover - reported by the caller
        task_utils = None
        _imported["task_utils_error"] = repr(e)
    _imported.update(BertConfig=BertConfig, M3PConfig=M3PConfig, BertForVLTasks=BertForVLTasks,
                     M3PForVLTasks=M3PForVLTasks, task_utils=task_utils)
    return types.SimpleNamespace(**_imported)


@contextlib.contextmanager
def cpu_cuda_noop():
    """task_utils.py calls `.cuda()` unconditionally (L309, L708); make it the identity on CPU."""
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


@contextlib.contextmanager
def bf16_operand_shim():
    """round the operands of F.linear and torch.matmul to bf16 and run the
    unmodified fp32 reference ("bf16 GEMM operands, fp32 accumulate, fp32 everything else")."""
    import torch
    import torch.nn.functional as F
    lin, mm = F.linear, torch.matmul

    def r(x):
        return x.bfloat16().to(x.dtype) if x.is_floating_point() else x

    def linear(x, w, b=None):
        return lin(r(x), r(w), b)

    def matmul(a, b):
        return mm(r(a), r(b))

    F.linear, torch.matmul = linear, matmul
    try:
        yield
    finally:
        F.linear, torch.matmul = lin, mm


@contextlib.contextmanager
def bf16_train_shim():
    """bf16 operand rounding in the forward AND the backward contractions (oracle/bf16_autograd.py) over the
    unmodified reference: the reference a bf16 training step's gradients are compared with."""
    import torch
    import torch.nn.functional as F
    from . import bf16_autograd as ba
    lin, mm = F.linear, torch.matmul
    F.linear, torch.matmul = ba.linear_bf16, ba.matmul_bf16
    try:
        yield
    finally:
        F.linear, torch.matmul = lin, mm
