"""Run the reference's OWN implementation of the shift-layer path on CPU.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py --impl reference, cpu_baseline, golden generation, tests).
The reference modules are imported unmodified from ``oracle/_ref/`` (staged by ``oracle/build_ref.py``; git-ignored,
travels to the GPU box) or, in the build container, from ``/root/reference``.  The reference tests
``torch.cuda.is_available`` without calling it (models/IPSRFunction.py:28,38, util/NonparametricShift.py:15,
models/InnerCos.py:19, models/InnerCos2.py:22) and calls ``.cuda()`` unconditionally, so it only runs on a CPU with
the two-line shim of SURVEY.md appendix B, applied inside ``cpu_shim()`` and undone afterwards (the product's CUDA
tensors must not be affected).
"""
import collections
import contextlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")


def reference_dir():
    """Directory holding the reference's ``models/`` and ``util/`` packages, or None."""
    for cand in (os.environ.get("IPSR_REFERENCE"), STAGED, "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "models", "IPSRFunction.py")):
            return cand
    return None


def available() -> bool:
    return reference_dir() is not None


@contextlib.contextmanager
def cpu_shim():
    """torch.cuda.FloatTensor -> torch.FloatTensor and Tensor.cuda() -> identity while the reference runs on CPU."""
    import torch
    had = hasattr(torch.cuda, "FloatTensor")
    old_ft = getattr(torch.cuda, "FloatTensor", None)
    old_cuda = torch.Tensor.cuda
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = old_cuda
        if had:
            torch.cuda.FloatTensor = old_ft
        else:
            del torch.cuda.FloatTensor


_mods = None


def modules():
    """(IPSR_model, InnerCos, InnerCos2, IPSRFunction module, util.util module) of the unmodified reference.
    The reference's packages are called ``models`` and ``util``: they are imported under those names from the staged
    directory and then removed from ``sys.path`` again."""
    global _mods
    if _mods is not None:
        return _mods
    root = reference_dir()
    if root is None:
        raise RuntimeError("the reference is not staged: run `python oracle/build_ref.py` where /root/reference exists")
    for name in ("models", "util"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(root):
            raise RuntimeError("a foreign top-level package %r is already imported; cannot import the reference" % name)
    sys.path.insert(0, root)
    try:
        with cpu_shim():
            from models.IPSR_model import IPSR_model          # noqa
            from models.InnerCos import InnerCos              # noqa
            from models.InnerCos2 import InnerCos2            # noqa
            import models.IPSRFunction as F                   # noqa
            import util.util as U                             # noqa
    finally:
        sys.path.remove(root)
    _mods = (IPSR_model, InnerCos, InnerCos2, F, U)
    return _mods


Ref = collections.namedtuple("Ref", ["relu4_3"])


def shift_fwd_bwd(x, ref, g, mask_global, triple_w=1.0, shift_sz=1, stride=1, mask_thred=1, threshold=5 / 16.0):
    """One forward + backward of the reference's ``IPSR_model`` on CPU tensors (models/IPSR_model.py:42-63,
    models/IPSRFunction.py:13-178).  Returns (out, grad_input, module)."""
    IPSR_model = modules()[0]
    with cpu_shim():
        m = IPSR_model(threshold, 1, shift_sz, stride, mask_thred, triple_w)
        m.set_mask(mask_global, 3, threshold)
        m.set_ref(Ref(ref))
        xin = x.detach().clone().requires_grad_(True)
        y = m(xin)
        y.backward(g)
    return y.detach(), xin.grad, m
