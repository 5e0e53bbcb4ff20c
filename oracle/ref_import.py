"""TEST INFRASTRUCTURE ONLY.  Import the UNMODIFIED reference (`compressai`) from /root/reference.

Used by oracle/gen_golden.py (to record golden vectors) and by the local-only pin tests; nothing
in stf_b200/, bench.py's product arm or the `-m gpu` tests touches this (the GPU box has no
/root/reference).

The reference needs three things that are not importable as-is:
  * compressai.ans / compressai._CXX  -> the reference's own C++ compiled by oracle/Makefile into
    oracle/_ref (reference setup.py:48-82); registered in sys.modules under the names the
    reference imports (entropy_models.py:13, stf.py:10).
  * timm==0.4.12 (requirements.txt:115) -> a shim of the three symbols the reference uses
    (stf.py:5): to_2tuple, trunc_normal_, DropPath (per-sample Bernoulli keep, scaled by 1/keep,
    identity in eval -- timm/models/layers/drop.py of that release).
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO_DIR = os.path.join(_HERE, "_ref")
# /root/reference in the build container; on the GPU box the byte-for-byte copy that `make -C oracle` left in
# oracle/_ref/compressai (a git-ignored build output, see oracle/Makefile)
REF_ROOT = os.environ.get("STF_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "compressai")) and os.path.isdir(os.path.join(REF_SO_DIR, "compressai")):
    REF_ROOT = REF_SO_DIR


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "compressai"))


def _load_ext(fullname: str, stem: str):
    suffix = importlib.machinery.EXTENSION_SUFFIXES[0]
    path = os.path.join(REF_SO_DIR, stem + suffix)
    if not os.path.exists(path):
        raise ImportError(f"{path} missing: run `make -C oracle` where /root/reference exists")
    loader = importlib.machinery.ExtensionFileLoader(fullname, path)
    spec = importlib.util.spec_from_loader(fullname, loader, origin=path)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def load_ref_ans():
    """The reference's compiled rANS extension alone (travels to the GPU box in oracle/_ref)."""
    if "compressai.ans" in sys.modules:
        return sys.modules["compressai.ans"]
    return _load_ext("compressai.ans", "ans")


def load_ref_cxx():
    if "compressai._CXX" in sys.modules:
        return sys.modules["compressai._CXX"]
    return _load_ext("compressai._CXX", "_CXX")


def _install_timm_shim():
    if "timm.models.layers" in sys.modules:
        return
    import torch
    import torch.nn as nn

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    class DropPath(nn.Module):
        def __init__(self, drop_prob=None):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if not self.drop_prob or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x.div(keep) * mask

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.to_2tuple, layers.trunc_normal_, layers.DropPath = to_2tuple, trunc_normal_, DropPath
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def import_reference():
    """Return the reference `compressai` package (imported from REF_ROOT, unmodified)."""
    if not reference_available():
        raise ImportError(f"{REF_ROOT} not present")
    if "compressai" in sys.modules and getattr(sys.modules["compressai"], "__file__", "").startswith(REF_ROOT):
        return sys.modules["compressai"]
    _install_timm_shim()
    sys.modules["compressai.ans"] = _load_ext("compressai.ans", "ans")
    sys.modules["compressai._CXX"] = _load_ext("compressai._CXX", "_CXX")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import compressai  # noqa: E402

    return compressai
