"""TEST INFRASTRUCTURE ONLY (oracle/): loader for the *real* reference under /root/reference.

Used only in the build container (the reference tree does not exist on the GPU box) to
(a) validate the CPU restatement in oracle/skit_oracle.py and (b) generate the golden
fixtures committed under tests/golden/ (see oracle/make_golden.py).

The reference imports a dozen packages that are absent here; they are
replaced by permissive stub modules *before* the reference is imported.  Nothing from the
reference is copied: it is imported in place from its read-only mount.
"""
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("VTS_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    """Module whose every attribute is a no-op callable / empty class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _noop(*a, **k):
            return None

        return _noop


def _install_stub(name, **attrs):
    m = _Stub(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


_LOADED = False


def load_reference():
    """Make `import models.networks` etc. resolve to the reference tree."""
    global _LOADED
    if _LOADED:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch.nn as nn

    class _LPIPS(nn.Module):  # lpips.LPIPS is constructed unconditionally by the model
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, a, b):
            raise RuntimeError("LPIPS is out of scope (SURVEY.md §8f); run with lambda_*_lpips 0")

    for name in ["tkinter", "turtle", "vision_aided_loss", "clip", "gspread",
                 "oauth2client", "oauth2client.service_account", "matplotlib",
                 "matplotlib.pyplot", "matplotlib.patches", "dominate", "dominate.tags",
                 "visdom", "wandb", "GPUtil"]:
        if name not in sys.modules:
            _install_stub(name)
    if "lpips" not in sys.modules:
        _install_stub("lpips", LPIPS=_LPIPS)
    if "torchmetrics" not in sys.modules:
        _install_stub("torchmetrics", MeanSquaredError=object)
        _install_stub("torchmetrics.functional")
    try:
        import torchvision.models.resnet as _r
        if not hasattr(_r, "model_urls"):
            _r.model_urls = {}
    except Exception:
        pass
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _LOADED = True


def ref_networks():
    load_reference()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        import models.networks as N  # noqa
    return N
