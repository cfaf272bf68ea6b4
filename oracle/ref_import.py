"""Import the UNMODIFIED reference package from /root/reference with import-time stubs.

TEST INFRASTRUCTURE ONLY.  This works only inside the build container (the GPU
box has no /root/reference); it is used by tests/golden/make_golden.py to
generate the committed golden vectors and by the `-m "not gpu"` pinning tests
when the reference happens to be present.

The reference needs pytorch_lightning, skimage, jpeg4py and lpips at import
time (master_thesis/utils.py:6-7, model_dfpn.py:6, dataset.py); none is in the
image, none is touched by the hot path, so they are replaced by empty modules.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MT_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "master_thesis"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def import_reference():
    """Returns the reference ``master_thesis`` package (CPU, unmodified)."""
    if "master_thesis" in sys.modules:
        return sys.modules["master_thesis"]
    if not reference_available():
        raise ImportError("reference not present at %s" % REFERENCE_ROOT)
    import torch.nn as nn

    class _LightningModule(nn.Module):
        def log(self, *a, **k):
            pass

    class _Dummy(object):
        def __init__(self, *a, **k):
            pass

    _stub("pytorch_lightning", LightningModule=_LightningModule,
          LightningDataModule=_Dummy, Trainer=_Dummy)
    sk = _stub("skimage")
    sk.metrics = _stub("skimage.metrics")
    sk.transform = _stub("skimage.transform")
    _stub("jpeg4py")
    _stub("lpips")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import master_thesis  # noqa: E402
    return master_thesis
