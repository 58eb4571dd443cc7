"""Path-import of the real reference modules (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box; nothing in
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this at run time.  It is used by
``tests/golden/make_golden.py`` (to generate fixtures) and by the ``not gpu`` tests
that cross-check the oracle layers against the live reference when it is present.

``import src.models`` fails in this image (timm / torch_geometric missing)
because ``src/models/__init__.py:1-6`` star-imports every model; ``base.py`` and
``loss.py`` load fine on their own by file path.
"""
import importlib.util
import os
import sys

_CANDIDATES = [os.environ.get("IA_REFERENCE_DIR"), "/root/reference"]


def reference_dir():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "src", "models", "base.py")):
            return c
    return None


def available():
    return reference_dir() is not None


def _load(name, rel):
    root = reference_dir()
    if root is None:
        raise RuntimeError("reference tree not present (expected on the build container only)")
    key = "_ia_ref_" + name
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(root, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


def base():
    """reference src/models/base.py (InnerProduct, VecSimClassificationHead, TwoTowerClassificationHead)."""
    return _load("base", "src/models/base.py")


def loss():
    """reference src/models/loss.py (HingeLoss, EuclideanDistanceLoss)."""
    return _load("loss", "src/models/loss.py")
