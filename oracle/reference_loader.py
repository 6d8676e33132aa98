"""TEST INFRASTRUCTURE (oracle) -- imports the reference's own models/sit.py and models/mpp.py, unmodified,
from /root/reference (present in the build container only) with ``oracle.vit_shim`` standing in for the absent
``vit_pytorch`` package.  Used to validate ``sit_oracle`` and to generate tests/golden fixtures.
"""
import importlib.util
import os
import sys

from . import vit_shim

REFERENCE_ROOT = os.environ.get("SVIT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "sit.py"))


def _load(name, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_models():
    """Returns (SiT, masked_patch_pretraining, mpp_module) classes of the reference."""
    if not available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    vit_shim.install_as_vit_pytorch()
    sit = _load("_svit_reference_sit", "models/sit.py")
    mpp = _load("_svit_reference_mpp", "models/mpp.py")
    return sit.SiT, mpp.masked_patch_pretraining, mpp
