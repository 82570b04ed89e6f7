"""Loads a module of the reference's concept_vit/ directory under a private name, so a shim module of the same public
name can re-export it with the hot-path functions swapped (shim/CLIP_og_utils.py, shim/utils.py, shim/og_utils.py)."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)


def reference_dir():
    return os.environ.get("MCD_REFERENCE_DIR", "/root/reference/concept_vit")


def load_reference_module(name):
    ref_dir = reference_dir()
    path = os.path.join(ref_dir, name + ".py")
    if not os.path.exists(path):
        raise ImportError("reference %s.py not found under %s (set MCD_REFERENCE_DIR)" % (name, ref_dir))
    if ref_dir not in sys.path:
        sys.path.append(ref_dir)          # its own imports (clip, data_utils) resolve as they do for the reference
    spec = importlib.util.spec_from_file_location("_reference_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
