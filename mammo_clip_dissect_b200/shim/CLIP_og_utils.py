"""Drop-in `CLIP_og_utils` module: re-exports the reference's own module (found through
MCD_REFERENCE_DIR, default /root/reference/concept_vit) with the two functions on the hot path
swapped for the B200 implementations:

    get_activation                   -> mammo_clip_dissect_b200.hooks.get_activation        (K4 pooling)
    get_similarity_from_activations  -> mammo_clip_dissect_b200.features.get_similarity_from_activations (K1 + scoring)

Everything else (save_activations, get_save_names, model / data plumbing) stays the reference's code.
The reference registers hooks through eval("target_model.<layer>.register_forward_hook(get_activation(...))")
inside its own module namespace, so the swap is made there as well.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

_REF_DIR = os.environ.get("MCD_REFERENCE_DIR", "/root/reference/concept_vit")
_REF_FILE = os.path.join(_REF_DIR, "CLIP_og_utils.py")
if not os.path.exists(_REF_FILE):
    raise ImportError("reference CLIP_og_utils.py not found under %s (set MCD_REFERENCE_DIR)" % _REF_DIR)
if _REF_DIR not in sys.path:
    sys.path.append(_REF_DIR)          # its own imports (clip, data_utils) resolve as they do for the reference

_spec = importlib.util.spec_from_file_location("_reference_CLIP_og_utils", _REF_FILE)
_ref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_ref)

from mammo_clip_dissect_b200.features import get_similarity_from_activations  # noqa: E402
from mammo_clip_dissect_b200.hooks import get_activation  # noqa: E402

_ref.get_activation = get_activation
_ref.get_similarity_from_activations = get_similarity_from_activations
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
