"""Drop-in `og_utils` module: re-exports the reference's own module (found through MCD_REFERENCE_DIR, default
/root/reference/concept_vit) with the two functions on the hot path swapped for the B200 implementations:

    get_activation                   -> mammo_clip_dissect_b200.hooks.get_activation        (K4 pooling)
    get_similarity_from_activations  -> mammo_clip_dissect_b200.features.get_similarity_from_activations (K1 + scoring;
                                        same keyword surface as reference og_utils.py:478-521: d_probe)

Everything else (save_activations, get_save_names, model / data plumbing) stays the reference's code.  The reference
registers hooks through eval("...register_forward_hook(get_activation(...))") inside its own module namespace
(og_utils.py:87), so the swap is made there as well.
"""
import functools

from _reference_module import load_reference_module

from mammo_clip_dissect_b200.features import get_similarity_from_activations as _gsfa
from mammo_clip_dissect_b200.hooks import get_activation

_ref = load_reference_module("og_utils")
get_similarity_from_activations = functools.partial(_gsfa, target_on_device=True)
functools.update_wrapper(get_similarity_from_activations, _gsfa)
_ref.get_activation = get_activation
_ref.get_similarity_from_activations = get_similarity_from_activations
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
