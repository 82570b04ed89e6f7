"""Drop-in `similarity` module: put this directory first on sys.path and the reference drivers'
`import similarity` / eval("similarity.<name>") (describe_clip_neurons.py:9,41) resolve to the
B200 kernels.  Same names and signatures as reference concept_vit/similarity.py."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

from mammo_clip_dissect_b200.similarity import (  # noqa: E402,F401
    cos_similarity, cos_similarity_cubed, cos_similarity_cubed_single, rank_reorder, soft_wpmi, wpmi)
