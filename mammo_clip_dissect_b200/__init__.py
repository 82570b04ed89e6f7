"""mammo_clip_dissect_b200 -- the neuron->concept scoring path of Mammo-CLIP Dissect as hand-written
sm_100a CUDA kernels behind a C ABI (include/mcd_b200.h), with the reference's Python call surface.

    from mammo_clip_dissect_b200 import similarity          # soft_wpmi, wpmi, cos_similarity(_cubed)
    from mammo_clip_dissect_b200.hooks import get_activation
    from mammo_clip_dissect_b200.features import similarity_matrix, get_similarity_from_activations

Importing the package does not load the CUDA library; the first op does, and raises if
libmcd_b200.so has not been built (python -m mammo_clip_dissect_b200.build).
"""
__version__ = "0.1.0"
