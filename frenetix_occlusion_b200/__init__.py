"""frenetix_occlusion_b200 -- B200-native per-planning-step occlusion assessment hot path.

Drop-in for the path behind the reference's ``frenetix_occlusion/interface.py`` (FOInterface,
the pluggable ``metrics/`` modules and the ``occlusion.yaml`` thresholds).  All arithmetic runs in
hand-written sm_100a CUDA kernels behind the C-ABI declared in ``include/fo_b200.h``; there is no
CPU fallback: importing the compute modules without the built library raises.
"""
__version__ = "0.1.0"
