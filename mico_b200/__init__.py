"""mico_b200 -- B200-native (sm_100a) implementation of the MiCo omni-modal transformer hot path.

The package holds only what the path needs: `csrc/` (hand-written CUDA kernels + the C-ABI in
include/mico_b200.h), `ops.py` (ctypes call wrappers) and the host-side mirror of the reference's
nn.Module surface (`mico.py`, `eva_vit.py`, `bert.py`, ...).  There is no CPU fallback.
"""
__version__ = "0.1.0"
