"""prisim_b200 -- B200-native visibility engine behind PRISim's observe / delay-spectrum API.

Host code (this package) mirrors the reference's Python surface for the hot path only; all
numerical work runs in hand-written sm_100a CUDA kernels behind the C-ABI in
``include/prisim_b200.h`` (``libprisim_b200.so``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
