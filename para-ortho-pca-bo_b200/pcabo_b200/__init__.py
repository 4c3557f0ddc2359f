"""pcabo_b200 -- host-side binding of the B200-native PCA-BO hot path (libpcabo_b200.so, C ABI in include/pcabo_b200.h).

PyTorch is used only for device memory, streams and torch.distributed; every numerical step of the hot path runs in
the hand-written sm_100a kernels under ../csrc. There is no CPU fallback: calling an operator without the compiled
library (or without a CUDA device) raises.
"""
from .lib import load_library, library_path, PcaboError, EXPORTED_SYMBOLS  # noqa: F401
