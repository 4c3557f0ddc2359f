"""pcabo_b200 -- host-side binding of the B200-native PCA-BO hot path (libpcabo_b200.so, C ABI in include/pcabo_b200.h).

PyTorch is used only for device memory, streams and torch.distributed; every numerical step of the hot path runs in
the hand-written sm_100a kernels under ../csrc. There is no CPU fallback: calling an operator without the compiled
library (or without a CUDA device) raises.
"""
import os as _os

# PipelinedEngines drives every sub-batch on two CUDA streams; with the default of 8 hardware work queues streams beyond
# the eighth share a queue and their kernels serialise (measured: 9 sub-batches ran one optimiser kernel after the other).
# Read by the driver when the CUDA context is created, so it only helps when this import precedes the first CUDA call.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .lib import load_library, library_path, PcaboError, EXPORTED_SYMBOLS  # noqa: F401
