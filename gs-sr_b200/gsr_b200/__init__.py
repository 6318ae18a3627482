"""gsr_b200 -- loader and thin ctypes binding of libgsr_b200.so (include/gsr_b200.h).

The drop-in Python packages next to this one (``diff_surfel_rasterization`` ...)
mirror the reference extensions' public API and call through here.  There is no
CPU or PyTorch fallback: if the CUDA library is missing or a call fails, an
exception is raised.
"""
from ._lib import lib, check, BUFFER_FN, TorchBuffers, ptr, LIB_PATH  # noqa: F401
