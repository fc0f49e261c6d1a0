"""movi_b200 -- B200 (sm_100a) implementation of Movi's query hot path (PML / ZML / count).

Python is only the test / bench harness here: the product is the C ABI in include/movi_gpu.h
(CUDA kernels, movi_b200/csrc) plus the C++ host side (movi_b200/host).  `MoveStructure` mirrors the
reference class of the same name for the query path (include/move_structure.hpp:89,150-151).
"""
from .api import HostIndex, MoveStructure, Reads, MoviGpuError, device_count  # noqa: F401
