"""pmp_b200: ctypes face of libpmp_b200.so (hand-written sm_100a CUDA kernels behind the C ABI
declared in include/pmp_b200.h).  PyTorch is used only for device memory and streams."""
from ._lib import (lib, PmpError, Arch, UnetEngine, sample, step, init_noise, unnormalize,
                   colchk_maze2d_swept, colchk_maze2d_points, colchk_kuka, pick_first_valid,
                   launch_count, device_count, profile_enable, profile_read, set_option, LIB_PATH)

__all__ = ["lib", "PmpError", "Arch", "UnetEngine", "sample", "step", "init_noise", "unnormalize",
           "colchk_maze2d_swept", "colchk_maze2d_points", "colchk_kuka", "pick_first_valid",
           "launch_count", "device_count", "profile_enable", "profile_read", "set_option", "LIB_PATH"]
