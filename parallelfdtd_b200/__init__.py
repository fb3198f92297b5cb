"""parallelfdtd_b200 -- B200-native time-stepping hot path of ParallelFDTD behind a C ABI.

`capi` is the ctypes door to libpfdtd_b200.so (CUDA kernels + C ABI, include/pfdtd.h);
`synth` generates voxelizer-style node volumes; `slabs` is the one-process-per-GPU z-slab driver.
There is no CPU implementation in this package: compute calls fail without a CUDA device.
"""
__version__ = "0.1"
