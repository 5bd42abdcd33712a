"""nlc_b200 — B200-native (sm_100a) implementation of Diffusion-NLC's per-timestep sampling loop.

Host side: Python mirrors of the reference interfaces (same names, arguments and state_dict keys);
device side: hand-written CUDA kernels behind the C ABI of include/nlc_b200.h (libnlc_b200.so).
"""
__version__ = "0.1.0"
