"""custom_diffusion360_b200 — B200-native (sm_100a) implementation of the pose-conditioned SDXL
UNet denoising step of customdiffusion360/custom-diffusion360.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (include/cd360.h) -> libcd360.so
  _lib.py    ctypes binding;  ops.py  torch-tensor front end of the ABI
  sgm/       host-side mirror of the reference's sgm module surfaces (UNetModel, SpatialTransformer,
             BasicTransformerBlock, NerfSDModule, DiscreteDenoiser, EulerEDMSampler, guiders, ...)
"""
__version__ = "0.1.0"
