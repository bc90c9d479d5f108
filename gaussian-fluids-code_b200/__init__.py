"""
gaussian-fluids-code_b200 — B200-native engine for the hot path of Gaussian Fluids (DvvCz/Gaussian-Fluids-Code):
evaluating the Gaussian Spatial Representation and running its per-timestep optimisation.

  gsr3d / gsr2d      drop-in classes for the reference's 3D/GSR.py and 2D/GSR.py
  engine, _lib       torch-tensor -> raw-pointer plumbing over the C ABI (include/gsr_b200.h)
  csrc/              hand-written sm_100a CUDA kernels + the extern "C" entry points
  build              nvcc build of csrc/ into csrc/libgsr_b200.so

Import as `gaussian_fluids_code_b200` (see the shim module of that name at the repo root).
"""
__version__ = '0.2.0'
