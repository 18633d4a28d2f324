"""deepmimic_b200 -- B200-native batched implementation of DeepMimic's per-step simulation hot path.

The compute path is hand-written sm_100a CUDA behind a C ABI (include/deepmimic_b200.h,
deepmimic_b200/libdeepmimic_b200.so).  This package holds only what that path needs on the host:
  capi.py    ctypes binding of the C ABI (torch tensors carry the device memory)
  env.py     mirror of the reference's Python env surface (R/env/deepmimic_env.py) over the C ABI
  assets.py  locating / unpacking the asset files
  build.py   in-tree build of the CUDA library and of the CPU oracle
There is no CPU fallback: importing works anywhere, running needs the built library and a CUDA device."""
from .assets import asset_root  # noqa: F401
