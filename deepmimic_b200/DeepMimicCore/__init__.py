"""Drop-in for the reference's SWIG package: `from DeepMimicCore import DeepMimicCore` then
`DeepMimicCore.cDeepMimicCore(enable_draw)` (R/env/deepmimic_env.py:3,10).  Put the directory that contains this package
(deepmimic_b200/) on sys.path in place of the reference's DeepMimicCore/ build directory."""
