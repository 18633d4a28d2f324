"""Stands where SWIG's generated DeepMimicCore.py stands in the reference (R/DeepMimicCore/Makefile:54-57): exposes cDeepMimicCore."""
import os

try:
    from ._DeepMimicCore import cDeepMimicCore  # noqa: F401
except ImportError as e:  # pragma: no cover
    raise ImportError("deepmimic_b200: the _DeepMimicCore extension is not built (run __graft_entry__.build()); "
                      "there is no CPU fallback. (%s)" % e)
