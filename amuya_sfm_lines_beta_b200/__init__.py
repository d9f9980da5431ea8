"""B200-native line bundle adjustment: the hot path of bgpantojar/amuya_sfm_lines_beta
(``src/bundle_adjustment.py``) behind the reference's own call signature.

    from amuya_sfm_lines_beta_b200.bundle_adjustment import run_bundle_adjustment   # drop-in
    from amuya_sfm_lines_beta_b200 import LineBA                                      # array-native

The compute lives in ``liblba.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/lba.h``).  Importing the package does not need a GPU; creating a ``LineBA``
does, and fails loudly without one.
"""
from .solver import LineBA, LMResult  # noqa: F401
from . import scenes, sharding  # noqa: F401

__all__ = ["LineBA", "LMResult", "scenes", "sharding"]
