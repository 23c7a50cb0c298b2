"""se_snmf_nat_b200 -- B200-native (sm_100a) implementation of the SE_SNMF_NAT enhancement hot path.

The product is the C-ABI CUDA library ``libsnmfnat.so`` (``include/snmfnat.h``); this package is the
host-side mirror of the reference's MATLAB interface on top of it (``api``), the settings-script reader
(``settings``) and the in-tree build (``build``).  There is no CPU implementation in here.
"""
from . import _lib  # noqa: F401
from .api import (Batch, Context, SnmfnatError, default_p, enhance_batch, filewise_run_IS16,  # noqa: F401
                  get_context, params_struct, sqrt_hann_periodic)

__all__ = ["Batch", "Context", "SnmfnatError", "default_p", "enhance_batch", "filewise_run_IS16", "get_context",
           "params_struct", "sqrt_hann_periodic"]
