"""rtlsdr_b200 -- B200 (sm_100a) implementation of rtl_power's per-hop scan pipeline.

The product is the C-ABI library ``librtlsdr_gpu_scan.so`` declared in
``include/rtlsdr_gpu_scan.h``; this package only builds it (``_build``) and
binds it for Python callers (``scan``).  There is no CPU fallback: without the
compiled CUDA library every entry point raises.
"""
from .scan import GpuScan, ScanError, lib_path, load_library  # noqa: F401
from .planner import Plan, plan_scan  # noqa: F401

__all__ = ["GpuScan", "ScanError", "Plan", "plan_scan", "lib_path", "load_library"]
