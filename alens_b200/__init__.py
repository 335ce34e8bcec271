"""alens_b200 -- B200-native (sm_100a) collision-constraint path of aLENS / SimToolbox.

The product is the C-ABI shared library ``libalens_b200.so`` (sources in ``alens_b200/csrc``,
interface in ``include/alens_b200.h``) plus the C++ drop-in headers in ``include/alens_b200/``.
This Python package is only the ctypes harness that tests and bench.py drive the library with.
It never imports anything from ``oracle/`` and has no CPU fallback.
"""
from .capi import Library, Context, Bcqp, AlensError, BLOCK_DTYPE, lib_path, build, comm_connect_local  # noqa: F401

__all__ = ["Library", "Context", "Bcqp", "AlensError", "BLOCK_DTYPE", "lib_path", "build"]
