"""cortenn_b200 — B200-native evaluator for Cortenn's LLO operators.

`cortenn_b200.age`  graph builders (the pybinder-generated `llo.age` surface)
`cortenn_b200.llo`  variable / evaluate / derive / seed / print_graph (`llo.llo`)
`cortenn_b200.capi` ctypes binding of the C-ABI library libllo_cuda.so

The native libraries are built in-tree by `cortenn_b200.build` (see
__graft_entry__.build()).  There is no CPU fallback: importing works without a
GPU (graphs can be built and derived), evaluating raises.
"""
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))

try:
    from . import _C  # noqa: F401
except ImportError as exc:  # pragma: no cover - build problem, fail loudly
    raise ImportError(
        "cortenn_b200: native extension missing or unloadable (%s). "
        "Run `python cortenn_b200/build.py` (needs nvcc + g++); there is no "
        "pure-python or CPU fallback." % (exc,)) from exc

from . import age, llo, capi  # noqa: E402,F401

__all__ = ["age", "llo", "capi"]
