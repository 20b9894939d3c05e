"""``shift_cuda`` next to ``shift.py``, for callers that put this directory itself on sys.path (the reference builds
the extension here, model/Temporal_shift/cuda/setup.py:4-14).  Same two functions as the top-level ``shift_cuda``."""
from shiftgcn_b200.shift import _native_backward as backward  # noqa: F401
from shiftgcn_b200.shift import _native_forward as forward  # noqa: F401
