"""Drop-in for the reference's compiled ``shift_cuda`` extension (model/Temporal_shift/cuda/shift_cuda.cpp:44-47,
built by cuda/setup.py and imported as a top-level module at cuda/shift.py:5): ``forward(input, xpos, ypos, stride)``
and ``backward(grad_output, input, output, xpos, ypos, stride)`` with the reference's argument order, return values
and RuntimeError behaviour, served by libshiftgcn_b200.so (``sgcn_shift_{fwd,bwd}_nchw_*``).  Resolves from the
repository root, like ``model.shift_gcn``; the reference's own unmodified cuda/shift.py runs on top of it."""
from shiftgcn_b200.shift import _native_backward as backward  # noqa: F401
from shiftgcn_b200.shift import _native_forward as forward  # noqa: F401
