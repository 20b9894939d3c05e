"""Drop-in for the reference's model/Temporal_shift/cuda/shift.py (``from cuda.shift import Shift`` after
``sys.path.append("./model/Temporal_shift/")``).  This directory deliberately has no __init__.py: the image
pre-imports NVIDIA's ``cuda`` namespace package, and a namespace portion is the only way ``cuda.shift`` resolves
(SURVEY.md App. C-3)."""
from shiftgcn_b200.shift import Shift, ShiftFunction, shift_cuda  # noqa: F401
