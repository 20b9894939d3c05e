from . import shift_gcn  # noqa: F401
