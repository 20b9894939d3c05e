from . import tools  # noqa: F401
from . import ntu_rgb_d  # noqa: F401
from . import mediapipe_pose  # noqa: F401
