"""Model-directory contract of the reference (src/model/model_selector.py:11-15): defines class ``STEREONET``."""
from dualpixelface_b200.stereonet import STEREONET  # noqa: F401
