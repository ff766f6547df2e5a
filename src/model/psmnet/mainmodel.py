"""Model-directory contract of the reference (src/model/model_selector.py:11-15): defines class ``PSMNET``."""
from dualpixelface_b200.models import PSMNET  # noqa: F401
