"""Model-directory contract of the reference (src/model/model_selector.py:11-15): defines class ``NNET``."""
from dualpixelface_b200.nnet import NNET  # noqa: F401
