"""Model-directory contract of the reference (README "model" section; src/model/model_selector.py:11-15):
``src/model/<name>/mainmodel.py`` defines class ``<NAME upper>``, constructed as ``CLS(option)``.
The class itself lives in dualpixelface_b200.models and runs the hot path on the sm_100a kernels."""
from dualpixelface_b200.models import STEREODPNET  # noqa: F401
