"""dualpixelface_b200 -- B200 (sm_100a) implementation of the DualPixelFace stereo hot path.

Cost-volume construction -> 3-D hourglass aggregation -> disparity regression (+ the ANM normal branch), behind the
reference's model-class contract.  Hand-written CUDA lives in ``csrc/`` and is reached through the C ABI declared in
``include/dpf_sm100.h`` (``_lib`` binds it with ctypes, ``ops`` wraps it for CUDA tensors).  No CPU fallback.
"""
__version__ = "0.1.0"
