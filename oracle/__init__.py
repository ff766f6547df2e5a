"""CPU oracle for the DualPixelFace stereo hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or the timed
CPU baseline) -- never as the thing shipped.  The product path
(``dualpixelface_b200``) must not import this package and fails loudly when
its CUDA library is missing.

Parity status: the reference ships no golden vectors or known-answer tests
(SURVEY.md section 4).  This oracle is pinned instead against outputs of the
reference's own code imported from ``/root/reference`` through the shim layer
in ``tests/golden/ref_shim.py``; the fixtures and the generating script are
committed under ``tests/golden/``.  The one piece that cannot be executed in
the build container is the reference's compiled 3-D deformable convolution
(CUDA only, no CPU path): the oracle restates
``src/module/dcn3d/src/cuda/deform_im2col_cuda.cuh``; that restatement is pinned
on the GPU box against the reference's own CUDA kernels, compiled from the
unmodified sources into the git-ignored ``oracle/_ref/DCN.so`` by
``oracle/build_ref_dcn.py`` (``tests/test_gpu_dcn_reference.py``: forward and all
four gradients within 1e-4).
"""
