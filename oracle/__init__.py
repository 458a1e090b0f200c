"""CPU oracle for the microImageLib hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and only as the checker.
The product (``microimagelib_b200``) never imports this package and has no CPU fallback.

Parity status: **pinned to the reference itself** (round 2).  The reference ships no tests, golden vectors or data
 and does not compile with CUDA 12 as it stands (legacy texture references in
``include/cukernel.cuh:29-35``; its CPU path needs FFTW, absent here), so

  * ``oracle/build_ref_gpu.py`` builds the reference's OWN libapi -- all 23 entry points, its real GPU path on cuFFT -- from
    the sources where they lie under ``/root/reference``, with a purely mechanical texture-object patch, into the
    git-ignored ``oracle/_ref/libapi_ref.so`` (``oracle/ref_gpu.py`` binds it);
  * ``tests/test_gpu_reference_pinned.py`` runs reference / oracle / product side by side on the GPU at the north-star
    tolerances; ``tests/golden/reference_vectors.npz`` (written from that library on a B200 by
    ``tests/golden/make_reference_golden.py``) pins the oracle on the CPU (``tests/test_reference_golden.py``);
  * the plain-C helpers (snapTransformSize, p2matrix, matrix2p, matrixmultiply, dof9tomatrix, checkmatrix) are compared bit
    for bit with the reference's compiled functions on the CPU (``tests/test_reference_host_helpers.py``);
  * ``src/api_powell.c`` is compiled unchanged into ``oracle/_ref/libpowell_ref.so`` (``oracle/Makefile``) and the optimiser
    restatement is checked against it bit for bit, with golden trajectories committed under ``tests/golden/``;
  * the texture-filter restatement in ``reg_oracle.c`` was fitted to the reference's own tex3D output (129 024 samples:
    99.9 % bit-identical, the rest one ulp).
The round-1 known-answer tests (delta image, flux conservation, identity transform ...) remain as drift guards.
"""
