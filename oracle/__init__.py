"""CPU oracle for the microImageLib hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and only as the checker.
The product (``microimagelib_b200``) never imports this package and has no CPU fallback.

Parity status: the reference ships no tests, golden vectors or data, and its
own GPU/CPU sources cannot be built in this image (CUDA-12 rejects the legacy texture references in
``include/cukernel.cuh:29-35``; FFTW is absent), so the numerical restatements here are
**parity unpinned** by the reference itself.  They are pinned instead by
  * mathematics (un-normalised DFT checked against scipy/pocketfft in float64),
  * known-answer tests minted in ``tests/`` (delta image, flux conservation, identity transform...),
  * the one reference file that does build: ``src/api_powell.c`` is compiled unchanged into
    ``oracle/_ref/libpowell_ref.so`` (see ``oracle/Makefile``) and the optimiser restatement is checked
    against it bit-for-bit, with golden trajectories committed under ``tests/golden/``.
"""
