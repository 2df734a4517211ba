"""CPU oracle for the embedding-vs-gallery matching path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker or the timed CPU baseline - never as a
fallback for the CUDA path (the product raises when its CUDA library is missing).

Parity status: the reference (bharatlytics/faceRecognition_InfrenceEngine) ships no
tests, fixtures or golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed verbatim in the
build container under module stubs (``oracle/ref_harness.py``) and frozen as
``tests/golden/*.npz`` by ``oracle/make_golden.py``.  Behaviour the reference does
not have (top-k > 1, Euclidean metric, bf16 gallery) is OUR definition and is
labelled "parity unpinned" where it is restated.

Modules
-------
synth            counter-based synthetic embeddings, bit-identical to the device generator
matcher_oracle   numpy restatement of the reference's normalise / scan / decide / store logic
ref_harness      runs the reference's own code under stubs (needs /root/reference; build box only)
make_golden      regenerates tests/golden/ from ref_harness
cpu_baseline     times the restated per-face Python loop on the host cores
"""
