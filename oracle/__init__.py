"""oracle -- CPU restatement of the reference's algorithms for the hot path.

TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import, build or
call anything in this package -- and only as the checker (or the timed CPU
baseline), never as part of the product path.  ``client_b200`` does not import
it; the CUDA path fails loudly when libtb200.so is missing instead of falling
back to this code.

Pinning status (see DESIGN.md "Oracle"):
  * wire encoders (``oracle.wire``): pinned by golden vectors generated from the
    reference Python client itself (``oracle/gen_golden.py`` ->
    ``tests/golden/wire_golden.json``) and by the reference's own known-answer
    tests.
  * image pack (``oracle.image``): literally the numpy arithmetic of
    src/python/examples/image_client.py:154-193; pinned by fixtures generated
    from that function.
  * random fill (``oracle.fill``): perf_analyzer's generator is not in the
    reference -> PARITY WITH perf_analyzer VALUES IS UNPINNED.  The Philox block
    function is pinned by the published Random123 known-answer vectors and by
    cuRAND's host generator.
"""
