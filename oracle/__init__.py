"""CPU oracle for the MVS photo-consistency hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs
of ``bench.py`` do, and there only as the checker / the timed CPU baseline.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 8c), so the restatements here are pinned against outputs of
the reference itself, produced by importing /root/reference in the build
container (``oracle/make_golden.py``) and committed under ``tests/golden/``.
"""
