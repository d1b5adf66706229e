"""CPU oracle for the MVS photo-consistency hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs
of ``bench.py`` do, and there only as the checker / the timed CPU baseline.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 8c), so the restatements here are pinned against outputs of
the reference itself, produced by importing /root/reference in the build
container (``oracle/make_golden.py``) and committed under ``tests/golden/``.

Modules: ``mode_a`` (the scorer, NumPy), ``mode_a.c`` + ``c_port`` (the same scorer restated independently in
plain C, built by gcc into ``oracle/_build/``; pinned to the same golden vectors), ``cameras``, ``expansion``,
``filter``, ``triangulate``, ``ref_port`` (the cost-faithful CPU arm), ``mode_b`` (spec of the extension).
"""
