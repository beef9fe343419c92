"""CPU oracle for the x-vector extraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
reported CPU baseline.  The product package
(``speaker-recognition-x-vectors_b200``) never imports this package and has
no CPU fallback.

Parity status: PINNED.  ``oracle/xvector_oracle.py`` is checked against
  * the reference's only known-answer fixtures (extra/time_context_test.py:3-39
    and the tdnn_layer.py:45-53 docstring example), and
  * outputs of the unmodified reference modules (``tdnn_layer.TdnnLayer``,
    ``main.XVectorModel``) imported in the build container by
    ``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``.
"""
