"""CPU oracle for the CoT-block hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import, call, link or execute it -- and there only as the checker
(or as the timed CPU baseline), never as the implementation that is measured or
shipped.  The product path (``cotnet_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Parity status
-------------
* ``agg_ref``  : restates ``cupy_layers/aggregation_zeropad.py:20-110`` and
  ``cupy_layers/aggregation_zeropad_mix.py:20-207`` (kernel index math) and the
  Unfold identities of the reference self-tests (``aggregation_zeropad.py:249-251``,
  ``aggregation_zeropad_mix.py:360-366``).  PINNED: the two restatements are checked
  against each other to 1e-9 in fp64 at the self-test shapes, exactly the gate the
  reference's own tests apply (the reference stores no golden vectors).
* ``cot_ref``  : restates ``models/cotnet.py:36-104`` (CotLayer), ``:106-178``
  (CoXtLayer) and ``models/cotnet_hybrid.py:48-116`` (CoTLayer).  PINNED against
  outputs of the reference's own, unmodified module code imported in the build
  container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
"""
