"""TEST INFRASTRUCTURE ONLY — CPU oracle for the ChAda-ViT / DINO hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / the reported CPU baseline.
The product path (``chadavit_b200``) never imports this package and fails loudly
when its CUDA library is missing.

Parity status: PINNED — ``oracle.chada_oracle`` is checked (tests/test_oracle_golden.py)
against outputs of the reference's own modules run in the build container
(fixtures in ``tests/golden/*.npz`` made by ``tests/golden/make_golden.py`` which
imports /root/reference through ``oracle.ref_loader``).
"""
