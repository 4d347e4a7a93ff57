"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the IIF classifier head.

Nothing under ``iif_b200/`` may import this package.  The only permitted
importers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` (as the checker / the timed CPU arm,
never as the product path).

Parity status: PINNED.  ``head_oracle`` is checked by ``tests/test_oracle_*.py``
against (i) outputs of the reference's own Python (``classification/custom.py``,
``mmdet/models/losses/{iif_loss,fasa_iif_loss,cross_entropy_loss,accuracy}.py``)
imported unmodified in the build container and frozen under ``tests/golden/``
by ``tests/golden/make_golden.py``, (ii) the reference's CSV weight tables
(``lvis_files/idf_1204.csv``, ``idf_1231.csv``, ``coco_files/idf_91.csv``) and
(iii) the known-answer tests of the reference's mmdet test-suite
(``tests/test_metrics/test_losses.py:8-32,186-240``).
"""
