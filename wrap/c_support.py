"""``import wrap.c_support`` — the import path the reference's trainers use (dim_red/triplet.py:143,
dim_red/angular.py:180) for its SWIG module (wrap/c_support.i, built per wrap/README.md:6-8 into
``wrap/_c_support.so`` + ``wrap/c_support.py``, a namespace package without ``__init__.py``).

With this repository's root on ``sys.path`` (or this directory copied next to the trainer) the unedited
trainer resolves the same name to the B200 implementation: everything is forwarded to
``gbnns_dim_red_b200.wrap.c_support`` (same positional signature, returns 0, tolerant of the angular
trainer's 9th argument; the measured accuracies are in ``last_results()``)."""
from gbnns_dim_red_b200.wrap.c_support import get_graphs_and_search_tests, last_results, search_tests  # noqa: F401
