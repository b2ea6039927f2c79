"""CPU oracle for the beam-optimisation hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm may
import anything from this package; the product package ``openpystruct_b200`` never does.
PARITY UNPINNED: the reference has no tests and its FE engine (OpenSeesPy) is not installable
here -- see oracle/opensees_shim.py.
"""
