"""CPU oracle for the Paella hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package; ``paella_b200`` never does.
"""
