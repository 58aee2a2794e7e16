"""TEST INFRASTRUCTURE ONLY: CPU oracle for the Hanabi hot path.

Nothing under ``hanabi_sad_b200/`` may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
"""
