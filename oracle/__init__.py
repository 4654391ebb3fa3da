"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the Skylarking/MARL hot path (learner step + matrix game).
Nothing under ``marl_b200/`` may import this package; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do, and only as the checker / the timed CPU baseline.
"""
