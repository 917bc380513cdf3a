"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the PRISim hot path.

Nothing in ``prisim_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker / the timed CPU baseline.
"""
