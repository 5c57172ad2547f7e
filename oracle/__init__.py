"""TEST INFRASTRUCTURE ONLY -- never imported by the product package ``stenos_b200``.

Two CPU checkers for the level-1 hot path of Thermadiag/stenos:

* ``oracle.port``  -- ctypes binding of ``oracle/stenos_oracle.c``, a plain-C restatement of the
  reference algorithm (each function cites the reference file:line it follows).
* ``oracle.ref``   -- ctypes binding of ``oracle/_ref/libstenos_ref.so``, the UNTOUCHED reference
  compiled from ``/root/reference`` by ``oracle/Makefile`` (``make ref``).  Used to pin the port and
  as the ``cpu_baseline`` / ``--impl reference`` arm of ``bench.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
