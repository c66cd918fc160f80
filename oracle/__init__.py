"""CPU oracle for the RVC-MAKER ``Synthesizer.infer`` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rvc_maker_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the timed CPU baseline.

Parity pin: the restatement in ``synth_oracle.py`` is checked against the
UNMODIFIED reference module (imported from ``/root/reference`` in the build
container by ``oracle/make_golden.py``) through the committed fixtures in
``tests/golden/`` (see ``tests/test_oracle_golden.py``).  The reference itself
ships no tests or golden vectors, so those fixtures are the pin.
"""
