"""CPU oracle for the scattering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``kymatio_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker or the timed CPU baseline - never as the product path.

Every function cites the reference file:line it restates (paths relative to
the kymatio reference tree).  Parity pinning: ``tests/golden/make_golden.py``
ran the unmodified reference (numpy frontend) in the build container and
committed its outputs under ``tests/golden/``; ``tests/test_oracle_*.py``
check this oracle against those vectors and against the reference's own
``test_data_{1,2,3}d.npz`` fixtures (re-saved verbatim as
``tests/golden/ref_fixture_*.npz``).
"""
