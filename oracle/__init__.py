"""CPU oracle for the AGILE3D hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
CPU baseline.  Nothing under ``agile3d_b200/`` imports it.

Parity status: the decoder / positional-encoding / graph half is pinned against
the unmodified reference files (tests/golden/make_golden.py runs them on top of
``oracle.me_ref``).  The MinkowskiEngine half is a restatement of SURVEY.md
Appendix A because MinkowskiEngine is neither vendored in /root/reference nor
installable here: **parity unpinned at the MinkowskiEngine boundary**.
"""
