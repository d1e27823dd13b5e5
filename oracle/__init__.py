"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy / scipy / torch-CPU) of the ARLib graph-CF hot path plus
a loader that imports the real reference from /root/reference when it is
mounted (builder container only).  Nothing under ``arlib_b200/`` may import
this package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker /
the timed CPU baseline -- never as the product path.

Parity status: the reference ships no tests or golden vectors,
so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the
builder container by ``oracle/make_golden.py`` and frozen under
``tests/golden/`` (see DESIGN.md "Oracle").
"""
