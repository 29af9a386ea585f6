"""CPU oracle for the psketch hot path — TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  See DESIGN.md §"Oracle" and the header of ``craft_oracle.c``.
"""
