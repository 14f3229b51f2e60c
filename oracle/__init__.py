"""CPU oracle for the Activation1d path -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Parity is pinned on tests/golden/ (outputs of the unmodified reference produced by
tests/golden/make_golden.py), because the reference ships no tests of its own for this path.
"""
