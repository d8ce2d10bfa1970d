"""oracle -- TEST INFRASTRUCTURE ONLY: CPU restatements of the reference's algorithms for the hot path (each function cites the
reference file:line it follows), the scripts that captured tests/golden/*.npz from the unmodified reference, and the Makefile
that compiles the reference's own accessmath_lib.c into oracle/_ref/.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product (lecturemath_b200/) never does."""
