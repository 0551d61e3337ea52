"""Test infrastructure only: a stand-in for the two Biopython calls the reference's scripts/generate_kmers.py makes, so that
the UNMODIFIED script can be run in the build container (Biopython is not installed) to produce golden vectors
(tests/golden/make_kmers_golden.py).  Never imported by the product."""
