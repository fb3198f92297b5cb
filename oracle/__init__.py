"""CPU oracle of the ParallelFDTD hot path.  TEST INFRASTRUCTURE ONLY: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / reference legs -- never from the product."""
