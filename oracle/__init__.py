"""CPU oracle of the X2I hot path.  TEST INFRASTRUCTURE ONLY (see each module's header):
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never from the product package x2i_b200/."""
