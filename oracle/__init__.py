"""TEST INFRASTRUCTURE: CPU checkers of the BMFR hot path (see oracle/bmfr_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
