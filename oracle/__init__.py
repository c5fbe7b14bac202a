"""TEST INFRASTRUCTURE (CPU oracle).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; reinlife_b200/ never does."""
