"""TEST INFRASTRUCTURE ONLY - CPU restatements of the reference's hot-path algorithms.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker or the CPU baseline - never on the product path.
"""
