"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithms for the sampling hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (difffacto_b200/) never does.
"""
