"""TEST INFRASTRUCTURE -- CPU oracle for the pyslam Gauss-Newton hot path.

Nothing under oracle/ is part of the product.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg
may import it, and there only as the checker / the reported CPU baseline.
"""
