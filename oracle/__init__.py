"""oracle/ -- TEST INFRASTRUCTURE, not product.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import anything from this package.  The product package
(`graph_colouring_with_rl_b200`) never imports it and fails loudly when its CUDA
extension is missing.
"""
