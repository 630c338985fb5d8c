"""CPU oracle for the potential-based diffusion planning hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`potential-motion-plan-release_b200/`) may import this package.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl
reference legs use it, and only as the checker / reported CPU baseline.

Parity status: the reference repository contains no tests, golden vectors or
known-answer fixtures for this path, so the oracle is
pinned against outputs of the *reference code itself*, imported in the build
container through `oracle/refshim.py` and frozen as fixtures under
`tests/golden/` by `tests/golden/gen_golden.py`.  The Kuka collision model is
the exception: the reference delegates it to pybullet 3.2.5 (not vendored, not
installable here), so the sphere-link model in `oracle/colchk_ref.py` is
self-defined -> "parity unpinned" for that one function.  Pinned beyond the
sampler: Maze2D collision checks (colchk_maze2d.npz), the replanner
(replan_maze2d.npz), the multi-horizon candidate selection
(pick_multi_horizon.npz), the concave layout arithmetic (concave_layouts.npz).
"""
