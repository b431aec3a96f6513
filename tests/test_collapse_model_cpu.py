"""The cost-optimal collapse of the binary hierarchy into 8-wide nodes (collapse_dp_kernel / widen_dp_kernel in csrc/bvh_build.cu,
VHR_COLLAPSE=1) restated in Python (tests/collapse_dp_model.py, same recurrences, same decision bytes, same gather loop) and checked
against an exhaustive search on small random trees: the dynamic programme finds the optimum, the children read off the decisions
cover every triangle once, and the emitted tree costs what the table says."""
import random

import collapse_dp_model as M


def test_dp_collapse_is_optimal_and_consistent():
    rng = random.Random(7)
    for _ in range(200):
        n = rng.randrange(2, 12)
        cn, ct = rng.choice([0.5, 1.0, 2.0]), rng.choice([0.7, 1.0, 1.5])
        root = M.rand_tree(n, rng)
        M.dp(root, cn, ct)
        seen = []
        total = M.emit_cost(root, cn, ct, seen)
        assert sum(seen) == n
        assert abs(total - root.cost[0]) <= 1e-9 * max(1.0, total)
        best = M.brute(root, cn, ct)
        assert abs(best - root.cost[0]) <= 1e-9 * max(1.0, best)


def test_costs_never_increase_with_more_roots():
    rng = random.Random(11)
    root = M.rand_tree(40, rng)
    M.dp(root, 1.0, 1.0)
    stack = [root]
    while stack:
        x = stack.pop()
        if x.l is None:
            continue
        assert all(a >= b for a, b in zip(x.cost, x.cost[1:]))
        assert all(0 <= d <= 7 for d in x.dec)
        stack += [x.l, x.r]
