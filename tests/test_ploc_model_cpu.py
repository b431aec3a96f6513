"""The PLOC hierarchy builder (ploc_nn_kernel / ploc_flag_kernel / ploc_merge_kernel in csrc/bvh_build.cu) restated in numpy, to pin the
two properties the CUDA code relies on: (1) with ties broken towards the smaller index, every round has at least one mutual
nearest-neighbour pair, so the loop always makes progress; (2) the scan-based numbering — round offset n - N plus the rank of the
absorbed partner among the absorbed clusters — hands out every internal id 0 .. n-2 exactly once, the last merge gets id 0 (the
root), and the result is one binary tree over all leaves, identical on every run (no atomics)."""
import numpy as np
import pytest


def _half_area(mn, mx):
    d = mx - mn
    return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]


def ploc(mn, mx, radius):
    n = len(mn)
    ids = np.arange(n - 1, 2 * n - 1)                  # unified ids: leaves are n-1 .. 2n-2
    child_l, child_r = np.full(n - 1, -1), np.full(n - 1, -1)
    parent = np.full(2 * n - 1, -1)
    N, rounds = n, 0
    while N > 1:
        rounds += 1
        nn = np.empty(N, int)
        for i in range(N):                             # ploc_nn_kernel: ascending scan, strict <, so ties keep the smaller index
            best, bj = np.inf, -1
            for j in range(max(0, i - radius), min(N, i + radius + 1)):
                if j == i:
                    continue
                ar = min(_half_area(np.minimum(mn[i], mn[j]), np.maximum(mx[i], mx[j])), 3.0e38)
                if ar < best:
                    best, bj = ar, j
            nn[i] = bj
        mutual = nn[nn] == np.arange(N)
        valid = ~(mutual & (np.arange(N) > nn))        # ploc_flag_kernel
        pos = np.concatenate([[0], np.cumsum(valid)])  # exclusive scan, pos[N] = next N
        assert pos[N] < N, "a round merged nothing"
        new_ids, new_mn, new_mx = np.empty(pos[N], int), np.empty((pos[N], 3)), np.empty((pos[N], 3))
        for i in range(N):                             # ploc_merge_kernel
            if not valid[i]:
                continue
            o, j = pos[i], nn[i]
            if nn[j] != i:
                new_ids[o], new_mn[o], new_mx[o] = ids[i], mn[i], mx[i]
                continue
            nid = n - 2 - ((n - N) + (j - pos[j]))
            assert child_l[nid] == -1, "internal id handed out twice"
            child_l[nid], child_r[nid] = ids[i], ids[j]
            parent[ids[i]] = parent[ids[j]] = nid
            new_ids[o], new_mn[o], new_mx[o] = nid, np.minimum(mn[i], mn[j]), np.maximum(mx[i], mx[j])
        ids, mn, mx, N = new_ids, new_mn, new_mx, pos[N]
    return child_l, child_r, parent, ids[0], rounds


@pytest.mark.parametrize("n,radius,seed", [(2, 8, 0), (3, 8, 1), (17, 4, 2), (64, 8, 3), (200, 8, 4), (200, 2, 5)])
def test_ploc_model_builds_one_tree_with_scan_numbering(n, radius, seed):
    rng = np.random.default_rng(seed)
    c = np.sort(rng.uniform(0, 10, n))[:, None] * np.array([1.0, 0.3, 0.1]) + rng.uniform(0, 0.5, (n, 3))
    mn, mx = c, c + rng.uniform(0.01, 0.3, (n, 3))
    if seed == 4:                                      # many exact ties: identical boxes
        mn[:] = mn[0]; mx[:] = mx[0]
    child_l, child_r, parent, root, rounds = ploc(mn.copy(), mx.copy(), radius)
    assert root == 0 and parent[0] == -1
    assert np.all(child_l >= 0) and np.all(child_r >= 0)
    # every node except the root has exactly one parent; every leaf is reachable from the root
    kids = np.concatenate([child_l, child_r])
    assert len(np.unique(kids)) == 2 * n - 2 and 0 not in kids
    seen, stack = 0, [0]
    while stack:
        x = stack.pop()
        if x >= n - 1:
            seen += 1
        else:
            stack += [child_l[x], child_r[x]]
    assert seen == n
    # deterministic: a second run gives the same arrays
    again = ploc(mn.copy(), mx.copy(), radius)
    assert np.array_equal(again[0], child_l) and np.array_equal(again[1], child_r)
    assert rounds <= n
