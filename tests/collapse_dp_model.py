"""Python model of collapse_dp_kernel / widen_dp_kernel (bvh_build.cu) + brute force, small random trees."""
import random, itertools, functools, math
INF = 3.0e38
MAXLEAF = 3

class Node:
    def __init__(s, l=None, r=None, area=0.0, cnt=1):
        s.l, s.r, s.area, s.cnt = l, r, area, cnt
        s.cost = None; s.dec = None

def rand_tree(nleaves, rng):
    nodes = [Node(area=rng.uniform(0.1, 1.0)) for _ in range(nleaves)]
    while len(nodes) > 1:
        i = rng.randrange(len(nodes) - 1)
        a, b = nodes[i], nodes[i + 1]
        p = Node(a, b, area=max(a.area, b.area) + rng.uniform(0.0, 1.0) * (a.area + b.area), cnt=a.cnt + b.cnt)
        nodes[i:i + 2] = [p]
    return nodes[0]

def C(x, i, ct):            # dp_cost_of
    return x.area * ct if x.l is None else x.cost[i - 1]

def dp(x, cn, ct):          # collapse_dp_kernel, post-order
    if x.l is None: return
    dp(x.l, cn, ct); dp(x.r, cn, ct)
    D = [None] * 9; K = [None] * 9
    for j in range(2, 9):
        best, bk = INF, 1
        for k in range(1, j):
            if k > 7 or j - k > 7: continue
            c = C(x.l, k, ct) + C(x.r, j - k, ct)
            if c < best: best, bk = c, k
        D[j], K[j] = best, bk
    c_int = x.area * cn + D[8]
    c_leaf = x.area * x.cnt * ct if x.cnt <= MAXLEAF else INF
    leaf = c_leaf <= c_int
    x.cost = [c_leaf if leaf else c_int]; x.dec = [0 if leaf else K[8]]
    x.leaf = leaf
    prev = x.cost[0]
    for i in range(2, 8):
        dd = 0
        if D[i] < prev: prev, dd = D[i], K[i]
        x.cost.append(prev); x.dec.append(dd)

def gather(x):              # widen_dp_kernel
    kids = []
    st = [(x.r, 8 - x.dec[0]), (x.l, x.dec[0])]
    while st and len(kids) < 8:
        n, share = st.pop()
        if n.l is None or share <= 1: kids.append(n); continue
        k = n.dec[share - 1]
        if k == 0: st.append((n, share - 1)); continue
        st.append((n.r, share - k)); st.append((n.l, k))
    assert not st, "stack not drained"
    return kids

def emit_cost(x, cn, ct, leaves_seen):      # total SAH of the emitted wide tree
    if x.l is None or x.leaf:
        leaves_seen.append(x.cnt)
        return x.area * x.cnt * ct
    kids = gather(x)
    assert 2 <= len(kids) <= 8
    return x.area * cn + sum(emit_cost(k, cn, ct, leaves_seen) for k in kids)

def cuts(x, budget):                       # all ways to represent x by <= budget roots
    out = [[x]]
    if x.l is not None and budget >= 2:
        for b in range(1, budget):
            for a in cuts(x.l, b):
                for c in cuts(x.r, budget - b):
                    out.append(a + c)
    return out

@functools.lru_cache(maxsize=None)
def brute(x, cn, ct):
    if x.l is None: return x.area * ct
    best = x.area * x.cnt * ct if x.cnt <= MAXLEAF else INF
    for cut in cuts(x, 8):
        if len(cut) < 2: continue
        c = x.area * cn + sum(brute(k, cn, ct) for k in cut)
        best = min(best, c)
    return best

if __name__ == "__main__":
    rng = random.Random(1)
    for trial in range(300):
        n = rng.randrange(2, 13)
        cn, ct = rng.choice([0.5, 1.0, 2.0]), rng.choice([0.7, 1.0, 1.5])
        root = rand_tree(n, rng)
        dp(root, cn, ct)
        seen = []
        tot = emit_cost(root, cn, ct, seen)
        assert sum(seen) == n, (sum(seen), n)
        assert abs(tot - root.cost[0]) <= 1e-9 * max(1, tot), (tot, root.cost[0])
        b = brute(root, cn, ct)
        assert abs(b - root.cost[0]) <= 1e-9 * max(1, b), (b, root.cost[0], n)
    print("ok")
