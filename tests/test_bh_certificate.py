"""The certificate behind the grouped Barnes-Hut walk's FP32 acceptance tests (nbody_b200/csrc/nb200_bh_group.cuh,
DESIGN.md 3.4), checked on the CPU with numpy emulating the kernel's arithmetic operation by operation:

    q   = fl32(c - O)                       node centre relative to the group's first target (subtraction in FP64)
    p   = fl32(x - O)                       target, likewise
    d   = fl32(p - q)                       per coordinate
    t32 = fma32(dz, dz, fma32(dy, dy, fma32(dx, dx, -fl32(w))))
    m   = fma32(fl32(w), 32 u, 32 u R^2 * 1.001),  u = 2^-24,  R = max |p| over the group

Claim: whenever |t32| > m, sign(t32) is the sign of the FP64 decision  d2 > w  with
d2 = fma(dz, dz, fma(dx, dx, dy*dy)) on absolute FP64 coordinates (bh_d2) -- so only tests with |t32| <= m need the FP64
re-evaluation. The cases below include random geometry at many scales, targets far from the origin (where absolute FP32
coordinates would be useless), and pairs placed within a few ulp of the decision boundary."""
import numpy as np

U = 2.0 ** -24
MU = np.float32(32 * U)


def fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64; the sum is rounded once to float64
    (53 bits) and once to float32 -- the double rounding can differ from a true FMA by one float32 ulp in rare ties,
    which the margin (tens of ulp) covers by construction."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def decide(x, c, w, origin, rad):
    """x: targets (T, 3), c: node centres (K, 3), w: radius_sqr (K,). Returns t32 (T, K), m (K,), exact decision (T, K)."""
    p = (x - origin).astype(np.float32)                              # (T, 3)
    q = (c - origin).astype(np.float32)                              # (K, 3)
    w32 = w.astype(np.float32)
    d = (p[:, None, :] - q[None, :, :]).astype(np.float32)           # float32 subtraction
    t = fma32(d[..., 0], d[..., 0], np.broadcast_to(-w32, d.shape[:2]))
    t = fma32(d[..., 1], d[..., 1], t)
    t = fma32(d[..., 2], d[..., 2], t)
    mur2 = np.float32(np.float32(MU * np.float32(rad * rad)) * np.float32(1.001))
    m = fma32(w32, np.full_like(w32, MU), np.full_like(w32, mur2))
    dd = x[:, None, :] - c[None, :, :]                               # FP64, absolute coordinates
    d2 = dd[..., 2] * dd[..., 2] + (dd[..., 0] * dd[..., 0] + dd[..., 1] * dd[..., 1])
    return t, m, d2 > w[None, :]


def group(rng, centre, extent, count=32):
    x = centre + rng.uniform(-extent, extent, (count, 3))
    origin = x[0].copy()
    rad = np.float32(np.abs((x - origin).astype(np.float32)).max())
    return x, origin, rad


def check(x, origin, rad, c, w):
    t, m, exact = decide(x, c, w, origin, rad)
    sure = np.abs(t) > m[None, :]
    wrong = sure & ((t > 0) != exact)
    assert not wrong.any(), "certified test with the wrong sign: t32 %r m %r" % (t[wrong][:3], np.broadcast_to(m, t.shape)[wrong][:3])
    return sure.mean()


def test_random_geometry_at_many_scales():
    rng = np.random.RandomState(1)
    sure_share = []
    for scale in (1e-3, 1e-1, 1.0, 10.0, 100.0):
        for centre_far in (0.0, 150.0, 1e4):                          # groups far from the coordinate origin too
            x, origin, rad = group(rng, np.full(3, centre_far), 0.5 * scale)
            k = 20000
            c = x[0] + rng.normal(0, 1, (k, 3)) * scale * rng.choice([0.3, 1, 3, 30], (k, 1))
            dist2 = ((c - x[0]) ** 2).sum(1)
            w = dist2 * rng.uniform(0.25, 4.0, k)                     # thresholds on both sides of the distance
            w[rng.rand(k) < 0.1] = 0.0                                # leaves
            sure_share.append(check(x, origin, rad, c, w))
    assert min(sure_share) > 0.98                                     # and almost every test IS certified


def test_pairs_within_a_few_ulp_of_the_boundary():
    """w set to the exact d2 of one target +- k FP64 ulp: that target's FP32 test must come out uncertain (or right),
    never certified-wrong; the other targets of the group see the same node at other distances."""
    rng = np.random.RandomState(2)
    for scale, centre in ((1.0, 0.0), (0.05, 120.0), (30.0, -80.0)):
        x, origin, rad = group(rng, np.full(3, centre), 0.4 * scale)
        k = 4000
        c = x[0] + rng.normal(0, 1, (k, 3)) * scale * 4
        j = rng.randint(0, 32, k)
        dd = x[j] - c
        d2 = dd[:, 2] * dd[:, 2] + (dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1])
        for ulps in (-3, -1, 0, 1, 3, 1000, -1000):
            w = d2.copy()
            for _ in range(abs(ulps) if abs(ulps) < 10 else 0):
                w = np.nextafter(w, np.inf if ulps > 0 else -np.inf)
            if abs(ulps) >= 10:
                w = d2 * (1 + ulps * 2.0 ** -52)
            t, m, exact = decide(x, c, w, origin, rad)
            own = np.abs(t[j, np.arange(k)]) <= m                     # the knife-edge pair itself is never certified
            assert own.all()
            check(x, origin, rad, c, w)


def test_coincident_and_self_pairs_are_never_certified():
    """d = 0 (the target's own leaf, or a body at the same place): t32 = -fl32(w) = 0 for a leaf, |t32| <= m: the FP64
    path decides (d2 > 0 is false: not accepted, no children)."""
    rng = np.random.RandomState(3)
    x, origin, rad = group(rng, np.array([40.0, -7.0, 3.0]), 0.7)
    t, m, exact = decide(x, x.copy(), np.zeros(32), origin, rad)
    own = np.arange(32)
    assert (np.abs(t[own, own]) <= m).all() and not exact[own, own].any()
    check(x, origin, rad, x.copy(), np.zeros(32))


def test_margin_is_small_enough_to_be_rare():
    """The uncertain band is 32 u (w + R^2) wide on either side of w: for a node at ten times the group's extent that is
    a relative band of ~2e-6 in d2."""
    rad = np.float32(0.5)
    w = np.float32(25.0)
    m = fma32(np.array([w]), np.array([MU]), np.array([np.float32(MU * rad * rad * np.float32(1.001))]))[0]
    assert m / w < 2.5e-6


# ---- the grouped walk's bookkeeping, restated in Python on the oracle's tree -----------------------------------------
def _grouped_walk(xyzr, n, leaves):
    """Work items (parent, mask) popped 32 at a time off a stack, both children tested against the group's targets,
    accept words -> interaction list, open words of internal children -> new items: the round structure of
    bh_walk_group with exact FP64 tests. Returns per target the SET of accepted nodes, visits and interactions."""
    pos = xyzr[leaves, :3]
    accepted = [set() for _ in leaves]
    visits = inter = 0

    def test(node, mask):
        d = pos - xyzr[node, :3]
        d2 = d[:, 2] * d[:, 2] + (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        acc = (d2 > xyzr[node, 3]) & mask
        return acc, mask & ~acc

    live = np.ones(len(leaves), dtype=bool)
    acc, opened = test(1, live)
    visits += int(live.sum())
    inter += int(acc.sum())
    for j in np.nonzero(acc)[0]:
        accepted[j].add(1)
    stack = [(1, opened)] if opened.any() and n > 1 else []
    while stack:
        block, stack = stack[-32:], stack[:-32]          # one item per lane
        pushed = []
        for parent, mask in block:
            for child in (2 * parent, 2 * parent + 1):
                acc, opened = test(child, mask)
                visits += int(mask.sum())
                inter += int(acc.sum())
                for j in np.nonzero(acc)[0]:
                    accepted[j].add(child)
                if opened.any() and child < n:           # a leaf that is not accepted has no children: dropped
                    pushed.append((child, opened))
        stack.extend(pushed)
    return accepted, visits, inter


def _traverse(xyzr, n, leaf):
    """nbody_space_heap_stackless::traverse (nbody_space_heap_stackless.cpp:3-28) for one target."""
    p = xyzr[leaf, :3]
    out, visits, curr, ts = [], 0, 1, 2 * n

    def skip(i):
        while i & 1:
            i >>= 1
        return i + 1

    while True:
        visits += 1
        d = p - xyzr[curr, :3]
        d2 = d[2] * d[2] + (d[0] * d[0] + d[1] * d[1])
        if d2 > xyzr[curr, 3]:
            out.append(curr)
            curr = skip(curr)
        else:
            curr = 2 * curr if 2 * curr < ts else skip(curr)
        if curr == 1:
            return out, visits


def test_grouped_bookkeeping_accepts_exactly_the_traversals_nodes(oracle64):
    """Every target of a group ends up with exactly the nodes its own stackless traversal accepts (as a set: the grouped
    walk adds them in another order), and the totals equal the C oracle's visit / interaction counts."""
    from conftest import load_golden_npz
    g = load_golden_npz("g1_n256")
    n = 256
    for ratio in (1.0, 3.1623, 10.0):
        t = oracle64.heap_build(g["y"], g["mass"], ratio)
        _, visits_c, inter_c = oracle64.fcompute_bh(g["y"], g["mass"], t)
        xyzr = np.asarray(t["xyzr"], dtype=np.float64)
        visits = inter = visits_t = inter_t = 0
        for first in range(0, n, 32):
            leaves = np.arange(n + first, n + first + 32)
            accepted, v, k = _grouped_walk(xyzr, n, leaves)
            visits += v
            inter += k
            for j, leaf in enumerate(leaves):
                want, vt = _traverse(xyzr, n, leaf)
                visits_t += vt
                inter_t += len(want)
                assert accepted[j] == set(want) and len(want) == len(set(want)), (ratio, leaf)
        assert (visits, inter) == (visits_t, inter_t), ratio
        # the C oracle evaluates d2 with fused multiply-adds (as the reference built with gcc -O3 does), numpy does not:
        # an equal-mass pair sitting exactly on d2 == radius_sqr may fall the other way, nothing else differs
        assert abs(visits - visits_c) <= 1e-3 * visits_c and abs(inter - inter_c) <= 1e-3 * inter_c, ratio
