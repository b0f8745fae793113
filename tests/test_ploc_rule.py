"""The pairing rule of the device BVH builder's PLOC rounds (csrc/k_bvh.cu ploc_nn_kernel / ploc_merge_kernel), restated in
numpy: every cluster picks, within R positions of its place in the Morton-ordered sequence, the neighbour whose union with it
has the smallest surface area; ties go to the nearer position, then to the pair whose lower position is even, then to the
lower position.  Two clusters that picked each other merge.  What the rule has to guarantee — and what this file checks on
the CPU, where the CUDA code cannot run — is (1) progress: some pair is always mutual, so every round shortens the
sequence; (2) no degenerate peeling: a run of equal distances (duplicate or gridded primitives) pairs up as (0,1), (2,3), ...
and halves per round instead of losing one pair off its end."""
import numpy as np

R = 16


def _area(mn, mx):
    d = mx - mn
    return 2.0 * (d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0])


def nearest(mn, mx):
    """the kernel's candidate order: distance 1 .. R; within a distance the pair with the even lower position first"""
    m = mn.shape[0]
    best = np.full(m, np.inf, dtype=np.float32)
    best_j = np.full(m, -1, dtype=np.int64)
    idx = np.arange(m)
    for d in range(1, R + 1):
        up_first = (d & 1) == 1
        for k in range(2):
            step = np.where((idx & 1) == 0, d, -d) if up_first else np.full(m, -d)
            if k == 1:
                step = -step
            j = idx + step
            ok = (j >= 0) & (j < m)
            jj = np.clip(j, 0, m - 1)
            a = _area(np.minimum(mn, mn[jj]), np.maximum(mx, mx[jj])).astype(np.float32)
            a = np.minimum(a, np.float32(3.0e38))
            take = ok & (a < best)
            best = np.where(take, a, best)
            best_j = np.where(take, j, best_j)
    return best_j


def one_round(mn, mx):
    m = mn.shape[0]
    nn = nearest(mn, mx)
    idx = np.arange(m)
    mutual = (nn >= 0) & (nn[np.clip(nn, 0, m - 1)] == idx)
    lower = mutual & (idx < nn)
    drop = mutual & (idx > nn)
    out_mn, out_mx = mn.copy(), mx.copy()
    out_mn[lower] = np.minimum(mn[lower], mn[nn[lower]])
    out_mx[lower] = np.maximum(mx[lower], mx[nn[lower]])
    keep = ~drop
    return out_mn[keep], out_mx[keep], int(lower.sum()), nn


def _run(mn, mx, max_rounds=400):
    rounds = 0
    while mn.shape[0] > 1:
        mn, mx, merges, _ = one_round(mn, mx)
        assert merges >= 1, "a round without a mutual pair"
        rounds += 1
        assert rounds <= max_rounds
    return rounds


def test_equal_distances_pair_up_level_by_level():
    for m in (2, 3, 17, 64, 1000):
        mn = np.zeros((m, 3), dtype=np.float32)
        mx = np.ones((m, 3), dtype=np.float32)
        _, _, merges, nn = one_round(mn, mx)
        assert merges == m // 2
        assert np.array_equal(nn[: 2 * (m // 2)], np.arange(2 * (m // 2)) ^ 1)  # (0,1), (2,3), ...
        assert _run(mn, mx) == int(np.ceil(np.log2(m)))


def test_every_round_merges_and_the_rounds_stay_logarithmic():
    rng = np.random.default_rng(3)
    for m in (5, 200, 5000):
        c = np.sort(rng.uniform(0, 1, (m, 1)), axis=0) * np.float32([1, 0.3, 0.1]) + rng.uniform(0, 0.02, (m, 3))
        mn = c.astype(np.float32)
        mx = (c + rng.uniform(0.0, 0.01, (m, 3))).astype(np.float32)
        assert _run(mn, mx) <= 6 * int(np.ceil(np.log2(m))) + 4
    # a regular grid in row order: many exact ties between the left and the right neighbour
    g = 32
    xs, ys = np.meshgrid(np.arange(g, dtype=np.float32), np.arange(g, dtype=np.float32), indexing="ij")
    mn = np.stack([xs.ravel(), ys.ravel(), np.zeros(g * g, dtype=np.float32)], axis=1)
    assert _run(mn, mn + np.float32(1.0)) <= 40
