"""The device's slab test (csrc/dev_accel.cuh box_geom) is not a transliteration of Bounds3::intersect_p_precomp
(bounds.rs:190-232): it folds the reference's compare-and-select chains into two 3-input min / max operations.  This file
pins the claim that the two give the same accept / reject decision for EVERY input, including the ones that produce NaN
(0 * inf: origin exactly on a slab plane of an axis the ray is parallel to), infinities and signed zeros:

  reference   t_min = tx0; t_max = tx1 * g;  miss if t_min > ty1 * g or ty0 > t_max;  t_min = ty0 > t_min ? ty0 : t_min; ...
              accept iff no miss and t_min < ray.t_max and t_max > 0      (a NaN never compares true, so a NaN that reaches
              the accumulators from the x axis sticks and rejects, one from y / z is ignored)
  device      t_min = max3(tx0, ty0, tz0);  t_max = min3(tx1, ty1, tz1) * g   (NaN operands ignored, as PTX max.f32 / min.f32 do;
              x -> fl(x * g) is monotonic, so scaling after the min is the min of the scaled values)
              accept iff !(t_min > t_max) and t_max > 0 and tx0, tx1 are numbers and t_min < ray.t_max

numpy float32 arithmetic is IEEE like the device's (-fmad=false) and np.fmax / np.fmin ignore NaN like max.f32 / min.f32."""
import numpy as np

F = np.float32
G = F(1.0) + F(2.0) * (F(3.0) * F(2.0 ** -24) / (F(1.0) - F(3.0) * F(2.0 ** -24)))  # 1 + 2 gamma(3), math.rs:8-10


def _slab_inputs(rng, n):
    special = np.array([0.0, -0.0, 1.0, -1.0, 0.5, 2.0, 1e-30, -1e-30, 1e30, -1e30, 3.0, -3.0, 1.0000001, 0.99999994], dtype=F)
    pick = lambda shape: np.where(rng.random(shape) < 0.5, special[rng.integers(0, special.size, shape)], rng.normal(size=shape).astype(F) * F(2.0)).astype(F)
    a, b = pick((n, 3)), pick((n, 3))
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    flat = rng.random((n, 3)) < 0.15
    hi = np.where(flat, lo, hi)  # degenerate (flat) boxes
    o = pick((n, 3))
    on_plane = rng.random((n, 3))
    o = np.where(on_plane < 0.2, lo, np.where(on_plane < 0.4, hi, o)).astype(F)  # origin exactly on a slab plane
    d = pick((n, 3))
    with np.errstate(divide="ignore"):
        inv = (F(1.0) / d).astype(F)  # +-inf for d = +-0
    t_ray = np.where(rng.random(n) < 0.3, F(np.inf), np.abs(pick((n,)))).astype(F)
    return lo, hi, o, inv, t_ray


def _planes(lo, hi, o, inv):
    neg = inv < 0  # dir_is_neg (ray.rs): inv_dir < 0
    with np.errstate(invalid="ignore", over="ignore"):
        t0 = ((np.where(neg, hi, lo) - o) * inv).astype(F)
        t1 = ((np.where(neg, lo, hi) - o) * inv).astype(F)
    return t0, t1


def reference_accept(lo, hi, o, inv, t_ray):
    t0, t1 = _planes(lo, hi, o, inv)
    with np.errstate(invalid="ignore", over="ignore"):
        t_min, t_max = t0[:, 0], (t1[:, 0] * G).astype(F)
        ty_max, tz_max = (t1[:, 1] * G).astype(F), (t1[:, 2] * G).astype(F)
        miss = (t_min > ty_max) | (t0[:, 1] > t_max)
        t_min = np.where(t0[:, 1] > t_min, t0[:, 1], t_min)
        t_max = np.where(ty_max < t_max, ty_max, t_max)
        miss |= (t_min > tz_max) | (t0[:, 2] > t_max)
        t_min = np.where(t0[:, 2] > t_min, t0[:, 2], t_min)
        t_max = np.where(tz_max < t_max, tz_max, t_max)
        return ~miss & (t_min < t_ray) & (t_max > 0), t_min


def device_accept(lo, hi, o, inv, t_ray):
    t0, t1 = _planes(lo, hi, o, inv)
    with np.errstate(invalid="ignore", over="ignore"):
        t_min = np.fmax(np.fmax(t0[:, 0], t0[:, 1]), t0[:, 2])
        t_max = (np.fmin(np.fmin(t1[:, 0], t1[:, 1]), t1[:, 2]) * G).astype(F)
        numbers = ~np.isnan(t0[:, 0]) & ~np.isnan(t1[:, 0])
        return ~(t_min > t_max) & (t_max > 0) & numbers & (t_min < t_ray), t_min


def test_min_max_slab_test_decides_like_the_reference_chain():
    rng = np.random.default_rng(2024)
    total = nan_cases = accepted = 0
    for _ in range(40):
        lo, hi, o, inv, t_ray = _slab_inputs(rng, 250_000)
        ref, ref_t = reference_accept(lo, hi, o, inv, t_ray)
        dev, dev_t = device_accept(lo, hi, o, inv, t_ray)
        assert np.array_equal(ref, dev), f"{np.count_nonzero(ref != dev)} decisions differ"
        # the entry distance that goes on the traversal stack: equal as a number wherever the box is accepted
        assert np.array_equal(ref_t[ref], dev_t[ref])
        t0, t1 = _planes(lo, hi, o, inv)
        nan_cases += int(np.count_nonzero(np.isnan(t0).any(axis=1) | np.isnan(t1).any(axis=1)))
        accepted += int(np.count_nonzero(ref))
        total += ref.size
    assert nan_cases > total // 20 and accepted > total // 50  # the sweep does reach the NaN paths and both outcomes
