"""Property test (CPU, numpy float32) of the elevation-band lower bound the closest-point scans prune rings with
(lins_assoc_az.cuh: slope_of / ring_lower_bound).  The kernel skips a ring when this bound exceeds the
search radius, so exactness of the search needs bound <= distance to EVERY target of the ring, for every geometry —
including the f32 rounding of the slopes and of the bound itself.  The GPU fuzz test checks the consequence (bit-equal
indices); this one checks the inequality directly over far more geometries than the scenes contain."""
import numpy as np

F = np.float32


def slope_of(p):  # z / sqrt(x*x + y*y), each operation rounded to f32 (the kernel's __fmul_rn / __fadd_rn / sqrtf / __fdiv_rn)
    x, y, z = (p[..., k].astype(F) for k in range(3))
    with np.errstate(divide="ignore", invalid="ignore"):
        return (z / np.sqrt((x * x).astype(F) + (y * y).astype(F), dtype=F)).astype(F)


def ring_lower_bound(lo, hi, zq, rq):  # the kernel's expression; fmaf emulated in f64 (exact product, one rounding)
    lo64, hi64, zq64, rq64 = (np.asarray(v, dtype=np.float64) for v in (lo, hi, zq, rq))
    with np.errstate(over="ignore", invalid="ignore"):
        a = (lo64 * rq64 - zq64).astype(F)
        b = (-hi64 * rq64 + zq64).astype(F)
    m = np.where(a > b, lo, hi).astype(F)
    with np.errstate(over="ignore", invalid="ignore"):
        inv = (1.0 / np.sqrt((m.astype(np.float64) * m.astype(np.float64) + 1.0).astype(F), dtype=F)).astype(F)
        bound = (np.maximum(a, b) * inv).astype(F) * F(0.998) - F(1.0e-4)
        unbounded = ~((hi - lo) < F(1.0e30))
    return np.where(unbounded, F(-1.0), bound).astype(F)


def ring_targets(rng, n, elev_deg, jitter_deg, rmin, rmax):
    """n points of one lidar ring: elevation elev +- jitter, any azimuth, ranges in [rmin, rmax]"""
    el = np.deg2rad(elev_deg + rng.uniform(-jitter_deg, jitter_deg, n))
    az = rng.uniform(-np.pi, np.pi, n)
    r = rng.uniform(rmin, rmax, n)
    return np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], -1).astype(F)


def check(rng, queries, targets):
    sl = slope_of(targets)
    finite = np.abs(sl) < F(1.0e30)
    lo, hi = (F(-3.0e38), F(3.0e38)) if not finite.all() else (sl.min(), sl.max())  # (az_build: a target on the z axis unbounds the band)
    q = queries.astype(F)
    rq = np.sqrt((q[:, 0] * q[:, 0]).astype(F) + (q[:, 1] * q[:, 1]).astype(F), dtype=F)
    bound = ring_lower_bound(np.full(len(q), lo, F), np.full(len(q), hi, F), q[:, 2], rq)
    d = np.sqrt(((q[:, None, :].astype(np.float64) - targets[None, :, :].astype(np.float64)) ** 2).sum(-1)).min(1)
    bad = ~(bound.astype(np.float64) <= d) & ~np.isnan(bound)  # (NaN: the kernel's `bound > B` is false, the ring is scanned)
    assert not bad.any(), (q[bad][:3], bound[bad][:3], d[bad][:3], lo, hi)
    return bound, d


def test_bound_never_exceeds_the_distance_to_any_target_of_the_ring():
    rng = np.random.default_rng(7)
    tight = 0
    for trial in range(300):
        elev = rng.uniform(-30, 30)
        t = ring_targets(rng, 400, elev, rng.choice([0.0, 0.05, 0.3]), 0.3, rng.choice([5.0, 40.0, 120.0]))
        # queries: other rings of the same sensor, points near the ring, points near the z axis, far points
        qs = [ring_targets(rng, 200, elev + rng.uniform(-8, 8), 0.2, 0.3, 120.0),
              t[rng.integers(0, len(t), 100)] + rng.normal(0, 0.05, (100, 3)).astype(F),
              np.concatenate([rng.normal(0, 1e-3, (50, 2)), rng.uniform(-5, 5, (50, 1))], 1),
              rng.uniform(-150, 150, (100, 3))]
        bound, d = check(rng, np.concatenate(qs).astype(F), t)
        tight += int(((bound > 0.5 * d) & (d > 0.2)).sum())
    assert tight > 1000  # the bound is not vacuous: it gets within a factor 2 of the true distance for many queries


def test_degenerate_bands_never_prune_wrongly():
    rng = np.random.default_rng(11)
    # a target on the z axis (rho = 0: slope +-inf / NaN) unbounds the band
    t = ring_targets(rng, 50, 10.0, 0.1, 1.0, 30.0)
    t[0] = (0.0, 0.0, 2.0)
    bound, _ = check(rng, rng.uniform(-20, 20, (500, 3)).astype(F), t)
    assert (bound == F(-1.0)).all()
    # near-vertical rings (slopes ~ 1e3 .. 1e6) and queries on the axis
    for elev in (89.0, 89.99, -89.9):
        t = ring_targets(rng, 200, elev, 0.001, 0.5, 50.0)
        q = np.concatenate([rng.uniform(-20, 20, (300, 3)), np.zeros((50, 3)), np.array([[0, 0, 3.0], [0, 0, -3.0]])]).astype(F)
        check(rng, q, t)
    # a single-target ring, a query exactly on it
    t = ring_targets(rng, 1, -15.0, 0.0, 10.0, 10.0)
    bound, d = check(rng, np.concatenate([t, t * F(1.5), rng.uniform(-30, 30, (200, 3)).astype(F)]), t)
    assert bound[0] <= 0 and d[0] == 0
