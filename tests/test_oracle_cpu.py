"""CPU suite, part 1: the oracle itself.

The reference ships no tests or golden vectors and cannot be compiled in this image, so the
oracle is pinned by (a) a second independent pure-Python restatement of the association on small scenes,
(b) brute-force vs kd-tree equality, (c) form A (reference-faithful M x M gain) vs form B (18 x 18) agreement,
(d) analytic cases, and (e) committed golden outputs (regression).  PARITY UNPINNED against the real reference.
"""
import numpy as np
import pytest

import pyref


def _scene(rng, n_rings=6, per_ring=40, jitter=0.02):
    """Ring-sorted target cloud on a wavy ground patch: (T,4) float32 x,y,z,intensity."""
    pts = []
    for r in range(n_rings):
        for k in range(per_ring):
            az = -0.6 + 1.2 * k / per_ring
            rad = 4.0 + 1.3 * r
            x, y = rad * np.cos(az), rad * np.sin(az)
            z = -1.2 + 0.05 * np.sin(3 * x) + jitter * rng.standard_normal()
            pts.append([x + jitter * rng.standard_normal(), y + jitter * rng.standard_normal(), z, r + 0.1 * rng.random() * 0.999])
    return np.asarray(pts, dtype=np.float32)


def test_association_matches_python_restatement(ob, defs):
    rng = np.random.default_rng(7)
    tgt_s = _scene(rng)
    tgt_c = _scene(rng, n_rings=8, per_ring=9, jitter=0.05)
    # queries: perturbed copies of some targets; n_query chosen to exercise the `j < surfPointsFlatNum` quirk:
    # with 25 queries only target indices < 25 are visited by the forward walk
    for n_q in (25, 90, 400):
        qi = rng.integers(0, len(tgt_s), size=min(n_q, 60))
        qs = tgt_s[qi].copy()
        qs[:, :3] += 0.05 * rng.standard_normal((len(qs), 3)).astype(np.float32)
        if n_q > len(qs):  # pad the query COUNT (the loop bound) with far-away points that match nothing
            pad = np.tile(np.array([[500.0, 500.0, 500.0, 0.05]], np.float32), (n_q - len(qs), 1))
            qs = np.concatenate([qs, pad])
        qc = tgt_c[rng.integers(0, len(tgt_c), size=20)].copy()
        qc[:, :3] += 0.05 * rng.standard_normal((20, 3)).astype(np.float32)
        state = np.zeros(19)
        state[:3] = [0.21, -0.03, 0.01]
        ax = np.array([0.002, -0.001, 0.015])
        th = np.linalg.norm(ax)
        state[6:9] = ax / th * np.sin(th / 2)
        state[9] = np.cos(th / 2)
        o = ob.Oracle(ob.LinsParams.shipped(), use_kdtree=False)
        o.set_map(defs.make_points(tgt_s[:, :3], tgt_s[:, 3]), defs.make_points(tgt_c[:, :3], tgt_c[:, 3]))
        out = o.associate(defs.make_points(qs[:, :3], qs[:, 3]), defs.make_points(qc[:, :3], qc[:, 3]), state, 0)
        for i in range(min(len(qs), 60)):
            sel = pyref.transform_to_start(qs[i, :3], qs[i, 3], state[:3], state[6:10])
            assert np.array_equal(sel, out["surf_sel"][i]), "pointSel must be bit-exact"
            assert tuple(out["surf_ind"][i]) == pyref.assoc_surf(sel, tgt_s, len(qs))
        for i in range(len(qc)):
            sel = pyref.transform_to_start(qc[i, :3], qc[i, 3], state[:3], state[6:10])
            assert np.array_equal(sel, out["corner_sel"][i])
            assert tuple(out["corner_ind"][i]) == pyref.assoc_corner(sel, tgt_c, len(qc))
        # the quirk is observable: some walks are cut by the query-count bound
        if n_q == 25:
            assert (out["surf_ind"][:, 1:] >= 0).any()


def test_kdtree_equals_bruteforce(ob, golden_batch):
    rng = np.random.default_rng(3)
    u = golden_batch.unit(1)
    o = ob.Oracle(ob.LinsParams.shipped())
    o.set_map(u["surf_less_flat"], u["corner_less_sharp"])
    for which, cloud in ((0, u["surf_less_flat"]), (1, u["corner_less_sharp"])):
        base = np.stack([cloud["x"], cloud["y"], cloud["z"]], 1)
        q = np.concatenate([base[rng.integers(0, len(base), 400)] + rng.normal(0, 0.3, (400, 3)),
                            rng.uniform(-60, 60, (200, 3)), base[:50]]).astype(np.float32)
        ik, dk = o.nn(which, q, True)
        ib, db = o.nn(which, q, False)
        assert np.array_equal(ik, ib)
        assert np.array_equal(dk, db)


def test_kdtree_tie_break_lowest_index(ob, defs):
    # duplicated targets: every query has exact ties; both searches must return the LOWEST index
    rng = np.random.default_rng(5)
    base = rng.uniform(-5, 5, (64, 3)).astype(np.float32)
    tgt = np.concatenate([base, base, base])
    inten = np.repeat(np.arange(3), 64).astype(np.float32)
    o = ob.Oracle(ob.LinsParams.shipped())
    o.set_map(defs.make_points(tgt, inten), defs.make_points(tgt[:8], inten[:8]))
    ik, _ = o.nn(0, base, True)
    ib, _ = o.nn(0, base, False)
    assert np.array_equal(ik, np.arange(64)) and np.array_equal(ib, np.arange(64))


def test_fixtures_have_no_exact_nn_ties(ob, golden_batch):
    """FLANN's tie-break is traversal dependent; ours is lowest-index.  The choice must be unobservable on the
    fixtures: the best and second-best squared distances of every query differ."""
    for i in range(golden_batch.n):
        u = golden_batch.unit(i)
        o = ob.Oracle(ob.LinsParams.shipped(), use_kdtree=False)
        o.set_map(u["surf_less_flat"], u["corner_less_sharp"])
        out = o.associate(u["surf_flat"], u["corner_sharp"], u["state"], 0)
        t = np.stack([u["surf_less_flat"][k] for k in "xyz"], 1)
        for sel, c in zip(out["surf_sel"], out["surf_ind"][:, 0]):
            d = ((t - sel) ** 2).astype(np.float32).sum(1)
            if c >= 0:
                assert (d == d[c]).sum() == 1


def test_form_a_equals_form_b(ob, golden_batch):
    """the 18x18 information form is the push-through identity of the reference's MxM gain."""
    for prm in (ob.LinsParams.shipped(), ob.LinsParams.shipped(num_iter=10, force_all_iters=1)):
        for i in range(golden_batch.n):
            u = golden_batch.unit(i)
            res = []
            for form in (ob.FORM_A, ob.FORM_B):
                o = ob.Oracle(prm)
                o.set_map(u["surf_less_flat"], u["corner_less_sharp"])
                res.append(o.ieskf(u["surf_flat"], u["corner_sharp"], u["state"], u["cov"], form))
            (sa, ca, ra), (sb, cb, rb) = res
            assert ra.iters == rb.iters and ra.converged == rb.converged and ra.diverged == rb.diverged
            assert list(ra.m_surf[: ra.iters]) == list(rb.m_surf[: rb.iters])
            assert np.abs(sa - sb).max() <= 1e-9
            assert np.abs(ca - cb).max() <= 1e-9 * max(1.0, np.abs(ca).max())


def test_golden_regression(ob, golden_batch, golden_out):
    for tag, prm in (("shipped", ob.LinsParams.shipped()), ("forced10", ob.LinsParams.shipped(num_iter=10, force_all_iters=1))):
        for i in range(golden_batch.n):
            u = golden_batch.unit(i)
            o = ob.Oracle(prm, use_kdtree=True)
            o.set_map(u["surf_less_flat"], u["corner_less_sharp"])
            so, co, rep, tr = o.ieskf_trace(u["surf_flat"], u["corner_sharp"], u["state"], u["cov"], ob.FORM_B)
            g = lambda k: golden_out[f"{tag}_{i}_{k}"]  # noqa: E731
            assert [rep.iters, rep.converged, rep.diverged, rep.has_nan] == list(g("iters"))
            assert np.array_equal(tr["surf_ind"], g("surf_ind")) and np.array_equal(tr["corner_ind"], g("corner_ind"))
            assert np.array_equal(tr["surf_mask"], g("surf_mask")) and np.array_equal(tr["corner_mask"], g("corner_mask"))
            assert np.allclose(so, g("state"), rtol=0, atol=1e-12)
            assert np.allclose(co, g("cov"), rtol=1e-9, atol=1e-15)
    if True:
        assert golden_out["forced10_0_iters"][0] == 10  # BASELINE.json configs[0]: "10 ESKF iters"


def test_pose_moves_toward_truth(ob, synth):
    """Sanity of the restated signs: from a deliberately poor prior the update must land near the truth."""
    b = synth.generate("config3", n=8, seed0=50, prior_vel_sigma=1.0)
    so, _, res, _, _ = ob.ieskf_batch(ob.LinsParams.shipped(), b, threads=4)
    prior = np.sqrt(((b.state[:, :3] - b.truth[:, :3]) ** 2).mean())
    post = np.sqrt(((so[:, :3] - b.truth[:, :3]) ** 2).mean())
    assert prior > 0.05 and post < 0.35 * prior


def test_identity_motion_exact_plane_is_rejected(ob, defs):
    """G1: zero motion, queries lying exactly in their target plane => res == 0 => rejected by `res != 0`
    (StateEstimator.hpp:942), so M = 0 and the update equals the prior difference (zero)."""
    xs, ys = np.meshgrid(np.arange(2.0, 12.0, 0.5), np.arange(-4.0, 4.0, 0.5))
    ring = np.clip(((xs - 2.0) // 2.5).astype(int), 0, 3)
    order = np.argsort(ring.ravel(), kind="stable")
    tgt = np.stack([xs.ravel(), ys.ravel(), np.full(xs.size, -1.0)], 1)[order].astype(np.float32)
    inten = ring.ravel()[order].astype(np.float32)
    q = np.array([[5.25, 0.25, -1.0], [6.25, 1.25, -1.0], [8.75, -2.25, -1.0]], np.float32)
    state = np.zeros(19)
    state[9] = 1.0
    state[18] = -9.81
    cov = np.eye(18).ravel() * 1e-4
    o = ob.Oracle(ob.LinsParams.shipped())
    o.set_map(defs.make_points(tgt, inten), defs.make_points(tgt[:6], inten[:6]))
    out = o.associate(defs.make_points(q, [1.0, 1.0, 2.0]), defs.make_points(q[:0], []), state, 0)
    assert (out["surf_ind"] >= 0).all() and not out["surf_mask"].any()
    so, co, rep = o.ieskf(defs.make_points(q, [1.0, 1.0, 2.0]), defs.make_points(q[:0], []), state, cov)
    assert rep.iters == 1 and rep.converged == 1 and rep.m_surf[0] == 0
    assert np.allclose(so, state)


def test_empty_inputs(ob, defs):
    e = defs.make_points(np.zeros((0, 3)), [])
    one = defs.make_points([[1.0, 2.0, 3.0]], [0.05])
    state = np.zeros(19)
    state[9] = 1.0
    cov = np.eye(18).ravel() * 1e-4
    for m_s, m_c, q_s, q_c in ((e, e, e, e), (e, e, one, one), (one, one, e, e)):
        o = ob.Oracle(ob.LinsParams.shipped())
        o.set_map(m_s, m_c)
        so, co, rep = o.ieskf(q_s, q_c, state, cov)
        assert rep.iters == 1 and rep.converged == 1 and rep.diverged == 0
        assert np.allclose(so, state)


def test_small_matrix_helpers(ob):
    rng = np.random.default_rng(1)
    L = ob.lib()
    for _ in range(20):
        B = rng.standard_normal((6, 6))
        A = B @ B.T + 0.1 * np.eye(6)
        E, V = np.zeros(6), np.zeros((6, 6))
        L.lins_oracle_sym_eig6(ob.ptr(np.ascontiguousarray(A)), ob.ptr(E), ob.ptr(V))
        assert np.allclose(E, np.linalg.eigvalsh(A), rtol=1e-10)
        assert np.allclose(A @ V, V * E, atol=1e-9)
        b, x = rng.standard_normal(6), np.zeros(6)
        L.lins_oracle_qr_solve6(ob.ptr(np.ascontiguousarray(A)), ob.ptr(b), ob.ptr(x))
        assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
    # boxPlus / boxMinus round trip (KalmanFilter.hpp:71-94)
    s = rng.standard_normal(19)
    s[6:10] /= np.linalg.norm(s[6:10])
    dx = 0.1 * rng.standard_normal(18)
    s2, back = np.zeros(19), np.zeros(18)
    L.lins_oracle_boxplus(ob.ptr(s), ob.ptr(dx), ob.ptr(s2))
    L.lins_oracle_boxminus(ob.ptr(s2), ob.ptr(s), ob.ptr(back))
    assert np.allclose(back, dx, atol=1e-12)


def test_transform_to_end_inverts_start(ob, defs):
    """transformToEnd (:1083-1101) = undo the end pose after transformToStart."""
    rng = np.random.default_rng(2)
    pts = defs.make_points(rng.uniform(-20, 20, (50, 3)), rng.integers(0, 16, 50) + 0.1 * rng.random(50))
    state = np.zeros(19)
    state[:3] = [0.3, -0.1, 0.02]
    ax = np.array([0.01, -0.02, 0.03])
    th = np.linalg.norm(ax)
    state[6:9], state[9] = ax / th * np.sin(th / 2), np.cos(th / 2)
    prm = ob.LinsParams.shipped()
    a, b = np.zeros(50, defs.POINT_DTYPE), np.zeros(50, defs.POINT_DTYPE)
    L = ob.lib()
    import ctypes as C
    L.lins_oracle_transform(C.byref(prm), ob.ptr(state), 0, ob.ptr(pts), 50, ob.ptr(a))
    L.lins_oracle_transform(C.byref(prm), ob.ptr(state), 1, ob.ptr(pts), 50, ob.ptr(b))
    th2 = 2 * np.arctan2(np.linalg.norm(state[6:9]), state[9])
    k = state[6:9] / np.linalg.norm(state[6:9])
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th2) * K + (1 - np.cos(th2)) * K @ K
    A = np.stack([a["x"], a["y"], a["z"]], 1).astype(float)
    Bv = np.stack([b["x"], b["y"], b["z"]], 1).astype(float)
    assert np.allclose(Bv, (A - state[:3]) @ R, atol=2e-5)
