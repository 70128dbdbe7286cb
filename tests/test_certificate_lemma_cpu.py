"""The exactness argument behind the search certificates of the fused kernel (lins_assoc_az.cuh: cert_accepted),
checked on a numpy model: if, at the query's new position, the stored winner re-evaluated exactly still beats the
stored runner-up (exact (f32 distance bits, index) keys), is inside the gate, and is closer than
bound - moved - 2e-4 m — bound = the third best distance at the search position — then a brute-force search at the
new position returns the same winner.  Includes the adversarial case the plain winner / runner-up margin could not
certify: a query almost midway between two neighbouring targets."""
import numpy as np

f32 = np.float32
GATE = f32(25.0)


def sqd(q, T):
    d = (q[None, :] - T).astype(f32)
    return ((d[:, 0] * d[:, 0]) + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def keys(q, T):
    return (sqd(q, T).view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.arange(len(T), dtype=np.uint64)


def test_two_front_runner_certificate_never_certifies_a_changed_answer():
    rng = np.random.default_rng(5)
    certified = flips_caught = 0
    for trial in range(4000):
        n = int(rng.integers(3, 60))
        T = (rng.standard_normal((n, 3)) * rng.uniform(0.05, 3.0)).astype(f32)
        p0 = (rng.standard_normal(3) * 0.5).astype(f32)
        if trial % 3 == 0:  # near-tie: the query starts almost midway between two targets
            a, b = T[0], T[1]
            p0 = ((a + b) / 2 + rng.standard_normal(3).astype(f32) * f32(1e-3)).astype(f32)
        k0 = keys(p0, T)
        order = np.argsort(k0)
        w, r = int(order[0]), int(order[1])
        if not sqd(p0, T)[w] < GATE:
            continue
        bound = np.sqrt(sqd(p0, T)[order[2]])  # everything but the two front-runners was at least this far
        step = (rng.standard_normal(3) * rng.uniform(1e-4, 0.3)).astype(f32)
        p1 = (p0 + step).astype(f32)
        moved = np.sqrt(sqd(p1, p0[None, :])[0])
        k1 = keys(p1, T)
        dw = sqd(p1, T)[w]
        ok = bool(dw < GATE) and bool(k1[w] < k1[r]) and bool(np.sqrt(dw) + moved + f32(2e-4) < bound)
        truth = int(np.argmin(k1))
        if ok:
            certified += 1
            assert truth == w, (trial, w, r, truth)
        elif truth != w:
            flips_caught += 1
    assert certified > 1000 and flips_caught > 50  # the rule is neither vacuous nor blind
