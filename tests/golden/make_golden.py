"""Generates the committed golden fixtures (run from the repo root: python tests/golden/make_golden.py).

The reference ships no tests / golden vectors and cannot be built here, so the goldens are
(a) synthetic inputs from tools/synth (seeded) and (b) the CPU oracle's outputs on them (form B, shipped params
and the "10 iterations forced" gate of BASELINE.json configs[0]).  PARITY UNPINNED (oracle/lins_oracle.hpp).
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
synth = importlib.import_module("lins---lidar-inertial-slam_b200.synth")
defs = importlib.import_module("lins---lidar-inertial-slam_b200.ctypes_defs")
from oracle import oracle_binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    b1 = synth.generate("config1", n=1, seed0=1)
    b3 = synth.generate("config3", n=3, seed0=1000)
    clouds = {k: np.concatenate([b1.clouds[k], b3.clouds[k]]) for k in defs.Batch.FIELDS}
    offsets = {k: np.concatenate([b1.offsets[k], b3.offsets[k][1:] + b1.offsets[k][-1]]) for k in defs.Batch.FIELDS}
    b = defs.Batch(clouds, offsets, np.concatenate([b1.state, b3.state]), np.concatenate([b1.cov, b3.cov]),
                   np.concatenate([b1.truth, b3.truth]))
    b.save(os.path.join(HERE, "units_inputs.npz"))
    b = defs.Batch.load(os.path.join(HERE, "units_inputs.npz"))  # what the tests will see
    out = {}
    for tag, prm in (("shipped", ob.LinsParams.shipped()), ("forced10", ob.LinsParams.shipped(num_iter=10, force_all_iters=1))):
        for i in range(b.n):
            u = b.unit(i)
            o = ob.Oracle(prm, use_kdtree=False)
            o.set_map(u["surf_less_flat"], u["corner_less_sharp"])
            so, co, rep, tr = o.ieskf_trace(u["surf_flat"], u["corner_sharp"], u["state"], u["cov"], ob.FORM_B)
            out[f"{tag}_{i}_state"] = so
            out[f"{tag}_{i}_cov"] = co
            out[f"{tag}_{i}_iters"] = np.array([rep.iters, rep.converged, rep.diverged, rep.has_nan])
            out[f"{tag}_{i}_m"] = np.array([list(rep.m_surf[: rep.iters]), list(rep.m_corner[: rep.iters])])
            out[f"{tag}_{i}_rnorm"] = np.array(rep.residual_norm[: rep.iters])
            out[f"{tag}_{i}_unorm"] = np.array(rep.update_norm[: rep.iters])
            out[f"{tag}_{i}_surf_ind"] = tr["surf_ind"]
            out[f"{tag}_{i}_corner_ind"] = tr["corner_ind"]
            out[f"{tag}_{i}_surf_mask"] = tr["surf_mask"]
            out[f"{tag}_{i}_corner_mask"] = tr["corner_mask"]
            out[f"{tag}_{i}_lin_state"] = tr["lin_state"]
    np.savez_compressed(os.path.join(HERE, "units_oracle_outputs.npz"), **out)
    print("wrote", {k: len(v) for k, v in b.clouds.items()})


if __name__ == "__main__":
    main()
