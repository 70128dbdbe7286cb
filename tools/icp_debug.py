"""ICP fallback: device loop vs the round-1 host loop (LINS_ICP_HOST_LOOP=1) vs the oracle, golden units (diagnostics; GPU box)."""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    capi = importlib.import_module("lins---lidar-inertial-slam_b200.capi")
    defs = importlib.import_module("lins---lidar-inertial-slam_b200.ctypes_defs")
    from oracle import oracle_binding as ob
    b = defs.Batch.load(os.path.join(ROOT, "tests", "golden", "units_inputs.npz"))
    prm = defs.LinsParams.shipped()
    g = capi.LinsGpu(prm)
    for i in range(b.n):
        u = b.unit(i)
        o = ob.Oracle(prm); o.set_map(u["surf_less_flat"], u["corner_less_sharp"]); g.set_map(u["surf_less_flat"], u["corner_less_sharp"])
        t0, q0 = u["state"][:3], u["state"][6:10]
        to, qo, ito, cvo = o.estimate_transform(u["surf_flat"], u["corner_sharp"], t0, q0)
        tg, qg, itg, cvg = g.estimate_transform(u["surf_flat"], u["corner_sharp"], t0, q0)
        print(os.environ.get("LINS_ICP_HOST_LOOP", "device"), i, "iters", itg, ito, "conv", cvg, cvo, "max |dt| %.3g" % np.abs(tg - to).max(), "|dq| %.3g" % np.abs(qg - qo).max())
else:
    for env in ({}, {"LINS_ICP_HOST_LOOP": "1"}):
        subprocess.call([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, **env))
