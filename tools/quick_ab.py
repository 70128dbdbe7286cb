"""A/B timing of the fused kernel on one resident 1000-unit batch (diagnostics; run on a GPU box).
Prints ms per launch (wall clock around a stream sync, best and median of N) and a checksum of the results so that
variants can be compared for bit-identical outputs."""
import importlib, sys, os, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
capi = importlib.import_module("lins---lidar-inertial-slam_b200.capi")
synth = importlib.import_module("lins---lidar-inertial-slam_b200.synth")
defs = importlib.import_module("lins---lidar-inertial-slam_b200.ctypes_defs")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
config = sys.argv[3] if len(sys.argv) > 3 else "config3"
b = synth.generate(config, n=n, seed0=int(os.environ.get("QA_SEED", "1000")))
if os.environ.get("QA_TILE"):
    b = b.tile(int(os.environ["QA_TILE"]))
g = capi.LinsGpu(defs.LinsParams.shipped())
g.batch_upload(b)
for _ in range(3):
    g.batch_run()
g.sync()
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); g.batch_run(); g.sync(); ts.append(1e3 * (time.perf_counter() - t0))
st, cov, res, _ = g.batch_download(states=True, covs=True)
h = hashlib.md5(res["iters"].tobytes() + res["flags"].tobytes() + st.tobytes() + cov.tobytes()).hexdigest()[:12]
its = int(res["iters"].sum())
print(f"{config} n={n}: best {min(ts):.3f} ms median {np.median(ts):.3f} ms  iterations {its}  ({its / min(ts) / 1e3:.3f} M it/s)  checksum {h}")
