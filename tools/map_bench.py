"""Row F2 alone: time lins_gpu_scan2map on one synthetic 50-key-frame unit (diagnostics / ncu target; GPU box)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
capi = importlib.import_module("lins---lidar-inertial-slam_b200.capi")
synth = importlib.import_module("lins---lidar-inertial-slam_b200.synth")
defs = importlib.import_module("lins---lidar-inertial-slam_b200.ctypes_defs")
kf = int(sys.argv[1]) if len(sys.argv) > 1 else 50
u = synth.generate_map_unit("config3", seed=40, n_keyframes=kf, sigma_t=0.1, sigma_r=0.01)
g = capi.LinsGpu(defs.LinsParams.shipped())
g.map_set(u.corner_map, u.surf_map)
for _ in range(3):
    T, rep = g.scan2map(u.corner_last, u.surf_last, u.guess)
t0 = time.perf_counter()
for _ in range(20):
    T, rep = g.scan2map(u.corner_last, u.surf_last, u.guess)
ms = (time.perf_counter() - t0) * 1e3 / 20
print("map %d + %d points, features %d + %d, LM iterations %d, %.3f ms per scan2map call" % (len(u.corner_map), len(u.surf_map), len(u.corner_last), len(u.surf_last), rep.iters, ms))
