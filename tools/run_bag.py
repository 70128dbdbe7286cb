#!/usr/bin/env python
"""tools/run_bag.py BAG [--lidar /velodyne_points] [--imu /imu/data] [--max-scans N] [--lidar-model 0|1]

BASELINE.json configs[1] runner (GPU box): replays a ROS1 bag through the restated front end (image projection, feature
extraction, IMU propagation) and the B200 IESKF update, prints the trajectory.  Uncompressed and lz4-compressed bags
are read directly; for bz2 run `python tools/bag_tool.py decompress IN.bag OUT.bag` first."""
import argparse, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ap = argparse.ArgumentParser()
ap.add_argument("bag"); ap.add_argument("--lidar", default="/velodyne_points"); ap.add_argument("--imu", default="/imu/data")
ap.add_argument("--max-scans", type=int, default=0); ap.add_argument("--lidar-model", type=int, default=0)
a = ap.parse_args()
synth = importlib.import_module("lins---lidar-inertial-slam_b200.synth")
out = synth.run_bag(a.bag, a.lidar, a.imu, a.max_scans, a.lidar_model)
print("scans", len(out["status"]), "IESKF updates", len(out["iters"]), "mean iterations %.2f" % (out["iters"].mean() if len(out["iters"]) else 0), "diverged", int(((out["flags"] & 2) != 0).sum()))
np.set_printoptions(precision=4, suppress=True)
for k, (st, g) in enumerate(zip(out["status"], out["global_est"])):
    print(k, int(st), g)
