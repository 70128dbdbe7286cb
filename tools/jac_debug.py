"""Compare the 28 sums of the two split-Jacobian kernels (LINS_JAC_VARIANT 0 = tensor-core fold, 2 = shuffle fold) on the golden units."""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    capi = importlib.import_module("lins---lidar-inertial-slam_b200.capi")
    defs = importlib.import_module("lins---lidar-inertial-slam_b200.ctypes_defs")
    b = defs.Batch.load(os.path.join(ROOT, "tests", "golden", "units_inputs.npz"))
    g = capi.LinsGpu(defs.LinsParams.shipped(num_iter=1, force_all_iters=1))
    g.batch_upload(b); g.batch_run(); g.sync()
    acc = g.batch_jacobian_pass(want_accum=True)
    np.save(sys.argv[2], acc)
else:
    subprocess.check_call([sys.executable, os.path.abspath(__file__), "child", "/tmp/jac_ref.npy"], env=dict(os.environ, LINS_JAC_VARIANT="2"))
    b = np.load("/tmp/jac_ref.npy")
    np.set_printoptions(linewidth=200, precision=3)
    for lib in [""] + sys.argv[1:]:
        env = dict(os.environ, LINS_JAC_VARIANT="0")
        if lib:
            env["LINS_GPU_LIB"] = os.path.join(ROOT, "variants", f"liblins_gpu_{lib}.so")
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "child", "/tmp/jac_new.npy"], env=env)
        a = np.load("/tmp/jac_new.npy")
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        print(lib or "main", "max rel diff per entry:", rel[:, :30].max(0))
