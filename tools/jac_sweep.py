"""Time the split Jacobian kernel for every launch-bounds variant (diagnostics; GPU box)."""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    capi = importlib.import_module("lins---lidar-inertial-slam_b200.capi")
    synth = importlib.import_module("lins---lidar-inertial-slam_b200.synth")
    defs = importlib.import_module("lins---lidar-inertial-slam_b200.ctypes_defs")
    b = synth.generate("config3", n=1000, seed0=1000).tile(5)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g = capi.LinsGpu(defs.LinsParams.shipped(), device=0, stream=stream.cuda_stream)
    g.batch_upload(b); g.batch_run(); torch.cuda.synchronize()
    for _ in range(3):
        g.batch_jacobian_pass()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        g.batch_jacobian_pass()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 20
    ns = int(b.offsets["surf_flat"][-1]); nc = int(b.offsets["corner_sharp"][-1])
    byts = 76 * ns + 56 * nc + 384 * b.n
    print("variant %s: %.1f us  %.0f GB/s algorithmic" % (os.environ.get("LINS_JAC_VARIANT", "0"), ms * 1e3, byts / ms / 1e6))
    if os.environ.get("LINS_GATHER_CAL"):
        # calibration: what a pure random 16-B gather (one 32-B DRAM sector per access) over a working set of the same size
        # reaches on this GPU — the access pattern that bounds the Jacobian kernel's target gathers
        n = 408 * 1000 * 1000 // 16
        table = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        m = 8 * 1000 * 1000
        idx = torch.randint(0, n, (m,), device="cuda")
        out = torch.empty((m, 4), dtype=torch.float32, device="cuda")
        for _ in range(3):
            torch.index_select(table, 0, idx, out=out)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            torch.index_select(table, 0, idx, out=out)
        e.record(); torch.cuda.synchronize()
        gms = a.elapsed_time(e) / 10
        print("random 16-B gather calibration: %.1f us for %d accesses -> %.0f GB/s of 32-B sectors (+ %.0f GB/s of index reads and output writes)" % (
            gms * 1e3, m, 32 * m / gms / 1e6, (8 + 16) * m / gms / 1e6))
else:
    for v in sys.argv[1:] or ["0", "1", "2", "3", "4"]:
        env = dict(os.environ, LINS_JAC_VARIANT=v)
        subprocess.call([sys.executable, os.path.abspath(__file__), "child"], env=env)
