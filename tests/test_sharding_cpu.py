"""N > 1 host-side logic on CPU: world_size-2 gloo run of the scan-sharded update + pose gather."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT, pkg


def test_shard_ranges_cover_everything():
    sh = pkg("sharding")
    for n in (0, 1, 7, 8, 1000, 1001):
        for w in (1, 2, 3, 8):
            r = [sh.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_gloo_pose_gather(tmp_path, ob, golden_batch):
    """Two processes, gloo: each runs its share (here through the CPU oracle as the stand-in processor), the
    gathered records must equal a single-process run over the whole batch."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, importlib
        sys.path.insert(0, {ROOT!r})
        import numpy as np
        import torch.distributed as dist
        from oracle import oracle_binding as ob
        defs = importlib.import_module("lins---lidar-inertial-slam_b200.ctypes_defs")
        sh = importlib.import_module("lins---lidar-inertial-slam_b200.sharding")
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        b = defs.Batch.load(os.path.join({ROOT!r}, "tests", "golden", "units_inputs.npz")).tile(2).subset(range(7))
        prm = ob.LinsParams.shipped()
        def process(sub):
            return ob.ieskf_batch(prm, sub, threads=1, want_cov=False)[2]
        rec = sh.run_sharded(b, rank, world, process, dist)
        if rank == 0:
            np.save({str(tmp_path / 'gathered.npy')!r}, rec)
        dist.barrier()
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", "29571", str(script)], env=env, timeout=300)
    rec = np.load(tmp_path / "gathered.npy")
    b = golden_batch.tile(2).subset(range(7))
    _, _, ref, _, _ = ob.ieskf_batch(ob.LinsParams.shipped(), b, threads=1, want_cov=False)
    assert np.array_equal(rec["scan_id"], np.arange(7))
    assert np.array_equal(rec["iters"], ref["iters"]) and np.array_equal(rec["flags"], ref["flags"])
    assert np.array_equal(rec["pose"], ref["pose"])
