"""Per-phase cycle breakdown of the fused kernel (diagnostics; run on a GPU box)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
capi = importlib.import_module("lins---lidar-inertial-slam_b200.capi")
synth = importlib.import_module("lins---lidar-inertial-slam_b200.synth")
defs = importlib.import_module("lins---lidar-inertial-slam_b200.ctypes_defs")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
b = synth.generate("config3", n=n, seed0=1000)
g = capi.LinsGpu(defs.LinsParams.shipped())
g.batch_upload(b); g.batch_run(); g.sync()
g.phase_cycles(enable=True)
g.batch_run(); g.sync()
t = g.phase_cycles(enable=False, read=True)
_, _, res, _ = g.batch_download(states=False, covs=False)
its = int(res["iters"].sum())
names = ["setup(prior,ring tables,index build)", "(loop top)", "tile stage (TMA)", "de-skew + certificates (P1)", "closest-point searches (P2)", "walk windows + walks (P3+P4)",
         "residual+fold", "block reduce", "tail's closing barrier", "exit cov+outputs"]
nq_avg = float(np.mean(np.diff(b.offsets["surf_flat"]) + np.diff(b.offsets["corner_sharp"])))  # queries per scan = searches of a first pass
tot = t[:10].sum() + t[18] + t[19] + t[21] + t[23] + t[25]
print(f"scans {n} iterations {its}  total CTA-cycles {tot:.3e}  per iteration {tot/its:.0f} cycles")
for k, nm in enumerate(names):
    per = t[k] / (n if k in (0, 9) else its)
    print(f"  {nm:40s} {100*t[k]/tot:5.1f}%   {per:9.0f} cycles per {'scan' if k in (0,9) else 'iteration'}")
print("serial tail per iteration: sums+A6 %.0f | M6+LU %.0f | K x, norms %.0f | logic+boxPlus %.0f | next consts+boxMinus %.0f | barrier %.0f" % (t[18]/its, t[19]/its, t[21]/its, t[23]/its, t[25]/its, t[8]/its))
print("first iteration %.0f cycles/scan | CTA busy: mean %.3e max %.3e (tail %.1f%%)" % (t[24]/n, t[26]/148, t[27], 100*(1-t[26]/148/max(t[27],1))))
print("first pass of a scan: P2 %.0f, P3+P4 %.0f cycles/scan | later passes: P2 %.0f, P3+P4 %.0f cycles/iteration" % (
    t[28]/n, t[29]/n, (t[4]-t[28])/max(its-n,1), (t[5]-t[29])/max(its-n,1)))
print("(the next three lines need a library built with -DLINS_SEARCH_DIAG=1; zeros otherwise)")
print("later passes, per search (lane-0 clocks): closest-point mean %.0f max %.0f cycles; searches with >= 64 bins: %.2f per iteration, mean %.0f cycles | walks mean %.0f cycles, wide (>= 64 bins) share of walk time %.1f%%" % (
    t[20]/max(t[10]-n*nq_avg,1), t[22], t[17]/max(its-n,1), t[16]/max(t[17],1), t[30]/max(t[11]-n*nq_avg,1), 100*t[31]/max(t[30],1)))
c = max(t[36], 1)
print("later-pass closest-point search, mean cycles: fetch+loads %.0f | setup+scan loop %.0f | arg-min %.0f | epilogue %.0f" % (t[32]/c, t[33]/c, t[34]/c, t[35]/c))
print("later-pass closest-point search: candidates of the busiest lane %.1f, of all lanes %.1f" % (t[37]/c, t[38]/c))
print("work lists per iteration: closest-point searches %.1f (ring-bins per search %.1f; first pass %.1f), walk searches %.1f (ring-bins per search %.1f; first pass %.1f)" % (
    t[10]/its, t[12]/max(t[10],1), t[14]/max(n,1)/nq_avg, t[11]/its, t[13]/max(t[11],1), t[15]/max(n,1)/nq_avg))
