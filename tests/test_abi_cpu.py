"""CPU suite, part 2: the C-ABI boundary and the host-side logic (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu, pkg


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lins_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lins_gpu_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(capi):
    capi.build()
    L = C.CDLL(capi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/lins_gpu.h but not exported"
    assert sorted(capi.EXPORTS) == names
    assert L.lins_gpu_abi_version() == 1


def test_struct_layouts_match_header(defs):
    assert C.sizeof(defs.LinsScanResult) == 64
    assert defs.POINT_DTYPE.itemsize == 32
    assert defs.POINT_DTYPE.fields["intensity"][1] == 16  # pcl::PointXYZI: intensity at byte 16
    assert C.sizeof(defs.LinsParams) == 48
    assert C.sizeof(defs.LinsReport) == 16 + 64 * 4 * 2 + 64 * 8 * 2
    assert C.sizeof(defs.LinsBatchDesc) == 8 + 10 * 8
    assert C.sizeof(defs.LinsMapReport) == 4 * 4 + 3 * 10 * 4  # row F2
    # the C side agrees (compile the header with gcc and print the sizes)
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "lins_gpu.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(lins_scan_result), sizeof(lins_params), sizeof(lins_report), sizeof(lins_batch_desc), sizeof(lins_map_report));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(t) for t in (defs.LinsScanResult, defs.LinsParams, defs.LinsReport, defs.LinsBatchDesc, defs.LinsMapReport)]


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(capi):
    """The product path must fail loudly without a device (never route through the oracle / a CPU path)."""
    with pytest.raises(capi.LinsError):
        capi.LinsGpu()


def test_create_rejects_bad_params(capi, defs):
    L = capi.lib()
    h = C.c_void_p()
    for kw in (dict(num_iter=65), dict(num_iter=-1), dict(icp_freq=0), dict(scan_period=0.0)):
        p = defs.LinsParams.shipped(**kw)
        assert L.lins_gpu_create(C.byref(p), 0, None, C.byref(h)) == -1
    assert L.lins_gpu_create(None, 0, None, C.byref(h)) == -1


def test_product_code_never_touches_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "lins---lidar-inertial-slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(#include|import|from)\b[^\n]*oracle", txt, flags=re.M):
                    bad.append(f)
    for f in ("tools/synth/lins_synth.cpp",):
        if re.search(r"^\s*#include[^\n]*oracle", open(os.path.join(ROOT, f)).read(), flags=re.M):
            bad.append(f)
    assert not bad, bad


def test_synth_is_deterministic_and_well_formed(synth, defs):
    a = synth.generate("config3", n=3, seed0=11, threads=1)
    b = synth.generate("config3", n=3, seed0=11, threads=3)
    for k in defs.Batch.FIELDS:
        assert np.array_equal(a.clouds[k].view(np.uint8), b.clouds[k].view(np.uint8))
        assert np.array_equal(a.offsets[k], b.offsets[k])
    assert np.array_equal(a.state, b.state) and np.array_equal(a.cov, b.cov)
    # reference feature-count limits (StateEstimator.hpp:749-758, :787-793): <= 2*6*16 sharp, <= 4*6*16 flat
    assert np.diff(a.offsets["corner_sharp"]).max() <= 192 and np.diff(a.offsets["surf_flat"]).max() <= 384
    # targets are ring-sorted (features are extracted ring by ring, :727)
    for i in range(a.n):
        u = a.unit(i)
        for k in ("surf_less_flat", "corner_less_sharp"):
            assert (np.diff(u[k]["intensity"].astype(int)) >= 0).all()
        P = u["cov"].reshape(18, 18)
        assert np.allclose(P, P.T) and np.linalg.eigvalsh(P).min() > -1e-12
        assert abs(np.linalg.norm(u["state"][6:10]) - 1) < 1e-12


def test_batch_roundtrip_and_subset(golden_batch, defs, tmp_path):
    p = str(tmp_path / "b.npz")
    golden_batch.save(p)
    b = defs.Batch.load(p)
    for k in defs.Batch.FIELDS:
        assert np.array_equal(b.clouds[k].view(np.uint8), golden_batch.clouds[k].view(np.uint8))
    s = golden_batch.subset([2, 0])
    assert s.n == 2 and np.array_equal(s.unit(1)["surf_flat"], golden_batch.unit(0)["surf_flat"])
    t = golden_batch.tile(3)
    assert t.n == 3 * golden_batch.n and np.array_equal(t.unit(golden_batch.n + 1)["state"], golden_batch.unit(1)["state"])


def test_host_small_linalg_and_filter(tmp_path):
    """The product's own host C++ (csrc/host): compile a tiny driver and check it against numpy."""
    import subprocess

    src = tmp_path / "t.cpp"
    host = os.path.join(ROOT, "lins---lidar-inertial-slam_b200", "csrc", "host")
    src.write_text(r'''
#include <cstdio>
#include "small_linalg.hpp"
#include "kalman_filter.hpp"
using namespace lins;
int main() {
  linalg::Mat<6> A; linalg::Vec<6> b;
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0 - 0.5; };
  linalg::Mat<6> B;
  for (auto& r : B) for (auto& e : r) e = rnd();
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double t = i == j ? 0.1 : 0; for (int k = 0; k < 6; ++k) t += B[i][k] * B[j][k]; A[i][j] = t; }
  for (auto& e : b) e = rnd();
  auto x = linalg::colPivQrSolve<6>(A, b);
  linalg::Vec<6> E; linalg::Mat<6> V, Vi;
  linalg::symmetricEigen<6>(A, E, V);
  linalg::inverse<6>(A, Vi);
  for (auto& r : A) for (auto e : r) printf("%.17g ", e);
  for (auto e : b) printf("%.17g ", e);
  for (auto e : x) printf("%.17g ", e);
  for (auto e : E) printf("%.17g ", e);
  for (auto& r : Vi) for (auto e : r) printf("%.17g ", e);
  // predictor: covariance stays symmetric PSD-ish and time advances
  filter::StatePredictor f;
  f.initialization(0.0, V3D(0,0,0), V3D(1,0,0), V3D(0,0,0), V3D(0,0,0), V3D(0,0,9.81), V3D(0,0,0.1));
  for (int k = 0; k < 40; ++k) f.predict(0.0025, V3D(0, 0, 9.81), V3D(0, 0, 0.1));
  printf("%.17g %.17g %.17g %.17g ", f.state_.rn_.x(), f.time_, f.covariance_(0, 0), f.covariance_(0, 3) - f.covariance_(3, 0));
  return 0;
}''')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", host, "-o", str(exe), str(src)])
    v = np.array(subprocess.check_output([str(exe)]).split(), dtype=float)
    A, b, x, E, Ai, rest = v[:36].reshape(6, 6), v[36:42], v[42:48], v[48:54], v[54:90].reshape(6, 6), v[90:]
    assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-9)
    assert np.allclose(E, np.linalg.eigvalsh(A), rtol=1e-10)
    assert np.allclose(Ai, np.linalg.inv(A), rtol=1e-8)
    assert abs(rest[0] - 0.1) < 1e-3 and abs(rest[1] - 0.1) < 1e-12 and rest[2] > 0 and rest[3] == 0
