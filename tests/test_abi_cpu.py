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
    assert C.sizeof(defs.LinsBatchDesc) == 8 + 10 * 8 + 8  # (+ point_format, padded)
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


def test_host_pool_and_map_host_code(tmp_path):
    """More of the product's host C++ (csrc/host): the persistent worker pool (every run executes fn on exactly n
    threads and returns only when all are done, across many generations and changing n) and the row-F2 host pieces
    (lins_cv_small.hpp / lins_map_host.hpp: cv::eigen / cv::solve restatements, the LM step) against OpenCV / numpy."""
    import subprocess

    host = os.path.join(ROOT, "lins---lidar-inertial-slam_b200", "csrc", "host")
    src = tmp_path / "p.cpp"
    src.write_text(r'''
#include <atomic>
#include <cstdio>
#include "host_pool.hpp"
#include "lins_map_host.hpp"
int main() {
  HostPool pool;
  long total = 0;
  for (int gen = 0; gen < 400; ++gen) {
    const int n = 1 + (gen * 7) % 13;
    std::atomic<int> calls(0), work(0);
    std::atomic<size_t> next(0);
    pool.run(n, [&] { ++calls; for (;;) { size_t i = next.fetch_add(1); if (i >= 1000) break; ++work; } });
    if (calls.load() != n || work.load() != 1000) { printf("BAD gen %d calls %d work %d\n", gen, calls.load(), work.load()); return 1; }
    total += work.load();
  }
  printf("pool %ld\n", total);
  // cv small kernels on a fixed input (compared with cv2 / numpy by the caller)
  float A[9] = {2.0f, 0.3f, -0.1f, 0.3f, 1.0f, 0.2f, -0.1f, 0.2f, 0.5f}, W[3], V[9];
  lins_cv::jacobi_eigen<3>(A, W, V);
  for (float w : W) printf("%.9g ", w);
  for (float v : V) printf("%.9g ", v);
  float Q[15] = {1.0f, 2.0f, 0.5f, 1.1f, 2.1f, 0.4f, 0.9f, 1.8f, 0.7f, 1.3f, 2.2f, 0.45f, 0.8f, 1.9f, 0.55f}, b[5] = {-1, -1, -1, -1, -1};
  const bool ok = lins_cv::qr_solve<5, 3>(Q, b);
  printf("%d %.9g %.9g %.9g ", (int)ok, b[0], b[1], b[2]);
  // one LM step on a well-conditioned system
  float AtA[36], AtB[6], T[6] = {0.01f, -0.02f, 0.03f, 1.f, 2.f, 3.f};
  for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) AtA[i * 6 + j] = (i == j ? 500.f + 10.f * i : 3.f / (1 + i + j)); AtB[i] = 0.5f * (i + 1); }
  lins::mapping::LmState st;
  float dR, dT;
  const bool conv = lins::mapping::lm_step(AtA, AtB, 0, T, st, dR, dT);
  printf("%d %d %.9g %.9g ", (int)conv, (int)st.isDegenerate, dR, dT);
  for (float t : T) printf("%.9g ", t);
  return 0;
}''')
    exe = tmp_path / "p"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", host, "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], timeout=120).decode().split()
    assert out[0] == "pool" and int(out[1]) == 400 * 1000
    v = np.array(out[2:], dtype=np.float64)
    A = np.array([[2.0, 0.3, -0.1], [0.3, 1.0, 0.2], [-0.1, 0.2, 0.5]], np.float32)
    W, V = v[:3], v[3:12].reshape(3, 3)
    assert np.allclose(W, np.linalg.eigvalsh(A.astype(np.float64))[::-1], rtol=1e-5)
    assert np.allclose(np.abs(V @ A @ V.T - np.diag(W)).max(), 0, atol=1e-5)
    Q = np.array([1.0, 2.0, 0.5, 1.1, 2.1, 0.4, 0.9, 1.8, 0.7, 1.3, 2.2, 0.45, 0.8, 1.9, 0.55], np.float32).reshape(5, 3)
    assert int(v[12]) == 1 and np.allclose(v[13:16], np.linalg.lstsq(Q.astype(np.float64), -np.ones(5), rcond=None)[0], rtol=2e-4)
    AtA = np.array([[500.0 + 10 * i if i == j else 3.0 / (1 + i + j) for j in range(6)] for i in range(6)])
    x = np.linalg.solve(AtA, 0.5 * np.arange(1, 7))
    T0 = np.array([0.01, -0.02, 0.03, 1, 2, 3])
    assert int(v[17]) == 0 and np.allclose(v[20:26], T0 + x, rtol=1e-5)
    try:  # where the real OpenCV is importable the f32 kernels are bit-identical to it
        import cv2
        _, w, vv = cv2.eigen(A)
        assert np.array_equal(w.ravel(), W.astype(np.float32)) and np.array_equal(vv, V.astype(np.float32))
        _, xq = cv2.solve(Q, -np.ones((5, 1), np.float32), flags=cv2.DECOMP_QR)
        assert np.array_equal(xq.ravel(), v[13:16].astype(np.float32))
    except ImportError:
        pass
