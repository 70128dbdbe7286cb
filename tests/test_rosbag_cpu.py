"""CPU suite: the dependency-free ROS1 bag v2.0 reader / writer (csrc/host/rosbag_reader.hpp).

The C++ implementation (product) and tools/bag_tool.py (pure Python, written separately from the published format
description) check each other: Python writes -> C++ reads, C++ writes -> Python reads, on sensor_msgs/PointCloud2,
sensor_msgs/Imu and cloud_msgs/cloud_info — the three messages the LINS front end consumes
(lins/src/lib/Estimator.cpp:36-45, lins/src/image_projection_node.cpp:83, cloud_msgs/msg/cloud_info.msg:1-12).
tests/golden/tiny.bag is the committed fixture (tools/bag_tool.py make-fixture).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bag_tool  # noqa: E402

FIX = os.path.join(ROOT, "tests", "golden", "tiny.bag")


@pytest.fixture(scope="module")
def baglib(defs):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools", "synth"), "liblins_bag.so"])
    L = C.CDLL(os.path.join(ROOT, "tools", "synth", "liblins_bag.so"))
    vp = C.c_void_p
    L.lins_bag_summary.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.lins_bag_read_imu.argtypes = [C.c_char_p, C.c_char_p, vp, C.c_int, C.POINTER(C.c_int)]
    L.lins_bag_read_cloud.argtypes = [C.c_char_p, C.c_char_p, C.c_int, vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.lins_bag_read_cloud_info.argtypes = [C.c_char_p, C.c_char_p, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.lins_bag_write_test.argtypes = [C.c_char_p] * 4 + [C.c_int, C.c_int]
    return L


def _summary(L, path):
    buf = C.create_string_buffer(1 << 16)
    rc = L.lins_bag_summary(path.encode(), buf, len(buf))
    return rc, buf.value.decode()


def test_committed_fixture_is_reproducible(tmp_path):
    p = tmp_path / "again.bag"
    conns, msgs, _ = bag_tool.fixture_contents()
    bag_tool.write_bag(str(p), conns, msgs)
    assert p.read_bytes() == open(FIX, "rb").read(), "tests/golden/tiny.bag is not what tools/bag_tool.py make-fixture writes"


def test_cpp_reader_decodes_the_python_written_fixture(baglib, defs):
    conns, msgs, truth = bag_tool.fixture_contents()
    T = bag_tool.FIXTURE_TOPICS
    rc, text = _summary(baglib, FIX)
    assert rc == 0, text
    rows = {ln.split()[0]: ln.split() for ln in text.strip().splitlines()}
    assert rows[T["lidar"]][1] == "sensor_msgs/PointCloud2" and int(rows[T["lidar"]][3]) == 3
    assert rows[T["imu"]][1] == "sensor_msgs/Imu" and int(rows[T["imu"]][3]) == 30
    assert rows[T["info"]][1] == "cloud_msgs/cloud_info" and int(rows[T["info"]][3]) == 3
    # Imu: stamp, acc, gyr bit-exact
    out = np.zeros((64, 8))
    n = C.c_int(0)
    assert baglib.lins_bag_read_imu(FIX.encode(), T["imu"].encode(), defs.ptr(out), 64, C.byref(n)) == 0 and n.value == 30
    want = np.array(truth["imu"])
    assert np.array_equal(out[:30, 1:7], want[:, 1:7]) and np.allclose(out[:30, 0], want[:, 0], atol=1e-9) and np.allclose(out[:30, 7], want[:, 0], atol=1e-9)
    # PointCloud2 (velodyne PointXYZIR, 32-byte step) -> PointXYZI; the ring field is ignored like pcl::fromROSMsg does
    for k, (stamp, xyz, inten) in enumerate(truth["clouds"]):
        pts = np.zeros(4096, defs.POINT_DTYPE)
        npts, st = C.c_int(0), C.c_double(0)
        assert baglib.lins_bag_read_cloud(FIX.encode(), T["lidar"].encode(), k, defs.ptr(pts), len(pts), C.byref(npts), C.byref(st)) == 0
        assert npts.value == len(xyz) and abs(st.value - stamp) < 1e-9
        p = pts[: npts.value]
        assert np.array_equal(np.stack([p["x"], p["y"], p["z"]], 1), xyz) and np.array_equal(p["intensity"], inten)
        assert (p["pad0"] == 1.0).all()  # PCL_ADD_POINT4D
    # cloud_info
    for k, (stamp, info) in enumerate(truth["info"]):
        sr, er = np.zeros(32, np.int32), np.zeros(32, np.int32)
        ori, gr, col, rg = np.zeros(3, np.float32), np.zeros(256, np.uint8), np.zeros(256, np.uint32), np.zeros(256, np.float32)
        nr, npt, st = C.c_int(0), C.c_int(0), C.c_double(0)
        assert baglib.lins_bag_read_cloud_info(FIX.encode(), T["info"].encode(), k, defs.ptr(sr), defs.ptr(er), 32, defs.ptr(ori), defs.ptr(gr), defs.ptr(col),
                                               defs.ptr(rg), 256, C.byref(nr), C.byref(npt), C.byref(st)) == 0
        assert nr.value == 16 and npt.value == len(info["rng"]) and abs(st.value - stamp) < 1e-9
        assert np.array_equal(sr[:16], info["start"]) and np.array_equal(er[:16], info["end"]) and np.allclose(ori, info["ori"])
        n_ = npt.value
        assert np.array_equal(gr[:n_], info["ground"]) and np.array_equal(col[:n_], info["col"]) and np.array_equal(rg[:n_], info["rng"])
    # a missing topic / message index is an error, not garbage
    assert baglib.lins_bag_read_cloud(FIX.encode(), b"/nope", 0, defs.ptr(pts), len(pts), C.byref(npts), C.byref(st)) != 0
    assert baglib.lins_bag_read_cloud(FIX.encode(), T["lidar"].encode(), 7, defs.ptr(pts), len(pts), C.byref(npts), C.byref(st)) != 0


def test_python_reader_decodes_what_the_cpp_writer_wrote(baglib, tmp_path):
    p = str(tmp_path / "cpp.bag")
    assert baglib.lins_bag_write_test(p.encode(), b"/velodyne_points", b"/imu/data", b"/segmented_cloud_info", 4, 500) == 0
    conns, msgs = bag_tool.read_bag(p)
    by_topic = {}
    for cid, t, d in msgs:
        by_topic.setdefault(conns[cid]["topic"], []).append((t, d))
    assert {c["type"] for c in conns.values()} == {"sensor_msgs/PointCloud2", "sensor_msgs/Imu", "cloud_msgs/cloud_info"}
    assert len(by_topic["/velodyne_points"]) == 4 and len(by_topic["/imu/data"]) == 40 and len(by_topic["/segmented_cloud_info"]) == 4
    for k, (t, d) in enumerate(by_topic["/velodyne_points"]):
        c = bag_tool.decode_pointcloud2(d)
        i = np.arange(500)
        assert abs(t - (100.0 + 0.1 * k + 0.1)) < 1e-6 and abs(c["header"]["stamp"] - t) < 1e-6
        assert np.array_equal(c["x"], (k + np.float32(0.001) * i.astype(np.float32)).astype(np.float32))
        assert np.array_equal(c["intensity"], ((i % 16).astype(np.float32) + np.float32(0.01) * np.float32(k)).astype(np.float32))
    im = bag_tool.decode_imu(by_topic["/imu/data"][13][1])  # scan 1, sample 3
    assert np.allclose(im["linear_acceleration"], [0.3, -0.2, 9.81]) and np.allclose(im["angular_velocity"], [0.003, 0.002, -0.003])
    ci = bag_tool.decode_cloud_info(by_topic["/segmented_cloud_info"][2][1])
    assert np.array_equal(ci["startRingIndex"], np.arange(16) * 7 + 2) and len(ci["range"]) == 500 and ci["ground"][3] == 1
    # and the C++ reader reads its own writer (index + chunk-info sections are skipped, several chunks are walked)
    rc, text = _summary(baglib, p)
    assert rc == 0 and "/imu/data sensor_msgs/Imu" in text


def test_compressed_and_damaged_bags(baglib, tmp_path):
    conns, msgs, _ = bag_tool.fixture_contents()
    z = str(tmp_path / "z.bag")
    bag_tool.write_bag(z, conns, msgs, compression="bz2")
    rc, text = _summary(baglib, z)
    assert rc == -3 and "bz2" in text  # LINS_BAG_E_COMPRESSED: reported, not mis-parsed
    # tools/bag_tool.py decompress makes it readable
    u = str(tmp_path / "u.bag")
    assert bag_tool.main(["bag_tool", "decompress", z, u]) == 0
    rc, text = _summary(baglib, u)
    assert rc == 0 and text.count("\n") == 3
    # truncated file: the complete records before the cut are still delivered; garbage is rejected
    raw = open(FIX, "rb").read()
    t = tmp_path / "t.bag"
    t.write_bytes(raw[: len(raw) // 2])
    rc, _ = _summary(baglib, str(t))
    assert rc in (0, -2)
    g = tmp_path / "g.bag"
    g.write_bytes(b"not a bag at all")
    assert _summary(baglib, str(g))[0] == -2
    assert _summary(baglib, str(tmp_path / "missing.bag"))[0] == -1


# ---- "lz4" chunk compression (one LZ4 frame per chunk, ros_comm's roslz4) ---------------------------------------------------
def _liblz4():
    import ctypes.util
    name = C.util.find_library("lz4")
    if not name:
        pytest.skip("no system liblz4 to pin the codec against")
    L = C.CDLL(name)
    L.LZ4F_compressFrameBound.restype = C.c_size_t
    L.LZ4F_compressFrameBound.argtypes = [C.c_size_t, C.c_void_p]
    L.LZ4F_compressFrame.restype = C.c_size_t
    L.LZ4F_compressFrame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    L.LZ4F_isError.argtypes = [C.c_size_t]
    L.LZ4F_createDecompressionContext.argtypes = [C.POINTER(C.c_void_p), C.c_uint]
    L.LZ4F_createDecompressionContext.restype = C.c_size_t
    L.LZ4F_decompress.restype = C.c_size_t
    L.LZ4F_decompress.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p]
    L.LZ4F_freeDecompressionContext.argtypes = [C.c_void_p]
    return L


def _payloads():
    rng = np.random.default_rng(2)
    conns, msgs, _ = bag_tool.fixture_contents()
    return [b"", b"a", b"abcd" * 3, b"x" * 13, b"\x00" * 70000, b"hello world, " * 9000,
            bytes(rng.integers(0, 256, 100000, dtype=np.uint8)),  # incompressible: stored raw
            bytes(rng.integers(0, 4, 300000, dtype=np.uint8)),    # several 64-KiB blocks, long literal / match lengths
            b"".join(m for _, _, m in msgs)]                      # real message bytes


def test_lz4_codec_round_trips_and_matches_the_system_liblz4():
    for data in _payloads():
        frame = bag_tool.lz4_frame_compress(data)
        assert bag_tool.lz4_frame_decompress(frame, len(data)) == data
    L = _liblz4()
    for data in _payloads():
        # the real library's frames (linked blocks by default) through our decoder
        cap = L.LZ4F_compressFrameBound(len(data), None)
        dst = C.create_string_buffer(cap)
        n = L.LZ4F_compressFrame(dst, cap, data, len(data), None)
        assert not L.LZ4F_isError(n)
        assert bag_tool.lz4_frame_decompress(dst.raw[:n], len(data)) == data
        # our frames (header checksum = our XXH32) through the real library's decoder
        frame = bag_tool.lz4_frame_compress(data)
        ctx = C.c_void_p()
        assert not L.LZ4F_isError(L.LZ4F_createDecompressionContext(C.byref(ctx), 100))
        out = C.create_string_buffer(max(1, len(data)))
        got, pos = b"", 0
        while pos < len(frame):
            dn, sn = C.c_size_t(len(out)), C.c_size_t(len(frame) - pos)
            src = (C.c_char * sn.value).from_buffer_copy(frame[pos:])
            rc = L.LZ4F_decompress(ctx, out, C.byref(dn), src, C.byref(sn), None)
            assert not L.LZ4F_isError(rc)
            got += out.raw[: dn.value]
            pos += sn.value
            if rc == 0:
                break
        L.LZ4F_freeDecompressionContext(ctx)
        assert got == data and pos == len(frame)


def test_cpp_reader_inflates_lz4_chunks(baglib, tmp_path):
    conns, msgs, _ = bag_tool.fixture_contents()
    z = str(tmp_path / "lz4.bag")
    bag_tool.write_bag(z, conns, msgs, compression="lz4", chunk_bytes=1 << 14)  # several chunks
    assert os.path.getsize(z) < os.path.getsize(FIX)
    ref_rc, ref_text = _summary(baglib, FIX)
    rc, text = _summary(baglib, z)
    assert rc == 0 and ref_rc == 0 and text == ref_text
    # every IMU sample and every cloud byte identical to the uncompressed fixture
    for path_a, path_b in [(FIX, z)]:
        outs = []
        for p in (path_a, path_b):
            imu = np.zeros((4096, 11))
            n = C.c_int(0)
            assert baglib.lins_bag_read_imu(p.encode(), b"/imu/data", imu.ctypes.data, 4096, C.byref(n)) == 0
            pts = np.zeros((1 << 16, 8), dtype=np.float32)
            m, st = C.c_int(0), C.c_double(0)
            assert baglib.lins_bag_read_cloud(p.encode(), b"/velodyne_points", 0, pts.ctypes.data, 1 << 16, C.byref(m), C.byref(st)) == 0
            outs.append((imu[: n.value].copy(), pts[: m.value].copy(), st.value))
        assert outs[0][0].shape == outs[1][0].shape and np.array_equal(outs[0][0], outs[1][0])
        assert outs[0][1].shape[0] > 0 and np.array_equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]
    # chunks compressed by the system liblz4 (linked blocks, its own match finder) read the same
    import ctypes.util
    if C.util.find_library("lz4"):
        L = _liblz4()

        def real(data):
            cap = L.LZ4F_compressFrameBound(len(data), None)
            dst = C.create_string_buffer(cap)
            n = L.LZ4F_compressFrame(dst, cap, data, len(data), None)
            assert not L.LZ4F_isError(n)
            return dst.raw[:n]

        zr = str(tmp_path / "lz4_real.bag")
        bag_tool.write_bag(zr, conns, msgs, compression="lz4", chunk_bytes=1 << 17, lz4_compress=real)
        rc, text = _summary(baglib, zr)
        assert rc == 0 and text == ref_text
        pts = np.zeros((1 << 16, 8), dtype=np.float32)
        m, st = C.c_int(0), C.c_double(0)
        assert baglib.lins_bag_read_cloud(zr.encode(), b"/velodyne_points", 0, pts.ctypes.data, 1 << 16, C.byref(m), C.byref(st)) == 0
        assert np.array_equal(pts[: m.value], outs[0][1])
    # the Python reader reads it back too, and a corrupted frame is rejected, not mis-parsed
    c2, m2 = bag_tool.read_bag(z)
    assert [x[2] for x in m2] == [x[2] for x in msgs]
    raw = bytearray(open(z, "rb").read())
    k = raw.find(b"\x04\x22\x4d\x18")
    assert k > 0
    bad = tmp_path / "bad.bag"
    raw[k + 40 : k + 60] = b"\xff" * 20
    bad.write_bytes(bytes(raw))
    rc, text = _summary(baglib, str(bad))
    assert rc == -2 and "lz4" in text
