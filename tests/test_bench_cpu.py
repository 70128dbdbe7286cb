"""bench.py host-side helpers (no GPU): the JSON contract's static parts, the reference arm's bounded sample, the NUMA
policy switch.  The timed legs themselves need a B200 and are exercised by the driver."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("lins_bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_thread_sweep_and_config():
    b = _bench()
    assert b.thread_sweep(128) == [32, 64, 128]
    assert b.thread_sweep(1) == [1]
    cfg = b.bench_config(1000, 1)
    assert "workload" in cfg and "config3" in cfg["workload"] and "model" not in cfg


def test_host_memory_policy_is_harmless():
    b = _bench()
    on = b.host_memory_policy(True)
    off = b.host_memory_policy(False)
    assert isinstance(on, str) and isinstance(off, str)  # one node / refused / interleaved: never raises


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ESKF iterations/sec" and d["value"] > 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
