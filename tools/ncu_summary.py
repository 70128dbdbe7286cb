"""Turn an .ncu-rep capture (one kernel, --set full) into the summary CSV kept under profiles/ and record the
kernel's measured DRAM traffic per launch in profiles/traffic.json (read by bench.py for roofline.traffic).

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_x_ncu_summary.csv fused|jacobian
"""
import csv, io, json, os, re, subprocess, sys

KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|gpu__time_duration|dram__bytes|dram__throughput|sm__throughput|launch__|"
                  r"sm__warps_active|smsp__issue_active|smsp__inst_executed\.(sum|avg\.per_cycle_active)|l1tex__t_sector_hit_rate|"
                  r"lts__t_sector_hit_rate|lts__t_bytes\.sum|l1tex__t_bytes\.sum|smsp__cycles_active\.avg|sm__cycles_elapsed|"
                  r"smsp__average_warp.*|sm__inst_executed_pipe_(fp64|fma|alu|lsu|xu).*sum$|sm__pipe_fp64_cycles_active.avg.pct|"
                  r"smsp__pcsamp_warps_issue_stalled)")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(rep, out_csv, key):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    traffic = 0.0
    inst = dur_ms = None
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for h, u, v in zip(hdr, units, vals):
            if KEEP.match(h):
                w.writerow([h, u, v])
            if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                traffic += float(v) * SCALE[u]
            if h == "smsp__inst_executed.sum":
                inst = float(v.replace(",", ""))
            if h == "gpu__time_duration.sum":
                dur_ms = float(v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
    tj = os.path.join(os.path.dirname(out_csv), "traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    d[key] = {"dram_bytes_per_launch": traffic, "source": os.path.basename(out_csv)}
    if inst is not None:
        d[key]["warp_instructions_per_launch"] = inst  # smsp__inst_executed.sum: bench.py's issue roofline
    if dur_ms is not None:
        d[key]["duration_ms_under_ncu"] = dur_ms
    json.dump(d, open(tj, "w"), indent=1, sort_keys=True)
    print(key, "traffic per launch: %.1f MB" % (traffic / 1e6), "->", out_csv)


if __name__ == "__main__":
    main(*sys.argv[1:4])
