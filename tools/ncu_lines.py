"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line (diagnostics).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_lines.py x.csv [top]"""
import csv, sys
csv.field_size_limit(10**9)
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; last = None
agg = {}  # (file, line) -> [samples, not-issued samples, warp instructions, source text, sass count]
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0] != "":
        try:
            key = (cur, int(r[0]))
            v = agg.setdefault(key, [0, 0, 0, r[1].strip()[:100], 0])
            v[0] += int(r[4]); v[1] += int(r[5]); v[2] += int(r[7])
            last = key
        except ValueError:
            last = None
    elif last is not None:
        agg[last][4] += 1
tot = sum(v[0] for v in agg.values()) or 1; toti = sum(v[2] for v in agg.values()) or 1; tots = sum(v[4] for v in agg.values())
print("total samples", tot, "warp instructions", toti, "sass instructions", tots)
print("--- by samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:22s}:{k[1]:4d} samp {v[0]:6d} ({100*v[0]/tot:4.1f}%) instr {100*v[2]/toti:4.1f}% sass {v[4]:4d} | {v[3]}")
print("--- by instructions")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{k[0]:22s}:{k[1]:4d} samp {100*v[0]/tot:4.1f}% instr {v[2]:10d} ({100*v[2]/toti:4.1f}%) sass {v[4]:4d} | {v[3]}")
print("--- by file")
byf = {}
for k, v in agg.items():
    f = byf.setdefault(k[0], [0, 0, 0]); f[0] += v[0]; f[1] += v[2]; f[2] += v[4]
for f, v in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print(f"{f:24s} samp {100*v[0]/tot:5.1f}% instr {100*v[1]/toti:5.1f}% sass {v[2]}")
