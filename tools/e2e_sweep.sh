# e2e_sweep.sh: bench.py's end-to-end figure for a few contexts:pack-threads[:i = interleaved host memory] settings on one box
cd $GRAFT_REPO_ROOT
for spec in ${E2E_SPECS:-3:21 4:16 4:12 6:10 3:16 2:32}; do
  n=$(echo $spec | cut -d: -f1); t=$(echo $spec | cut -d: -f2); il=$(echo $spec | cut -d: -f3)
  if [ "$il" = i ]; then il=1; else il=0; fi
  out=$(LINS_NUMA_INTERLEAVE=$il LINS_E2E_CONTEXTS=$n LINS_PACK_THREADS=$t timeout 300 python -u bench.py --steps 12 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$spec" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print("%s: device %.2f M  e2e %.2f M  e2e_packed16 %.2f M  (%s)" % (sys.argv[1], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["e2e_packed16"]["value"] / 1e6, d["e2e"]["numa"]))
PY
done
