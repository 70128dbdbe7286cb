# GPU parity suite + quick A/B (+ phase profile with PP=1) over the number of resident units per CTA
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
if [ -z "$NOTEST" ]; then
(time timeout 600 python -u -m pytest tests -m gpu -x -q) > gpurun_out/t1.log 2>&1
tail -n 6 gpurun_out/t1.log
fi
for s in ${SLOTS:-1 2 3}; do
  echo "=== LINS_SLOTS=$s"
  LINS_SLOTS=$s LINS_VERBOSE=1 timeout 120 python -u tools/quick_ab.py 1000 12 2>&1 | sort | uniq -c | sort -rn | head -3
  if [ -n "$PP" ]; then LINS_SLOTS=$s timeout 120 python -u tools/phase_profile.py 2>&1 | tail -n 17; fi
done
