cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for ne in 6 3 9; do
  (LINS_E2E_CONTEXTS=$ne timeout 300 python -u bench.py --steps 12 --warmup 3 --no-cpu-baseline) > gpurun_out/bench_e$ne.log 2>&1
  echo "contexts $ne: $(grep -o '"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*' gpurun_out/bench_e$ne.log | tr '\n' ' ')"
done
