# run_ab2.sh "name[:ENV=VAL,ENV=VAL]" ... : quick_ab.py (+ optional phase profile) per library variant
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/ab2.log
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [ "$spec" != "$v" ]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  if [ "$v" = main ]; then lib=""; else lib="LINS_GPU_LIB=$GRAFT_REPO_ROOT/variants/liblins_gpu_$v.so"; fi
  echo "=== $spec" >> gpurun_out/ab2.log
  (env $lib $envs LINS_VERBOSE=1 timeout 120 python -u tools/quick_ab.py 1000 12 2>&1 | sort | uniq -c | sort -rn | head -4) >> gpurun_out/ab2.log 2>&1
  if [ -n "$AB_PP" ]; then (env $lib $envs timeout 120 python -u tools/phase_profile.py 2>&1 | head -20) >> gpurun_out/ab2.log 2>&1; fi
done
cat gpurun_out/ab2.log
