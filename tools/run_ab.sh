cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = main ]; then unset LINS_GPU_LIB; else export LINS_GPU_LIB=$GRAFT_REPO_ROOT/variants/liblins_gpu_$v.so; fi
  echo "=== variant $v" >> gpurun_out/ab.log
  (timeout 150 python -u tools/phase_profile.py) >> gpurun_out/ab.log 2>&1
  (timeout 200 python -u bench.py --steps 10 --warmup 3 --no-cpu-baseline | grep -o '"ms_per_step": [0-9.]*') >> gpurun_out/ab.log 2>&1
done
cat gpurun_out/ab.log
