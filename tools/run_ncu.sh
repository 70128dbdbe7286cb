cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lins_ieskf -s 4 -c 1 -o gpurun_out/prof_r01_tail python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_tail.log 2>&1
(time timeout 150 python -u tools/phase_profile.py) > gpurun_out/pp.log 2>&1
tail -n 14 gpurun_out/pp.log
