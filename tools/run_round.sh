# run_round.sh TAG [tests] : bench (+ reference arm), ncu launch list and one full capture of the fused kernel, named per round/tag
cd $GRAFT_REPO_ROOT
TAG=${1:-r02_a}
mkdir -p gpurun_out
md5sum lins---lidar-inertial-slam_b200/liblins_gpu.so | cut -c1-8 > gpurun_out/build_id.txt
if [ -n "$2" ]; then
(time timeout 600 python -u -m pytest tests -m gpu -x -q) > gpurun_out/t1.log 2>&1
(time timeout 120 python -u -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')") > gpurun_out/smoke.log 2>&1
fi
(time timeout 400 python -u bench.py --steps 12 --warmup 3) > gpurun_out/bench_$TAG.log 2>&1
if [ -z "$NOREF" ]; then (time timeout 300 python -u bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/bench_ref_$TAG.log 2>&1; fi
if [ -z "$NONCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lins_ieskf -s 4 -c 1 -o gpurun_out/prof_${TAG}_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_d1.log 2>&1
fi
cat gpurun_out/build_id.txt; tail -n 3 gpurun_out/t1.log; tail -n 3 gpurun_out/smoke.log; tail -n 5 gpurun_out/bench_$TAG.log | cut -c1-3000; tail -n 4 gpurun_out/bench_ref_$TAG.log | cut -c1-1500; ls gpurun_out | tail -n 8
