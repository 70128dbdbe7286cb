cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
md5sum lins---lidar-inertial-slam_b200/liblins_gpu.so | cut -c1-8 > gpurun_out/build_id.txt
(time timeout 420 python -u -m pytest tests -m gpu -x -q) > gpurun_out/t1.log 2>&1
(timeout 100 python -u tools/map_bench.py) > gpurun_out/map.log 2>&1
(time timeout 280 python -u bench.py --steps 12 --warmup 3 --no-cpu-baseline) > gpurun_out/bench.log 2>&1
cat gpurun_out/build_id.txt; tail -n 4 gpurun_out/t1.log; cat gpurun_out/map.log; grep -o '"value": [0-9.]*, "unit": "iterations/s", "n_gpus\|"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"ms_per_call": [0-9.]*' gpurun_out/bench.log
