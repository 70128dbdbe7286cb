cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
(time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 12 --warmup 3) > gpurun_out/bench_n$N.log 2>&1
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1) > gpurun_out/bench_ref_n$N.log 2>&1
grep -o '"value": [0-9.]*, "unit": "iterations/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"ms_per_call": [0-9.]*\|"impl": "reference"' gpurun_out/bench_n$N.log gpurun_out/bench_ref_n$N.log; tail -n 3 gpurun_out/bench_n$N.log | cut -c1-200
