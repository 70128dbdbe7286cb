cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 420 python -u -m pytest tests -m gpu -x -v --timeout 150 --timeout-method thread -k "not shim") > gpurun_out/t1.log 2>&1
(time LINS_SEQ_VERBOSE=1 timeout 150 python -u -m pytest tests -m gpu -x -v -s -k shim) > gpurun_out/t2.log 2>&1
(time timeout 150 python -u tools/phase_profile.py) > gpurun_out/pp.log 2>&1
(time timeout 280 python -u bench.py --steps 10 --warmup 3) > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/t1.log gpurun_out/t2.log; tail -12 gpurun_out/pp.log; tail -2 gpurun_out/bench.log | cut -c1-400
