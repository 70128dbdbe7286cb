cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 150 python -u tools/phase_profile.py) > gpurun_out/pp.log 2>&1
tail -n 10 gpurun_out/pp.log
