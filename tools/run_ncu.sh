cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
md5sum lins---lidar-inertial-slam_b200/liblins_gpu.so | cut -c1-8 > gpurun_out/build_id.txt
ncu --set full --clock-control none --import-source on -k regex:lins_ieskf -s 4 -c 1 -o gpurun_out/prof_r01_b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_b.log 2>&1
cat gpurun_out/build_id.txt; ls -la gpurun_out | tail -n 4
