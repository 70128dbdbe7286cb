cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01_map2.csv python tools/map_bench.py > gpurun_out/map_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lins_map_knn -s 3 -c 1 -o gpurun_out/prof_r01_map2_knn python tools/map_bench.py > gpurun_out/prof_map1.log 2>&1
python tools/map_bench.py
