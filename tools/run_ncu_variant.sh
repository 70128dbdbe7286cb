# run_ncu_variant.sh OUTNAME "name[:ENV=VAL,...]": ncu --set full capture of one fused launch of a library variant
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
out=$1; spec=$2
v=${spec%%:*}; envs=""
if [ "$spec" != "$v" ]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
if [ "$v" = main ]; then lib=""; else lib="LINS_GPU_LIB=$GRAFT_REPO_ROOT/variants/liblins_gpu_$v.so"; fi
env $lib $envs timeout 600 ncu --set full --clock-control none --import-source on -k regex:lins_ieskf -s 3 -c 1 -f -o gpurun_out/$out python tools/quick_ab.py 1000 2 > gpurun_out/$out.log 2>&1
tail -n 3 gpurun_out/$out.log
