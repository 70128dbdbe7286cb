#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...]: variants/liblins_gpu_NAME.so for A/B runs (LINS_GPU_LIB selects it)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o variants/liblins_gpu_$name.so "lins---lidar-inertial-slam_b200/csrc/cuda/lins_gpu.cu"
