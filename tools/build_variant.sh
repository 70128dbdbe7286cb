#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...]: variants/liblins_gpu_NAME.so for A/B runs (LINS_GPU_LIB selects it)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
python tools/build_variant.py "$name" "$@"
