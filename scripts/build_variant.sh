#!/bin/bash
# usage: scripts/build_variant.sh NAME [-DFLAG ...]  -> amuya_sfm_lines_beta_b200/variants/liblba_NAME.so
cd /root/repo/amuya_sfm_lines_beta_b200/csrc
name=$1; shift
mkdir -p ../variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -shared -o ../variants/liblba_$name.so lba.cu -ldl
