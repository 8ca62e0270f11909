#!/bin/bash
# usage: build_variant.sh name -Dflags...
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -o /root/repo/benchmarks/probes/lib_$name.so /root/repo/gxbeam_b200/csrc/gxb_lib.cu /root/repo/gxbeam_b200/csrc/gxb_section.cu && echo built $name
