#!/bin/bash
# A/B tuning builds of the headline kernel family (32 lanes x 4 nodes): compiles pr_ensemble_m4.cu with extra nvcc
# flags into a small shared object that the regular library loads in place of its own copy of that family:
#   tools/ab_build.sh <name> "<extra nvcc flags>"
#   PR_M4_VARIANT=flow_sim_b200/csrc/variants/m4_<name>.so python tools/run_headline.py ...
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../flow_sim_b200/csrc"
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr \
     -Xfatbin -compress-all -DPR_VARIANT_SHIM $flags -shared -cudart static -o variants/m4_$name.so pr_ensemble_m4.cu 2> variants/m4_$name.ptxas.log
grep -A2 "ILi32ELi4ELi16ELb0ELi1ELb1ELb0ELb0" variants/m4_$name.ptxas.log | tail -2
