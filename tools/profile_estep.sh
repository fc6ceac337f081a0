#!/bin/bash
# usage: tools/profile_estep.sh <tag>   — ncu --set full capture of one E-step kernel pair (iteration-1 state) on the
# GPU box, then per-line attribution here against a listing built from the CURRENT source.
set -e
TAG=$1
cd /root/repo
/usr/local/graft/bin/gpurun --timeout 900 -- "timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bfgs_kernel|post_group' -s 2 -c 2 -o gpurun_out/prof_$TAG python tools/gpu_perf.py --iters 2 > gpurun_out/ncu_$TAG.log 2>&1; tail -1 gpurun_out/ncu_$TAG.log" | tail -2
cd gpurun_out
ncu -i prof_$TAG.ncu-rep --page source --csv > src_${TAG}_all.csv 2>/dev/null
# ncu prints the two kernels one after another: one csv per kernel
python -c "
t = open('src_${TAG}_all.csv').read().split('\"Kernel Name\"')
open('src_${TAG}_bfgs.csv', 'w').write('\"Kernel Name\"' + t[1])
open('src_${TAG}_post.csv', 'w').write('\"Kernel Name\"' + t[3])
"
mkdir -p sass
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -DSTM_KPL=2 -cubin ../strutopy_b200/csrc/estep_inst.cu -o sass/kpl2_$TAG.cubin 2>/dev/null
nvdisasm -gi -c sass/kpl2_$TAG.cubin > sass/kpl2_$TAG.sass
ncu -i prof_$TAG.ncu-rep --page details 2>/dev/null | grep -E "bfgs_kernel|post_group_kernel|Duration|Executed Ipc Active|Registers Per|Warp Cycles Per Issued|Theoretical Active Warps"
python ../tools/ncu_hot.py src_${TAG}_bfgs.csv sass/kpl2_$TAG.sass _ZN3stm11bfgs_kernelILi2ELi5E 30
python ../tools/ncu_hot.py src_${TAG}_post.csv sass/kpl2_$TAG.sass _ZN3stm17post_group_kernelILi2E 20
