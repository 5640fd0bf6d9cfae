#!/bin/bash
# ncu --set full captures of the kernels tools/gpu_round.sh does not cover: tcgen05 dgrad / wgrad, fused vertex front,
# dense-tile aggregate.  Usage (under gpurun): bash tools/ncu_extra.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:gemm_tf32x3_kernel -s 4 -c 1 -o $OUT/${TAG}_dgrad python tools/tc_gemm_probe.py > $OUT/${TAG}_ncu_dgrad.log 2>&1
timeout 300 $NCU -k regex:wgrad_tf32x3_kernel -s 4 -c 1 -o $OUT/${TAG}_wgrad python tools/tc_gemm_probe.py > $OUT/${TAG}_ncu_wgrad.log 2>&1
timeout 300 $NCU -k regex:vertex_front_fwd -s 4 -c 1 -o $OUT/${TAG}_front python tools/vertex_front_bench.py > $OUT/${TAG}_ncu_front.log 2>&1
timeout 300 $NCU -k regex:gcn_aggregate_union -s 4 -c 1 -o $OUT/${TAG}_dense python tools/agg_bench.py 256 3 > $OUT/${TAG}_ncu_dense.log 2>&1
ls -la $OUT/${TAG}_*.ncu-rep
