#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list, ncu full captures of the dominant kernels, dev profiles.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc >> $OUT/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > $OUT/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chamfer_nn_filter -s 3 -c 1 -f -o $OUT/${TAG}_chamfer \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > $OUT/${TAG}_ncu_chamfer.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gcn_aggregate_tile -s 6 -c 1 -f -o $OUT/${TAG}_agg \
    python tools/agg_bench.py 256 2 > $OUT/${TAG}_ncu_agg.log 2>&1
timeout 600 python tools/recon_profile.py 16 > $OUT/${TAG}_recon_profile.txt 2>&1
timeout 600 python tools/gemm_check.py 31184,300,300 31184,448,300 > $OUT/${TAG}_gemm_check.txt 2>&1
timeout 600 python tools/agg_bench.py 256 20 > $OUT/${TAG}_agg_bench.txt 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_smoke.log | tail -2; cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 600 python tools/chamfer_sweep.py > $OUT/${TAG}_chamfer_sweep.txt 2>&1
timeout 300 python tools/vertex_front_bench.py > $OUT/${TAG}_vertex_front.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgemm_fwd_tma -s 3 -c 1 -f -o $OUT/${TAG}_fwd \
    python tools/fwd_one.py > $OUT/${TAG}_ncu_fwd.log 2>&1
timeout 300 python tests/torch_gpu_baselines.py > $OUT/${TAG}_torch_gpu_baselines.txt 2>&1
timeout 400 python tools/pruned_check.py > $OUT/${TAG}_pruned_check.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chamfer_pruned -s 6 -c 2 -f -o $OUT/${TAG}_pruned \
    python tools/pruned_profile.py pruned cube 256 10000 > $OUT/${TAG}_ncu_pruned.log 2>&1
for a in "pruned cube 256 10000" "pruned sphere 256 10000" "pruned sphere 16 50000" "pruned sphere 1 100000"; do
    timeout 200 python tools/pruned_profile.py $a 2>&1 | grep -v Warn | grep "us  \|pruned " >> $OUT/${TAG}_pruned_kernels.txt; done
(echo "# tail tiles on (default)"; timeout 200 python tools/fwd_tail_probe.py; echo "# PTK_FWD_TAIL=0: 64-row tiles everywhere"; PTK_FWD_TAIL=0 timeout 200 python tools/fwd_tail_probe.py) > $OUT/${TAG}_fwd_tail_probe.txt 2>&1
