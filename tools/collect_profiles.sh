#!/bin/bash
# Copy the outputs of one tools/gpu_round.sh pass (gpurun_out/<tag>_*) into profiles/<round>_* (tracked, what DESIGN.md cites).
# Usage: bash tools/collect_profiles.sh <tag in gpurun_out> <round prefix, e.g. r02>
TAG=$1; R=$2; G=gpurun_out; P=profiles
cp $G/${TAG}_bench.json $P/${R}_bench.json
cp $G/${TAG}_bench_ref.json $P/${R}_bench_ref.json
cp $G/${TAG}_pytest_gpu.log $P/${R}_pytest_gpu.log
cp $G/${TAG}_recon_profile.txt $P/${R}_recon_step_kernels.txt
cp $G/${TAG}_gemm_check.txt $P/${R}_gemm_check.txt
cp $G/${TAG}_agg_bench.txt $P/${R}_gcn_aggregate_bench.txt
cp $G/${TAG}_chamfer_sweep.txt $P/${R}_chamfer_sweep.txt
cp $G/${TAG}_vertex_front.txt $P/${R}_vertex_front.txt
cp $G/${TAG}_torch_gpu_baselines.txt $P/${R}_torch_gpu_baselines.txt
python tools/launch_summary.py $G/${TAG}_launches.csv "python bench.py --steps 3 --warmup 3 --no-cpu --no-extra" > $P/${R}_ncu_launches.txt
python tools/ncu_summary.py $G/${TAG}_chamfer.ncu-rep > $P/${R}_ncu_chamfer_filter.txt
python tools/ncu_summary.py $G/${TAG}_agg.ncu-rep > $P/${R}_ncu_gcn_aggregate_tile.txt
python tools/ncu_summary.py $G/${TAG}_fwd.ncu-rep > $P/${R}_ncu_sgemm_fwd_tma.txt
python - <<PY
import csv, io, json, subprocess
out = subprocess.run(["ncu", "-i", "$G/${TAG}_chamfer.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, units, r = rows[0], rows[1], rows[2]
def val(name):
    i = hdr.index(name); v = float(r[i].replace(",", "")); u = units[i].lower()
    return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)
json.dump({"kernel": "chamfer_nn_filter_tma_kernel", "pairs": 256, "points": 10000,
           "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
           "source": "profiles/${R}_ncu_chamfer_filter.txt (ncu --set full, dram__bytes_read.sum / dram__bytes_write.sum, one launch)"},
          open("$P/chamfer_scan_traffic.json", "w"), indent=1)
PY
ls $P/${R}_*
cp $G/${TAG}_pruned_check.txt $P/${R}_chamfer_pruned_check.txt
cp $G/${TAG}_pruned_kernels.txt $P/${R}_chamfer_pruned_kernels.txt
python tools/ncu_summary.py $G/${TAG}_pruned.ncu-rep > $P/${R}_ncu_chamfer_pruned.txt
ls $P/${R}_*pruned*
cp $G/${TAG}_fwd_tail_probe.txt $P/${R}_fwd_tail_probe.txt
