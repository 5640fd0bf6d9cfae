"""Dev tool: condense an .ncu-rep (ncu --set full) into the few lines DESIGN.md / bench.py cite.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_bytes\.sum|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_(fma|alu|lsu|fmaheavy|fmalite|uniform|xu)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_(fma|fmaheavy|fmalite|alu|tensor.*)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__(registers_per_thread|grid_size|block_size|occupancy_limit.*|waves_per_multiprocessor)|"
    r"smsp__inst_executed\.sum|sm__cycles_active\.avg|smsp__cycles_active\.avg|smsp__average_warp.*_per_issue_active.*|"
    r"smsp__inst_executed_op_.*\.sum|sm__sass_thread_inst_executed_op_f(add|mul|fma)_pred_on\.sum)$")


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# {path}: {len(data)} launch(es) profiled with ncu --set full --clock-control none")
    for r in data:
        print(f"\n## {r[name_i][:120]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            if KEEP.match(h):
                print(f"{h:95s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
