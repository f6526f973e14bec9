"""Key metrics of ncu --set full reports (.ncu-rep): one line block per report.
    python profiles/ncu_summary.py gpurun_out/a.ncu-rep [...]        # also writes <name>_raw.csv under profiles/ with --export
"""
import csv
import io
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
]
export = "--export" in sys.argv
for rep in [a for a in sys.argv[1:] if a.endswith(".ncu-rep")]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    if export:
        dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.path.basename(rep).replace(".ncu-rep", "_ncu_full_raw.csv"))
        open(dst, "w").write(out)
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print("== %s: %s" % (os.path.basename(rep), d.get("Kernel Name", "?")[:90]))
        for k in KEYS:
            if k in d and d[k] not in ("", None):
                print("   %-95s %s %s" % (k, d[k], u.get(k, "")))
