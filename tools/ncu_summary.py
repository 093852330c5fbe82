"""Extract the judged metrics from .ncu-rep files into a markdown table (run here, no GPU needed)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]
print("| report | kernel | " + " | ".join(k.split(".")[0].replace("__", " ") for k in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")].replace("void unnamed>::", "")[:60]
        vals = []
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals.append(f"{row[i]} {units[i]}".strip())
            else:
                vals.append("-")
        print(f"| {path.split('/')[-1]} | `{name}` | " + " | ".join(vals) + " |")
