"""Summarise one `ncu --set full` report (.ncu-rep) into the JSON + text that profiles/ keeps:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name   ->  name.json, name.txt"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def to_bytes(u, v):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


txt, js = [], {}
for k in want:
    if k in col:
        u, v = col[k]
        txt.append("%-70s %-10s %s" % (k, u, v))
        js[k] = {"unit": u, "value": v}
if "dram__bytes_read.sum" in col:
    js["dram_bytes_per_launch"] = to_bytes(*col["dram__bytes_read.sum"]) + to_bytes(*col["dram__bytes_write.sum"])
    txt.append("dram bytes per launch (read + write): %.3e" % js["dram_bytes_per_launch"])
for short, key in (("tensor_pipe_pct_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   ("tensor_pipe_pct_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")):
    if key in col:
        js[short] = float(col[key][1].replace(",", ""))
open(out + ".txt", "w").write("\n".join(txt) + "\n")
json.dump(js, open(out + ".json", "w"), indent=1)
print("\n".join(txt))
