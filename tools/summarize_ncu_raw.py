"""Condenses `ncu -i X.ncu-rep --page raw --csv` files into one JSON summary (the metrics the notes in profiles/ quote):
   python tools/summarize_ncu_raw.py out.json a.raw.csv b.raw.csv ..."""
import csv
import json
import re
import sys

KEEP = {
    "gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "regs", "launch__grid_size": "grid",
    "launch__block_size": "block", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_pipe_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active": "tmem_pipe_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "smsp__inst_executed.sum": "warp_insts",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
}
out = []
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    ik = h.index("Kernel Name")
    for r in rows[2:]:
        e = {"kernel": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("zk::", "").strip(), "source": path.split("/")[-1]}
        for i, name in enumerate(h):
            if name in KEEP and r[i] != "":
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if KEEP[name].startswith("dram_r") or KEEP[name].startswith("dram_w"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                if KEEP[name] == "time_us":
                    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                e[KEEP[name]] = v
        if "dram_read" in e and "dram_write" in e:
            e["dram_bytes"] = e["dram_read"] + e["dram_write"]
        out.append(e)
json.dump(out, open(sys.argv[1], "w"), indent=1)
for e in out:
    print(f"{e['kernel'][:44]:44s} {e.get('time_us', 0):8.1f}us grid {int(e.get('grid', 0)):6d} x {int(e.get('block', 0)):4d} regs {int(e.get('regs', 0)):3d} "
          f"sm {e.get('sm_throughput_pct', 0):5.1f}% imad(fmaheavy) {e.get('fmaheavy_pipe_pct', 0):5.1f}% alu {e.get('alu_pipe_pct', 0):5.1f}% tensor {e.get('tensor_pipe_pct', 0):5.1f}% "
          f"warps {e.get('warps_active_pct', 0):5.1f}% dram {e.get('dram_pct', 0):5.1f}% ({e.get('dram_bytes', 0) / 1e6:8.1f} MB) L2hit {e.get('l2_hit_pct', 0):5.1f}%")
