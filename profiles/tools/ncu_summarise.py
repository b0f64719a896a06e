"""Turn one gpurun capture directory (profiles/tools/r2_run13.sh) into the tracked evidence under profiles/r2:
  <tag>_ncu_full_c4.txt       the metrics of the `ncu --set full` capture that the docs quote, one column per kernel
  <tag>_launch_summary_c4.txt per-kernel mean time and share of the launch-list pass
  traffic_c4.json             DRAM bytes per launch, read by bench.py for roofline.traffic, stamped with the commit
usage: python profiles/tools/ncu_summarise.py gpurun_out/r2n v10 <commit>"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

src, tag, commit = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "profiles", "r2")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]

rows = list(csv.reader(open(os.path.join(src, "full_c4_raw.csv"))))
head, units, data = rows[0], rows[1], rows[2:]
col = {c: i for i, c in enumerate(head)}
names = [r[col["Kernel Name"]] for r in data]
with open(os.path.join(OUT, f"{tag}_ncu_full_c4.txt"), "w") as f:
    f.write(f"commit {commit}; ncu --set full --clock-control none --import-source on, C4 dam-break, step 4, one launch per kernel\n")
    f.write("%-86s %-8s %s\n" % ("Kernel Name", "", [n[:24] for n in names]))
    for m in WANT:
        if m in col:
            f.write("%-86s %-8s %s\n" % (m, units[col[m]], [r[col[m]] for r in data]))


def short(n):
    m = re.match(r"(?:void )?([A-Za-z_0-9]+)", n)
    return m.group(1)


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_s(v, unit):
    return float(v) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[unit]


kern = OrderedDict()
for r in data:
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    kern[short(r[col["Kernel Name"]])] = dict(dram_bytes=rd + wr, dram_read_bytes=rd, dram_write_bytes=wr,
                                              ncu_duration_s=to_s(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]]),
                                              kernel=r[col["Kernel Name"]][:80])
# algorithmic bytes per launch (DESIGN.md section 3 / BASELINE.md section 3), C4: N particles, G cells
N, G = 32768000, 256 ** 3
ALG = {"k_p2g1_cell": 64 * N + 16 * G, "k_p2g2_cell": 52 * N + 28 * G, "k_g2p_cell": 72 * N + 16 * G}
for k, a in ALG.items():
    if k in kern:
        kern[k]["algorithmic_bytes"] = a
        kern[k]["dram_over_algorithmic"] = round(kern[k]["dram_bytes"] / a, 3)
with open(os.path.join(OUT, "traffic_c4.json"), "w") as f:
    json.dump(dict(source=f"profiles/r2/{tag}_ncu_full_c4.txt (one `ncu --set full --clock-control none --import-source on` capture of "
                          "step 4 of the C4 dam-break, one launch per kernel)",
                   commit=commit, workload="c4", kernels=kern), f, indent=1)

# launch list: "ID","Process ID",...,"Kernel Name",...,"Metric Name","Metric Unit","Metric Value"
lines = [l for l in open(os.path.join(src, "launches_c4.csv")) if l.startswith('"')]
lr = list(csv.reader(lines))
lc = {c: i for i, c in enumerate(lr[0])}
agg = OrderedDict()
for r in lr[1:]:
    if r[lc["Metric Name"]] != "gpu__time_duration.sum":
        continue
    n = re.sub(r"\(.*", "", r[lc["Kernel Name"]]).replace("void ", "")
    t = to_s(r[lc["Metric Value"]].replace(",", ""), {"nsecond": "ns", "usecond": "us", "msecond": "ms", "second": "s",
                                                         "ns": "ns", "us": "us", "ms": "ms", "s": "s"}[r[lc["Metric Unit"]]])
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
with open(os.path.join(OUT, f"{tag}_launch_summary_c4.txt"), "w") as f:
    f.write(f"commit {commit}; ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 2 --warmup 3 "
            "--no-cpu-baseline --no-extras --evolved-at 0 (scene set-up kernels included; cold-cache serialised times)\n")
    for n, (c, t) in agg.items():
        f.write("%-44s n=%3d mean=%10.1f us  share=%5.1f%%\n" % (n, c, t / c * 1e6, 100 * t / tot))
print(json.dumps({k: (round(v["dram_bytes"] / 1e9, 3), round(v["ncu_duration_s"] * 1e6, 1)) for k, v in kern.items()}))
