cd $GRAFT_REPO_ROOT
O=gpurun_out/r2s; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1232 -c 48 --csv --log-file $O/launches_evolved.csv python bench.py --presteps 100 --steps 2 --warmup 3 --no-cpu-baseline --no-extras --evolved-at 0 > $O/ncu_evolved.log 2>&1
python - $O/launches_evolved.csv <<'PY'
import csv,sys,re
from collections import OrderedDict
lines=[l for l in open(sys.argv[1]) if l.startswith('"')]
r=list(csv.reader(lines)); c={k:i for i,k in enumerate(r[0])}
agg=OrderedDict()
for x in r[1:]:
    n=re.sub(r"\(.*","",x[c["Kernel Name"]]).replace("void ","")
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(x[c["Metric Value"]].replace(",",""))
for n,(k,t) in agg.items(): print("%-40s n=%3d mean=%10.1f us"%(n,k,t/k/1e3))
PY
ncu --set full --clock-control none --import-source on -k regex:'k_p2g1_cell|k_g2p_cell|k_rank_place|k_block_order' --launch-skip 408 --launch-count 4 -o $O/full_evolved python bench.py --presteps 100 --steps 2 --warmup 3 --no-cpu-baseline --no-extras --evolved-at 0 > $O/ncu_full_evolved.log 2>&1
ncu -i $O/full_evolved.ncu-rep --page raw --csv > $O/full_evolved_raw.csv 2>/dev/null
python - $O/full_evolved_raw.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1]))); h=rows[0]; col={c:i for i,c in enumerate(h)}
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio",
"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
for m in want:
    if m in col: print("%-80s %s"%(m[:80],[r[col[m]][:22] for r in rows[2:]]))
PY
