cd $GRAFT_REPO_ROOT
O=gpurun_out/r2r; mkdir -p $O
python bench.py --no-cpu-baseline > $O/bench_timing2.json 2> $O/bench_timing2.err; tail -2 $O/bench_timing2.err
python - $O/bench_timing2.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "ms/step %.3f (phase pass %.3f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()})
    for k,c in l["configs"].items(): print("   ",k,"ms %.4f (phase pass %.4f) G %.2f"%(c["ms_per_step"],c["ms_per_step_phase_pass"],c["value"]/1e9), {a:round(b,4) for a,b in c["phase_ms"].items()})
    print("    evolved ms %.3f (%.3f)"%(l["evolved"]["ms_per_step"],l["evolved"]["ms_per_step_phase_pass"]), "weak ms %.3f"%l["weak"]["ms_per_step"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 -c 80 --csv --log-file $O/launches_evolved.csv python bench.py --presteps 100 --steps 2 --warmup 3 --no-cpu-baseline --no-extras --evolved-at 0 > $O/ncu_evolved.log 2>&1
python - $O/launches_evolved.csv <<'PY'
import csv,sys,re
from collections import OrderedDict
lines=[l for l in open(sys.argv[1]) if l.startswith('"')]
r=list(csv.reader(lines)); c={k:i for i,k in enumerate(r[0])}
agg=OrderedDict()
rows=r[1:]
# keep only the last 120 launches (the steps after the 100 presteps are not captured by -c 300 unless small) -> print all
for x in rows:
    n=re.sub(r"\(.*","",x[c["Kernel Name"]]).replace("void ","")
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(x[c["Metric Value"]].replace(",",""))
for n,(k,t) in agg.items(): print("%-40s n=%3d mean=%10.1f %s"%(n,k,t/k,r[1][c["Metric Unit"]]))
PY
