cd $GRAFT_REPO_ROOT
O=gpurun_out/r2u; mkdir -p $O
for m in 0 15 13 2; do
for w in c4 c2; do
st=20; [ $w = c2 ] && st=200
MPM_PDL_MASK=$m python bench.py --workload $w --steps $st --warmup 5 --no-cpu-baseline --no-extras --evolved-at 0 > $O/${w}_mask$m.json 2> $O/${w}_mask$m.err
python - $O/${w}_mask$m.json $m $w <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[3], "mask", sys.argv[2], "ms/step %.4f (phase pass %.4f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), {k:round(v,4) for k,v in l["phase_ms"].items() if k!='exchange'})
PY
done
done
