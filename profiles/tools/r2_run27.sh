cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ac; mkdir -p $O
for w in c3 c4; do for b in 8 4; do
st=20; [ $w = c3 ] && st=30
MPM_BLOCK_EDGE=$b timeout 300 python bench.py --workload $w --steps $st --warmup 5 --no-cpu-baseline --no-extras --evolved-at 0 > $O/${w}_b$b.json 2> $O/${w}_b$b.err
python - $O/${w}_b$b.json $w $b <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "B", sys.argv[3], "ms/step %.4f (phase pass %.4f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), {k:round(v,4) for k,v in l["phase_ms"].items() if k!='exchange'})
except Exception as e: print(sys.argv[2], sys.argv[3], "FAILED", e)
PY
done; done
tail -3 $O/c4_b4.err
