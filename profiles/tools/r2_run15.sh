cd $GRAFT_REPO_ROOT
O=gpurun_out/r2p; mkdir -p $O
for it in 0 2 3 4 6 8; do
MPM_BLOCK_SPLIT_ITEMS=$it python bench.py --workload c3 --steps 50 --warmup 5 --no-cpu-baseline --no-extras --evolved-at 0 > $O/c3_items$it.json 2> $O/c3_items$it.err
python - $O/c3_items$it.json $it <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("items", sys.argv[2], "ms/step %.4f"%l["ms_per_step"], {k:round(v,4) for k,v in l["phase_ms"].items()})
PY
done
