cd $GRAFT_REPO_ROOT
O=gpurun_out/r2o; mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "parts or reproducible or oracle or invariants or stable" > $O/pytest_parts.log 2>&1; tail -3 $O/pytest_parts.log
python bench.py --no-cpu-baseline > $O/bench_split.json 2> $O/bench_split.err; tail -2 $O/bench_split.err
MPM_BLOCK_SPLIT_ITEMS=0 python bench.py --no-cpu-baseline > $O/bench_nosplit.json 2> $O/bench_nosplit.err
python - $O/bench_split.json $O/bench_nosplit.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "ms/step %.3f"%l["ms_per_step"], "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()})
    for k,c in l["configs"].items(): print("   ",k,"ms %.4f G %.2f"%(c["ms_per_step"],c["value"]/1e9), {a:round(b,4) for a,b in c["phase_ms"].items()})
    print("    evolved ms %.3f"%l["evolved"]["ms_per_step"], "weak ms %.3f"%l["weak"]["ms_per_step"])
PY
