cd $GRAFT_REPO_ROOT
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q > $O/pytest_mr.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mr.log
grep -n "FAILED\|passed\|failed\|^E  " $O/pytest_mr.log | tail -12
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "zero_copy or binning or reproducible" > $O/pytest_p.log 2>&1; grep -n "FAILED\|passed\|failed\|^E  " $O/pytest_p.log | tail -8
python bench.py --no-cpu-baseline --no-extras > $O/bench_default.json 2> $O/bench_default.err; tail -3 $O/bench_default.err
python - $O/bench_default.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "ms/step %.3f"%l["ms_per_step"], "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "frac %.3f"%l["p2g_g2p_frac"], "e2e %.2f"%(l["e2e"]["value"]/1e9))
    for k in ("evolved","binning"):
        if k in l: print("   ",k, json.dumps(l[k])[:700])
PY
