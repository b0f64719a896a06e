cd $GRAFT_REPO_ROOT
O=gpurun_out/r2f; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "binning or reproducible or bench_line or config3 or bad_positions or overflow or checkpoint" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -n "FAILED\|passed\|failed" $O/pytest.log | tail -8
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -3 $O/bench_default.err
MPM_ATOMIC_BINNING=1 python bench.py --no-cpu-baseline --no-extras > $O/bench_atomic.json 2> $O/bench_atomic.err
python - $O/bench_default.json $O/bench_atomic.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "ms/step %.3f"%l["ms_per_step"], "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "frac %.3f"%l["p2g_g2p_frac"], "e2e %.2f"%(l["e2e"]["value"]/1e9))
    for k in ("evolved","configs","weak","cpu_baseline","same_config_pair","binning"):
        if k in l: print("   ",k, json.dumps(l[k])[:600])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --evolved-at 0 > $O/ncu_launches.log 2>&1
