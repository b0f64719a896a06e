cd $GRAFT_REPO_ROOT
O=gpurun_out/r2e; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -n "FAILED\|passed\|failed\|STABLE_BINNING" $O/pytest.log | tail -15
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_new.json 2> $O/bench_new.err
MPM_ATOMIC_BINNING=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_atomic.json 2> $O/bench_atomic.err
python bench.py --steps 20 --warmup 5 --presteps 100 --no-cpu-baseline > $O/bench_new_evolved.json 2> $O/bench_new_evolved.err
MPM_ATOMIC_BINNING=1 python bench.py --steps 20 --warmup 5 --presteps 100 --no-cpu-baseline > $O/bench_atomic_evolved.json 2> $O/bench_atomic_evolved.err
python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_c2.json 2>&1
python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_c3.json 2>&1
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms/step %.3f"%l["ms_per_step"], "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "frac %.3f"%l["p2g_g2p_frac"], "e2e %.2f"%(l["e2e"]["value"]/1e9))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_p2g1_cell|k_p2g2_cell|k_g2p_cell|k_rank_count|k_rank_place|k_block_order' --launch-skip 28 --launch-count 7 -o $O/full_c4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
