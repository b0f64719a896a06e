cd $GRAFT_REPO_ROOT
O=gpurun_out/r2z; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -n "FAILED\|passed\|failed\|^E  " $O/pytest.log | tail -12
python bench.py > $O/bench_c4_default.json 2> $O/bench_default.err; tail -3 $O/bench_default.err
MPM_ATOMIC_BINNING=1 python bench.py --no-cpu-baseline --no-extras > $O/bench_c4_atomic_binning.json 2> $O/bench_atomic.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
python - $O/bench_c4_default.json $O/bench_c4_atomic_binning.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "ms/step %.3f"%l["ms_per_step"], "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "frac %.3f"%l["p2g_g2p_frac"], "e2e %.2f"%(l["e2e"]["value"]/1e9))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --evolved-at 0 > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_p2g1_cell|k_p2g2_cell|k_g2p_cell|k_rank_count|k_rank_place|k_block_order|k_fix_far' --launch-skip 32 --launch-count 8 -o $O/full_c4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --evolved-at 0 > $O/ncu_full.log 2>&1
ncu -i $O/full_c4.ncu-rep --page raw --csv > $O/full_c4_raw.csv 2>/dev/null
ls -la $O | head -30
python profiles/tools/baseline_table_c1_c2.py > $O/baseline_c1_c2.json 2> $O/baseline_c1_c2.err; cat $O/baseline_c1_c2.json
