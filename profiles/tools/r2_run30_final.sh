cd $GRAFT_REPO_ROOT
O=gpurun_out/r2final; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -n "FAILED\|passed\|failed\|^E  " $O/pytest.log | tail -6
python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-400 $O/bench_reference.json
python - $O/bench_default.json <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.3f G %.2f frac %.3f roofline %s %.3f e2e %.2f q16 %.2f launches %d"%(l["ms_per_step"], l["value"]/1e9, l["p2g_g2p_frac"], l["roofline"]["kernel"], l["roofline"]["frac"], l["e2e"]["value"]/1e9, l["e2e_q16"]["value"]/1e9, l["gpu_launches"]))
print("keys", sorted(l.keys()))
print("cpu_baseline", l["cpu_baseline"]["value"], l["cpu_baseline"]["cores"], "clocks", l["clocks"])
PY
