cd $GRAFT_REPO_ROOT
O=gpurun_out/r2y; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
python profiles/tools/evolved_binning_trace.py > $O/evolved_trace.txt 2>&1; grep "^step   2[1-6]\|^step  10[1-6]\|^step  14[1-6]" $O/evolved_trace.txt | cut -c1-150
python bench.py --no-cpu-baseline --no-extras > $O/bench_atomic.json 2> $O/bench_atomic.err
python - $O/bench_atomic.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "ms/step %.3f (phase pass %.3f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()})
PY
