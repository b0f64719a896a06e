cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c; mkdir -p $O
python profiles/tools/dev_bin_debug.py > $O/bin_debug.log 2>&1
cat $O/bin_debug.log
timeout 900 python -m pytest tests -m gpu -q -k "reproducible or stable_sort or overflow" > $O/pytest.log 2>&1; tail -30 $O/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_new.json 2> $O/bench_new.err
python - $O/bench_new.json <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.3f"%l["ms_per_step"], {k:round(v,3) for k,v in l["phase_ms"].items()})
PY
