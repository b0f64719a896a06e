cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ag; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -4 $O/pytest.log
for v in overlap serial; do
if [ $v = serial ]; then export MPM_NO_BIN_OVERLAP=1; fi
python bench.py --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
python - $O/bench_$v.json $v <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step %.4f (phase pass %.4f) G %.2f"%(l["ms_per_step"], l["ms_per_step_phase_pass"], l["value"]/1e9), {k:round(v,4) for k,v in l["phase_ms"].items() if k!='exchange'})
for k,c in l["configs"].items(): print("   ",k,"ms %.4f G %.2f sort %.4f"%(c["ms_per_step"],c["value"]/1e9,c["phase_ms"]["sort"]))
print("    evolved ms %.3f"%l["evolved"]["ms_per_step"], "weak ms %.3f"%l["weak"]["ms_per_step"], "unordered", l["binning"])
PY
done
