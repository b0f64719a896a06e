cd $GRAFT_REPO_ROOT
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "quantised or programmatic or timing_levels" > $O/pytest_new.log 2>&1; tail -5 $O/pytest_new.log
python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python - $O/bench.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "ms/step %.3f (phase pass %.3f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "e2e %.2f"%(l["e2e"]["value"]/1e9), "e2e_q16 %.2f"%(l["e2e_q16"]["value"]/1e9), "launches", l["gpu_launches"])
    for k,c in l["configs"].items(): print("   ",k,"ms %.4f (phase pass %.4f) G %.2f"%(c["ms_per_step"],c["ms_per_step_phase_pass"],c["value"]/1e9))
    print("    evolved ms %.3f (%.3f)"%(l["evolved"]["ms_per_step"],l["evolved"]["ms_per_step_phase_pass"]), "weak ms %.3f G %.2f"%(l["weak"]["ms_per_step"], l["weak"]["value"]/1e9))
PY
