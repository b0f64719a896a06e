cd $GRAFT_REPO_ROOT
O=gpurun_out/r2t; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
python bench.py --no-cpu-baseline > $O/bench_pdl.json 2> $O/bench_pdl.err; tail -2 $O/bench_pdl.err
MPM_NO_PDL=1 python bench.py --no-cpu-baseline > $O/bench_nopdl.json 2> $O/bench_nopdl.err
python - $O/bench_pdl.json $O/bench_nopdl.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "ms/step %.3f (phase pass %.3f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "e2e %.2f"%(l["e2e"]["value"]/1e9))
    for k,c in l["configs"].items(): print("   ",k,"ms %.4f (phase pass %.4f) G %.2f"%(c["ms_per_step"],c["ms_per_step_phase_pass"],c["value"]/1e9), {a:round(b,4) for a,b in c["phase_ms"].items()})
    print("    evolved ms %.3f (%.3f)"%(l["evolved"]["ms_per_step"],l["evolved"]["ms_per_step_phase_pass"]), "weak ms %.3f"%l["weak"]["ms_per_step"])
PY
