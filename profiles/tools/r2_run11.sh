cd $GRAFT_REPO_ROOT
O=gpurun_out/r2l; mkdir -p $O
for v in default ru3 ru4; do
  if [ $v = default ]; then unset MPM_B200_LIB; else export MPM_B200_LIB=$PWD/mls-mpm-godot_b200/build/ab/libmpm_$v.so; fi
  python bench.py --no-cpu-baseline --no-extras > $O/bench_$v.json 2> $O/bench_$v.err
done
unset MPM_B200_LIB
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "cell or binning or reproducible" > $O/pytest.log 2>&1; grep -n "FAILED\|passed\|failed\|^E  " $O/pytest.log | tail -5
python - $O/bench_default.json $O/bench_ru3.json $O/bench_ru4.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.3f"%l["ms_per_step"], "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "frac %.3f"%l["p2g_g2p_frac"], "evolved %.3f"%l["evolved"]["ms_per_step"], {k:round(v,3) for k,v in l["evolved"]["phase_ms"].items()})
    except Exception as e: print(f, "FAILED", e)
PY
