set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench_new.json 2> gpurun_out/r2a/bench_new.err
MPM_B200_LIB=$PWD/mls-mpm-godot_b200/build/ab/libmpm_r1.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench_r1.json 2> gpurun_out/r2a/bench_r1.err
MPM_B200_LIB=$PWD/mls-mpm-godot_b200/build/ab/libmpm_ctas4.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench_ctas4.json 2> gpurun_out/r2a/bench_ctas4.err
MPM_ATOMIC_BINNING=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench_atomic.json 2> gpurun_out/r2a/bench_atomic.err
python bench.py --steps 20 --warmup 5 --presteps 100 --no-cpu-baseline > gpurun_out/r2a/bench_new_evolved.json 2> gpurun_out/r2a/bench_new_evolved.err
MPM_B200_LIB=$PWD/mls-mpm-godot_b200/build/ab/libmpm_r1.so python bench.py --steps 20 --warmup 5 --presteps 100 --no-cpu-baseline > gpurun_out/r2a/bench_r1_evolved.json 2> gpurun_out/r2a/bench_r1_evolved.err
python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench_c2.json 2>&1
python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench_c3.json 2>&1
for f in gpurun_out/r2a/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms/step %.3f"%l["ms_per_step"], "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "frac %.3f"%l["p2g_g2p_frac"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
