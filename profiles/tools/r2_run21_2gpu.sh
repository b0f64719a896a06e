cd $GRAFT_REPO_ROOT
N=${1:-2}
O=gpurun_out/r2v; mkdir -p $O
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -q > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_2gpu.log
grep -n "FAILED\|passed\|failed\|skipped\|^E  " $O/pytest_2gpu.log | tail -8
fi
for m in 63 0; do
MPM_PDL_MASK=$m MPM_BENCH_ALLRANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 --no-extras --evolved-at 0 --no-cpu-baseline > $O/bench_n${N}_mask$m.json 2> $O/bench_n${N}_mask$m.err
grep -o "\[rank [0-9]\] n_local=[0-9]* cells=[0-9]* ms_step=[0-9.]* sort=[0-9.]* p2g1=[0-9.]* p2g2=[0-9.]* update=[0-9.]* g2p=[0-9.]* exchange=[0-9.]* (mass [0-9.]* momentum [0-9.]* migration [0-9.]*)" $O/bench_n${N}_mask$m.err | sort | head -8
python - $O/bench_n${N}_mask$m.json $m <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("mask", sys.argv[2], "ms/step %.4f (phase pass %.4f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), "G %.2f"%(l["value"]/1e9), "e2e %.2f"%(l["e2e"]["value"]/1e9))
PY
done
