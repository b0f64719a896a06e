cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ad; mkdir -p $O
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -q > $O/pytest_2gpu.log 2>&1; tail -3 $O/pytest_2gpu.log
for rb in 0 3; do
MPM_BENCH_ALLRANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$rb bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --evolved-at 0 --no-cpu-baseline --rebalance $rb > $O/bench_n2_rb$rb.json 2> $O/bench_n2_rb$rb.err
grep -o "\[rank [0-9]\] n_local=[0-9]* cells=[0-9]* ms_step=[0-9.]* sort=[0-9.]* p2g1=[0-9.]* p2g2=[0-9.]* update=[0-9.]* g2p=[0-9.]* exchange=[0-9.]* (mass [0-9.]* momentum [0-9.]* migration [0-9.]*)" $O/bench_n2_rb$rb.err | sort | head -2
python - $O/bench_n2_rb$rb.json $rb <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("rebalance", sys.argv[2], "ms/step %.4f (phase pass %.4f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), "G %.2f"%(l["value"]/1e9))
except Exception as e: print("FAILED", e)
PY
tail -2 $O/bench_n2_rb$rb.err | cut -c1-300
done
