cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ae; mkdir -p $O
for rb in 0 4; do
pre=0; [ $rb = 0 ] && pre=24
MPM_BENCH_ALLRANKS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$rb bench.py --gpus 8 --steps 20 --warmup 5 --workload c5 --no-extras --evolved-at 0 --no-cpu-baseline --rebalance $rb --presteps $pre > $O/bench_c5_rb$rb.json 2> $O/bench_c5_rb$rb.err
grep -o "\[rank [0-9]\] n_local=[0-9]* cells=[0-9]* ms_step=[0-9.]* sort=[0-9.]* p2g1=[0-9.]* p2g2=[0-9.]* update=[0-9.]* g2p=[0-9.]* exchange=[0-9.]* (mass [0-9.]* momentum [0-9.]* migration [0-9.]*)" $O/bench_c5_rb$rb.err | sort | head -8
python - $O/bench_c5_rb$rb.json $rb <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("rebalance", sys.argv[2], "ms/step %.4f (phase pass %.4f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), "G %.2f"%(l["value"]/1e9))
except Exception as e: print("FAILED", e)
PY
done
