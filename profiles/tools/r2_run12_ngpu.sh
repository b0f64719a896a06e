cd $GRAFT_REPO_ROOT
N=$1
O=gpurun_out/r2n8; mkdir -p $O
MPM_BENCH_ALLRANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err
grep "^\[rank" $O/bench_c4_n$N.err | tr ']' '\n' | grep -c n_local
grep -o "\[rank [0-9]\] n_local=[0-9]* cells=[0-9]* ms_step=[0-9.]* sort=[0-9.]* p2g1=[0-9.]* p2g2=[0-9.]* update=[0-9.]* g2p=[0-9.]* exchange=[0-9.]*" $O/bench_c4_n$N.err | head -8
if [ "$N" = "8" ]; then
MPM_BENCH_ALLRANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --workload c5 --no-extras --evolved-at 0 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err
grep -o "\[rank [0-9]\] n_local=[0-9]* cells=[0-9]* ms_step=[0-9.]* sort=[0-9.]* p2g1=[0-9.]* p2g2=[0-9.]* update=[0-9.]* g2p=[0-9.]* exchange=[0-9.]*" $O/bench_c5_n$N.err | head -8
fi
python - $O/bench_c4_n$N.json $O/bench_c5_n$N.json <<'PY'
import json,sys,os
for f in sys.argv[1:]:
    if not os.path.exists(f): continue
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.3f"%l["ms_per_step"], "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "e2e %.2f"%(l["e2e"]["value"]/1e9))
        for k in ("evolved","weak"):
            if k in l: print("   ",k, "ms %.3f G %.2f"%(l[k]["ms_per_step"], l[k]["value"]/1e9))
    except Exception as e: print(f,"FAILED",e)
PY
