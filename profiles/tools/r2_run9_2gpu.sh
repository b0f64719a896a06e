cd $GRAFT_REPO_ROOT
O=gpurun_out/r2i; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -q -k "nccl or two_gpus" > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_2gpu.log
grep -n "FAILED\|passed\|failed\|skipped\|^E  " $O/pytest_2gpu.log | tail -8
MPM_BENCH_ALLRANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
tail -4 $O/bench_n2.err
python - $O/bench_n2.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "ms/step %.3f"%l["ms_per_step"], "G %.2f"%(l["value"]/1e9), {k:round(v,3) for k,v in l["phase_ms"].items()}, "e2e %.2f"%(l["e2e"]["value"]/1e9))
    for k in ("evolved","weak","binning"):
        if k in l: print("   ",k, json.dumps(l[k])[:500])
PY
