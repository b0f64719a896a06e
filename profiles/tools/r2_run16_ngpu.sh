cd $GRAFT_REPO_ROOT
N=$1
O=gpurun_out/r2q; mkdir -p $O
for it in 0 4 8; do
MPM_BLOCK_SPLIT_ITEMS=$it MPM_BENCH_ALLRANKS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$it bench.py --gpus $N --steps 20 --warmup 5 --no-extras --evolved-at 0 --no-cpu-baseline > $O/c4_n${N}_items$it.json 2> $O/c4_n${N}_items$it.err
echo "items $it"
grep -o "\[rank [0-9]\] n_local=[0-9]* cells=[0-9]* ms_step=[0-9.]* sort=[0-9.]* p2g1=[0-9.]* p2g2=[0-9.]* update=[0-9.]* g2p=[0-9.]* exchange=[0-9.]* (mass [0-9.]* momentum [0-9.]* migration [0-9.]*)" $O/c4_n${N}_items$it.err | sort | head -8
done
