cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ab; mkdir -p $O
cd mls-mpm-godot_b200; cp libmpm_b200.so /tmp/keep.so; cd ..
for v in keep rank10 rank12 keep; do
if [ $v = keep ]; then cp /tmp/keep.so mls-mpm-godot_b200/libmpm_b200.so; else cp mls-mpm-godot_b200/build/ab/libmpm_$v.so mls-mpm-godot_b200/libmpm_b200.so; fi
python bench.py --no-cpu-baseline --no-extras --evolved-at 0 > $O/bench_$v.json 2> $O/bench_$v.err
python - $O/bench_$v.json $v <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step %.4f (phase pass %.4f)"%(l["ms_per_step"], l["ms_per_step_phase_pass"]), {k:round(v,4) for k,v in l["phase_ms"].items() if k!='exchange'})
PY
done
cp /tmp/keep.so mls-mpm-godot_b200/libmpm_b200.so
