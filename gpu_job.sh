cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast.py -m gpu -x -q -s 2>&1 | tail -12
for w in kms2 kms8block; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    print('$w', round(d['value'],1), d['mode'], {k:round(v,2) for k,v in d['stage_ms_last_step'].items()}, 'frac', round(d['roofline']['frac'],3), d['decrypt_check'])
except Exception as e:
    print('$w FAILED', e); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
