cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_strict.py -m gpu -x -q 2>&1 | tail -5
for w in kms2 cggi; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    print('$w', round(d['value'],1), {k:round(v,2) for k,v in d['stage_ms_last_step'].items()}, d['decrypt_check'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
