set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -s 2>&1 | tail -12
python bench.py --workload kms8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kms8.json 2> gpurun_out/bench_kms8.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_kms8.json'))
print(d['value'], d['e2e']['value'], d['stage_ms_last_step'], d['roofline']['frac'], d['decrypt_check'], d['keygen_and_upload_s'])
PY
tail -3 gpurun_out/bench_kms8.err
