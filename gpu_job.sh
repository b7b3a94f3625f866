set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fast.py -m gpu -x -q -s 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fast2.json 2> gpurun_out/bench_fast2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_fast2.json'))
print(d['value'], d['stage_ms_last_step'], d['roofline']['frac'], d['clocks'])
PY
tail -3 gpurun_out/bench_fast2.err
ncu --set full --clock-control none --import-source on -k regex:k_phase1 -s 1 -c 1 -o gpurun_out/prof_phase1_v2 python bench.py --steps 1 --warmup 1 --batch 197 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
