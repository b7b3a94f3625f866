set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fast.py -m gpu -x -q -s 2>&1 | tail -25
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 2500 gpurun_out/bench_fast.json; tail -5 gpurun_out/bench_fast.err
