set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
free -g | head -2; nproc
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/r2a_tests.log
cat gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
