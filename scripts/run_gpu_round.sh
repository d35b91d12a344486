set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TESTS=${TESTS:-tests}
timeout 1500 python -m pytest $TESTS -m gpu -q --tb=short -x --maxfail=${MAXFAIL:-100} 2>&1 > gpurun_out/r2_tests_full.log
tail -5 gpurun_out/r2_tests_full.log
if [ -n "$BENCH" ]; then
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 1500 gpurun_out/r2_bench.json
tail -5 gpurun_out/r2_bench.err
fi
