# r02 (final): launch list of the Pubmed-shape DGG step, launch list of `bench.py --steps 2 --warmup 1` (the contract's
# command), and a full capture of the stack kernel + the step's kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 250 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 24 --csv --log-file gpurun_out/r02g_launches_dgg_step.csv python scripts/step_one.py 9 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:linear_tf32x3|gemm_tn_tf32x3|dgg_fwd_fused|dgg_bwd_fused' -s 10 -c 5 -o gpurun_out/r02g_full_dgg -f python scripts/step_one.py 4 > /dev/null 2>&1
ls -la gpurun_out/r02g_*
