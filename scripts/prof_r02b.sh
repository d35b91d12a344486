# r02 (second half): launch list + full ncu capture of the Pubmed-shape DGG step after the encoder-backward fusion and
# the second-generation fused edge kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 250 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 24 --csv --log-file gpurun_out/r02b_launches_dgg_step.csv python scripts/step_one.py 9
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:linear_tf32x3|gemm_tn_tf32x3|dgg_fwd_fused|dgg_bwd_fused' -s 10 -c 5 -o gpurun_out/r02b_full_dgg -f python scripts/step_one.py 4
ls -la gpurun_out/r02b_*
