set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 250 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/r02_launches_dgg_step.csv python scripts/step_one.py 6 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 250 -c 400 --csv --log-file gpurun_out/r02_launches_models.csv python scripts/prof_models.py 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"linear_tf32x3|gemm_tn_tf32x3|dgg_fwd_fused|dgg_bwd_fused" -s 8 -c 5 -o gpurun_out/r02_full_dgg -f python scripts/step_one.py 4 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"spmm_|gat_|edge_mlp|row_firstk" -s 14 -c 14 -o gpurun_out/r02_full_models -f python scripts/prof_models.py 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*.csv
