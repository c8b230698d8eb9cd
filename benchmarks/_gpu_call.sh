cd $GRAFT_REPO_ROOT
timeout 120 python benchmarks/configs.py 2>&1 | tee gpurun_out/r1_configs_1_3_4_v23.json | cut -c1-220
