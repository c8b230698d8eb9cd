cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in 1 2 3; do timeout 900 python benchmarks/configs.py --configs 3 2>&1 | cut -c100-200; done
timeout 900 python benchmarks/configs.py 2>&1 | tee gpurun_out/r1_configs_1_3_4_v22.json | cut -c1-200
python benchmarks/adjoint_breakdown.py 2>&1 | tail -3 | cut -c1-250
