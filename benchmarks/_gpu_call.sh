cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo runs; timeout 900 python benchmarks/configs.py --configs 3 2>&1 | cut -c1-300
echo noruns; B2SV_ADJOINT_RUNS=0 timeout 900 python benchmarks/configs.py --configs 3 2>&1 | cut -c1-300
