cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "adjoint" 2>&1 | tail -4
for i in 1 2; do timeout 900 python benchmarks/configs.py --configs 3 2>&1 | cut -c100-260; done
python benchmarks/adjoint_breakdown.py 2>&1 | tail -3 | cut -c1-250
