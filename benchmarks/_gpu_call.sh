cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "adjoint or vjp" 2>&1 | tail -3
for i in 1 2; do timeout 900 python benchmarks/configs.py --configs 3 2>&1 | cut -c100-200; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_ncu_launches_adjoint_v23.csv python benchmarks/adjoint_once.py > gpurun_out/r1_ncu_adjoint.log 2>&1
