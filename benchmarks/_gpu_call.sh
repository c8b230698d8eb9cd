cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1c_pytest_gpu.log
cat gpurun_out/r1c_pytest_gpu.log
timeout 900 python benchmarks/tile_experiments.py > gpurun_out/r1c_tile_experiments.jsonl 2>&1
cat gpurun_out/r1c_tile_experiments.jsonl
