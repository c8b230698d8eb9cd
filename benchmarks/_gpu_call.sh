cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r1_pytest_gpu_v21.log
python -c "import __graft_entry__ as g; g.smoke()"
