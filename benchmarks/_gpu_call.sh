cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo new; python benchmarks/single_pass.py
echo auto; timeout 900 python benchmarks/configs.py 2>&1 | tee gpurun_out/r1_configs_1_3_4_v18.json | cut -c1-300
echo base; cd _base; timeout 900 python benchmarks/configs.py --configs 3 2>&1 | cut -c1-300; cd ..
timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200
