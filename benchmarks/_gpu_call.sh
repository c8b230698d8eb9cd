cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r1_pytest_gpu_v22.log
timeout 600 python bench.py > gpurun_out/r1_bench_v22.json 2> gpurun_out/r1_bench_v22.err; tail -c 400 gpurun_out/r1_bench_v22.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference_v22.json 2>/dev/null; cut -c1-200 gpurun_out/r1_bench_reference_v22.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_ncu_launches_v22.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_launches_v22.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
