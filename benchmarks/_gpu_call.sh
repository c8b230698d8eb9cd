cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r1_pytest_gpu_v20.log
timeout 600 python bench.py > gpurun_out/r1_bench_v20.json 2> gpurun_out/r1_bench_v20.err; tail -c 600 gpurun_out/r1_bench_v20.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference_v20.json 2>/dev/null; cut -c1-300 gpurun_out/r1_bench_reference_v20.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_ncu_launches_v20.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_launches_v20.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_exec -s 20 -c 2 -f -o gpurun_out/r1_tile_v20 python benchmarks/ncu_target.py 30 > gpurun_out/r1_ncu_v20.log 2>&1
tail -2 gpurun_out/r1_ncu_v20.log
./benchmarks/micro/inplace_stream | tee gpurun_out/r1_inplace_stream.jsonl
