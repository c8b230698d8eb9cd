set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_final_pytest_gpu.log
cat gpurun_out/r2_final_pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err
tail -c 600 gpurun_out/r2_final_bench_1gpu.json; tail -3 gpurun_out/r2_final_bench_1gpu.err
timeout 600 python benchmarks/configs.py --configs 1,3,4 --reps 5 > gpurun_out/r2_final_configs.json 2>/dev/null
cat gpurun_out/r2_final_configs.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2>/dev/null
cut -c1-400 gpurun_out/r2_final_bench_reference.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-adjoint --dtype c64 --qubits 31 > gpurun_out/r2_final_bench_c64.json 2>/dev/null
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-adjoint --workload controlled --layers 2 > gpurun_out/r2_final_bench_controlled.json 2>/dev/null
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-adjoint --workload controlled --layers 2 --dtype c64 --qubits 31 > gpurun_out/r2_final_bench_controlled_c64.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_launches_adjoint.csv python benchmarks/adjoint_once.py > /dev/null 2>&1
ls -la gpurun_out/r2_ncu_launches_adjoint.csv
