set -x
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 600 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2b_pytest_gpu.log 2>&1
tail -30 gpurun_out/r2b_pytest_gpu.log
timeout 120 python tools/pcie_ceiling.py > gpurun_out/r2b_pcie_ceiling.json 2> gpurun_out/r2b_pcie_ceiling.err; cat gpurun_out/r2b_pcie_ceiling.json
( time timeout 420 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
tail -c 1500 gpurun_out/r2b_bench_n1.json; tail -5 gpurun_out/r2b_bench_n1.err
