set -x
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; free -g | head -2; nproc
timeout 1500 python -m pytest tests -q -m gpu --durations=25 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/r2a_pytest.log
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; tail -c 3000 gpurun_out/r2a_bench_n1.json
for rows in 1 2; do timeout 300 python bench.py --workload sens --sens-rows $rows --steps 50 --no-cpu > gpurun_out/r2a_sens_m3d_rows$rows.json 2> gpurun_out/r2a_sens_m3d_rows$rows.err; tail -c 1500 gpurun_out/r2a_sens_m3d_rows$rows.json; done
timeout 300 python bench.py --workload sens-hog1p --steps 200 > gpurun_out/r2a_sens_hog1p.json 2> gpurun_out/r2a_sens_hog1p.err; tail -c 1500 gpurun_out/r2a_sens_hog1p.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sens_matvec -s 3 -c 2 -o gpurun_out/r2a_k_sens python bench.py --workload sens --steps 5 --warmup 2 --no-cpu > gpurun_out/r2a_ncu_sens.log 2>&1; tail -3 gpurun_out/r2a_ncu_sens.log
