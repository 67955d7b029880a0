set -x
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; free -g | head -2; nproc
timeout 1700 python -m pytest tests -q -m gpu --durations=25 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/r2a_pytest.log
for rows in 1 2; do timeout 300 python bench.py --workload sens --sens-rows $rows --steps 50 --no-cpu > gpurun_out/r2a_sens_m3d_rows$rows.json 2> gpurun_out/r2a_sens_m3d_rows$rows.err; tail -c 1500 gpurun_out/r2a_sens_m3d_rows$rows.json; done
timeout 300 python bench.py --workload sens-hog1p --steps 200 > gpurun_out/r2a_sens_hog1p.json 2> gpurun_out/r2a_sens_hog1p.err; tail -c 1500 gpurun_out/r2a_sens_hog1p.json
timeout 300 python bench.py --workload sens-hog1p --sens-rows 2 --steps 200 --no-cpu > gpurun_out/r2a_sens_hog1p_rows2.json 2>&1; tail -c 600 gpurun_out/r2a_sens_hog1p_rows2.json
SAN_TIMEOUT=270 tools/sanitize.sh gpurun_out
