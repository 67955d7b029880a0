set -x
timeout 900 python -m pytest tests -v -m gpu --durations=15 --timeout=300 -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?"; grep -E "FAILED|ERROR" gpurun_out/r2m_pytest.log | head -20; tail -22 gpurun_out/r2m_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; tail -c 600 gpurun_out/r2m_bench_n1.json; tail -3 gpurun_out/r2m_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --solve-method bdf --solve-t 2.0 > gpurun_out/r2m_ncu_bench.log 2>&1; tail -2 gpurun_out/r2m_ncu_bench.log | cut -c1-300; wc -l gpurun_out/r2m_launches.csv
SAN_TIMEOUT=240 tools/sanitize.sh gpurun_out
