set -x
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_solve.py tests/test_gpu_examples.py tests/test_gpu_sens.py tests/test_gpu_statespace.py -m gpu -q -x ) > gpurun_out/r2e_pytest.log 2>&1
tail -8 gpurun_out/r2e_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --solve-method bdf > gpurun_out/r2e_bench_bdf.json 2>gpurun_out/r2e_bench_bdf.err
python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_bdf.json').read().strip().splitlines()[-1]); s=d['solve']; print('bdf wall', s['wall_s'], s['steps'], s['rhs_evals'], s['launches'], 'api', s.get('solve_api_wall_s'), s.get('solve_api',{}).get('breakdown_s'))"
NCME_BDF_PROFILE=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --solve-method bdf 2>&1 | grep -i "bdf profile\|blocked\|wall" | head -5
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_bdf|k_gm|k_matrix_diag' -c 2600 --csv --log-file gpurun_out/r2e_launches_bdf.csv python bench.py --steps 3 --warmup 3 --no-cpu --solve-method bdf > gpurun_out/r2e_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r2e_launches_bdf.csv | head -30
