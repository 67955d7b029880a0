mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_solve.py -m gpu -q -x ) > gpurun_out/r2l_pytest.log 2>&1
tail -4 gpurun_out/r2l_pytest.log
for rep in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --solve-method bdf > gpurun_out/r2l_bench_bdf_$rep.json 2>gpurun_out/r2l_bench_bdf_$rep.err
python -c "
import json; d=json.loads(open('gpurun_out/r2l_bench_bdf_$rep.json').read().strip().splitlines()[-1]); s=d['solve']; print('bdf wall', s['wall_s'], s['steps'], s['rhs_evals'], s['launches'], 'api integrate', s['solve_api']['breakdown_s']['integrate'], s['mean_x'])"
done
