set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/r2m_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2m_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 420 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
tail -3 gpurun_out/r2m_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2m_bench_n1.json').read().strip().splitlines()[-1]); s=d['solve']; print('matvec', d['ms_per_step'], d['value'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e'].get('link'), 'bdf', s['wall_s'], s['all_methods']['bdf']['wall_s_runs'], 'dp5', s['all_methods']['dp5']['wall_s_runs'], 'api', s['solve_api_wall_s'], s['solve_api']['breakdown_s'], 'cpu', d['cpu_baseline']['value'], 'tele', d['parity_configs']['telegraph_adaptive_solve_ms']['best'])"
