set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_solve.py tests/test_gpu_sens.py -m gpu -q -x ) > gpurun_out/r2f_pytest.log 2>&1
tail -12 gpurun_out/r2f_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --solve-method bdf > gpurun_out/r2f_bench_bdf_n1.json 2>gpurun_out/r2f_bench_bdf_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2f_scale_n2.json 2>gpurun_out/r2f_scale_n2.err
for f in r2f_bench_bdf_n1 r2f_scale_n2; do python -c "
import json; d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); s=d.get('solve') or {}; print('$f', 'matvec ms', d['ms_per_step'], 'GB/s', d['value'], 'e2e', d['e2e']['value'], 'solve', s.get('wall_s'), s.get('steps'), s.get('rhs_evals'), 'all', {k:v['wall_s'] for k,v in (s.get('all_methods') or {}).items()}, 'api', s.get('solve_api_wall_s'), 'assemble', d['config'].get('assemble_s'))"; done
tail -3 gpurun_out/r2f_scale_n2.err
