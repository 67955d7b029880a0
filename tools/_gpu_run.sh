set -x
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_headline.py tests/test_gpu_matvec.py -m gpu -q -x ) > gpurun_out/r2h_pytest.log 2>&1
tail -4 gpurun_out/r2h_pytest.log
for rep in 1 2; do
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu --no-solve > gpurun_out/r2h_e2e_$rep.json 2>gpurun_out/r2h_e2e_$rep.err
python -c "
import json; d=json.loads(open('gpurun_out/r2h_e2e_$rep.json').read().strip().splitlines()[-1]); print('rep $rep e2e ms', d['e2e']['ms_per_step'], 'GB/s', d['e2e']['value'], d['e2e']['link']['both_directions_floor_ms'], d['e2e']['link']['e2e_frac_of_link_floor'], 'checksum', d['e2e'].get('checksum_sum_y_states'))"
done
