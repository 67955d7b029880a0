mkdir -p gpurun_out
for q in 1 2; do
NCME_HOST_PIPE_QUEUES=$q timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu --no-solve > gpurun_out/r2k_q$q.json 2>gpurun_out/r2k_q$q.err
python -c "
import json; d=json.loads(open('gpurun_out/r2k_q$q.json').read().strip().splitlines()[-1]); print('queues $q e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'], 'floor', d['e2e']['link']['both_directions_floor_ms'], 'checksum', d['e2e']['checksum_sum_y_states'])"
done
NCME_HOST_PIPE_QUEUES=2 NCME_HOST_PIPE_TRACE=1 timeout 200 python bench.py --steps 10 --warmup 5 --no-cpu --no-solve 2>&1 >/dev/null | grep "host pipe"
