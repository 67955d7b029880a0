set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2c_pytest_gpu.log 2>&1
tail -25 gpurun_out/r2c_pytest_gpu.log
for mode in ramp uniform ramp8 ; do
  case $mode in
    ramp) export -n NCME_HOST_PIPE_UNIFORM; unset NCME_HOST_PIPE_UNIFORM; unset NCME_HOST_PIPE_CHUNKS;;
    uniform) export NCME_HOST_PIPE_UNIFORM=1; unset NCME_HOST_PIPE_CHUNKS;;
    ramp8) unset NCME_HOST_PIPE_UNIFORM; export NCME_HOST_PIPE_CHUNKS=8;;
  esac
  timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu --no-solve > gpurun_out/r2c_e2e_$mode.json 2>gpurun_out/r2c_e2e_$mode.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2c_e2e_$mode.json').read().strip().splitlines()[-1]); print('$mode e2e ms', d['e2e']['ms_per_step'], 'GB/s', d['e2e']['value'], 'matvec ms', d['ms_per_step'], 'checksum', d['e2e'].get('checksum_sum_y_states'))"
done
