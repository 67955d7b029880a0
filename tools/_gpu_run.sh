set -x
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_headline.py tests/test_gpu_matvec.py -m gpu -q -x ) > gpurun_out/r2d_pytest.log 2>&1
tail -8 gpurun_out/r2d_pytest.log
for mode in zc nozc zc_uniform zc12; do
  unset NCME_HOST_PIPE_UNIFORM NCME_HOST_PIPE_CHUNKS NCME_HOST_ZEROCOPY
  case $mode in
    nozc) export NCME_HOST_ZEROCOPY=0;;
    zc_uniform) export NCME_HOST_PIPE_UNIFORM=1;;
    zc12) export NCME_HOST_PIPE_CHUNKS=12;;
  esac
  timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu --no-solve > gpurun_out/r2d_e2e_$mode.json 2>gpurun_out/r2d_e2e_$mode.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2d_e2e_$mode.json').read().strip().splitlines()[-1]); print('$mode e2e ms', d['e2e']['ms_per_step'], 'GB/s', d['e2e']['value'], 'matvec ms', d['ms_per_step'], 'checksum', d['e2e'].get('checksum_sum_y_states'))"
done
