set -x
for v in k1 forced; do
  if [ $v = forced ]; then export NCME_FORCE_SHARDED_KERNEL=1; fi
  timeout 300 python bench.py --steps 200 --warmup 10 --no-solve --no-cpu > gpurun_out/r2d_n1_$v.json 2> gpurun_out/r2d_n1_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2d_n1_$v.json').read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['value'], d['per_step']['median_ms'], d['config']['assemble_s'], d['config']['expand_s'])"
done
unset NCME_FORCE_SHARDED_KERNEL
timeout 2700 python -m pytest tests -v -m gpu --durations=60 --timeout=900 -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?"; grep -E "PASSED|FAILED|ERROR|SKIPPED" gpurun_out/r2d_pytest.log | grep -v PASSED | head -20; tail -75 gpurun_out/r2d_pytest.log
