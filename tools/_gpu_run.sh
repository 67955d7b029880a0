set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
run() { name=$1; n=$2; lv=$3; shift; shift; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $n --steps 20 --warmup 5 --solve-method bdf --levels $lv > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err; grep "ncme bdf profile" gpurun_out/r2l_$name.err | tail -1; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2l_$name.json").read().strip().splitlines()[-1])
    s=d.get("solve") or {}
    print("$name", d["config"]["states"], round(d["ms_per_step"],5), "solve", s.get("wall_s"), s.get("steps"), s.get("rejected"), s.get("rhs_evals"), s.get("launches"), s.get("mean_x"), "api", s.get("solve_api_wall_s"), (s.get("solve_api") or {}).get("breakdown_s"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2l_$name.err").read()[-2500:])
PY
}
run n4_L310_fused 4 310 NCME_BDF_FUSED_MAX_ROWS_SHARDED=100000000
run n4_L310_classic 4 310 NCME_BDF_NO_SHARDED_FUSED=1 NCME_BDF_PROFILE=1
run n4_L390_fused 4 390 NCME_BDF_FUSED_MAX_ROWS_SHARDED=100000000
