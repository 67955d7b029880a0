set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
run() { name=$1; n=$2; shift; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $n --steps 20 --warmup 5 --solve-method bdf > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2f_$name.json").read().strip().splitlines()[-1])
    s=d.get("solve") or {}
    print("$name", round(d["ms_per_step"],5), round(d["value"]), {k:(round(v,5) if isinstance(v,float) else v) for k,v in d["per_step"].items() if k!="note"}, d["gpu_launches"], "assemble", d["config"]["assemble_s"], "solve", s.get("wall_s"), s.get("rhs_evals"), s.get("launches"), "api", s.get("solve_api_wall_s"), (s.get("solve_api") or {}).get("breakdown_s"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2f_$name.err").read()[-2500:])
PY
}
run n4_one 4 NCME_X=1
run n4_two 4 NCME_P2P_TWO_LAUNCH=1
run n4_one_hostred 4 NCME_NO_DEV_ALLREDUCE=1
run n2_one 2 NCME_X=1
run n2_two 2 NCME_P2P_TWO_LAUNCH=1
