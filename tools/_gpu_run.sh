set -x
timeout 200 python tools/telegraph_profile.py 2>&1 | tail -60
for nc in 16 8 4; do NCME_HOST_PIPE_CHUNKS=$nc timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-solve > gpurun_out/r2n_e2e_$nc.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2n_e2e_$nc.json').read().strip().splitlines()[-1]); print('chunks $nc e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'matvec', d['ms_per_step'])"; done
SAN_SEL="test_bdf_fused or test_telegraph_example or test_adaptive_solve_reference_tests or test_prune_by_mass_matches_oracle or test_sens_telegraph or test_fixed_space_solve" SAN_TOOLS=racecheck SAN_TIMEOUT=330 tools/sanitize.sh gpurun_out
