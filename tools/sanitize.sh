#!/bin/bash
# compute-sanitizer over the small-size GPU parity tests (race / sync / memory checks of the
# last-CTA-done sink reduction, the single-CTA and cooperative BDF step kernels, k_expand_small, the prune kernels).
# Usage (on a GPU box):  tools/sanitize.sh [outdir]      logs -> <outdir>/sanitizer_<tool>.log
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SEL=${SAN_SEL:-'test_fspmat_jl_on_gpu or test_rectangular_telegraph_all_kernel_variants or test_device_resident_and_matvecadd or test_reference_kats or test_expand_delete_sequence_index_exact or test_sens_telegraph or test_sens_poisson or test_fixed_space_solve or test_adaptive_solve_reference_tests or test_bdf_fused or test_prune_by_mass_matches_oracle or test_telegraph_example'}
for tool in ${SAN_TOOLS:-memcheck synccheck racecheck}; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  NCME_SANITIZER=1 timeout ${SAN_TIMEOUT:-300} compute-sanitizer --tool $tool $extra --target-processes all --error-exitcode 0 \
      --log-file "$OUT/sanitizer_$tool.raw.log" \
      python -m pytest tests -x -q -m gpu -k "$SEL" -p no:cacheprovider > "$OUT/sanitizer_$tool.pytest.log" 2>&1
  echo "== $tool: exit $? ; $(tail -1 "$OUT/sanitizer_$tool.pytest.log")"
  # keep the verdict lines and the first reports only (raw logs can be huge)
  { grep -E "ERROR SUMMARY|RACECHECK SUMMARY|========= (Error|Warning|Race|Invalid|Uninitialized|Barrier|Program hit)" "$OUT/sanitizer_$tool.raw.log" | sort | uniq -c | sort -rn | head -40;
    echo "---- first 120 lines of the raw log"; head -120 "$OUT/sanitizer_$tool.raw.log"; } > "$OUT/sanitizer_$tool.log"
  rm -f "$OUT/sanitizer_$tool.raw.log"
done
