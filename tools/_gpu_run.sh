SAN_SEL="test_bdf_fixed_space or test_sens_incremental_rebuild_after_adapt" SAN_TOOLS=racecheck SAN_TIMEOUT=95 tools/sanitize.sh gpurun_out
head -12 gpurun_out/sanitizer_racecheck.log
tail -3 gpurun_out/sanitizer_racecheck.pytest.log
