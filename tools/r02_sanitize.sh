# compute-sanitizer over the new round-2 device code (small inputs): memcheck + racecheck + synccheck of smoke(), memcheck of a
# slice of the parity tests (balanced forward kernel, window walk/packing, compaction, chains)
mkdir -p gpurun_out
T=${1:-r02s}
CS=/usr/local/cuda/bin/compute-sanitizer
( timeout 600 $CS --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke ) > gpurun_out/${T}_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/${T}_memcheck_smoke.log
( timeout 900 $CS --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke ) > gpurun_out/${T}_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 gpurun_out/${T}_racecheck_smoke.log
( timeout 600 $CS --tool synccheck --error-exitcode 9 python __graft_entry__.py smoke ) > gpurun_out/${T}_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?"; tail -3 gpurun_out/${T}_synccheck_smoke.log
( timeout 1200 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py -q -x -k "forward_level_sweep_both or overwritten or remove_null or a_window_deeper or chain_matches or test_random_tapes" ) > gpurun_out/${T}_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 gpurun_out/${T}_memcheck_tests.log
