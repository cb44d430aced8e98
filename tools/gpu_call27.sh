# round 2, call 27: compute-sanitizer over the host-packed upload (k_unpack_ids_blk) and the coalesced path
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_hostpack.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r02af_compute_sanitizer_memcheck_hostpack.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02af_compute_sanitizer_memcheck_hostpack.log
tail -6 gpurun_out/r02af_compute_sanitizer_memcheck_hostpack.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_hostpack.py -x -q -m gpu -p no:cacheprovider -k "equals_plain and dlrm_cat" > gpurun_out/r02af_compute_sanitizer_racecheck_hostpack.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02af_compute_sanitizer_racecheck_hostpack.log
tail -6 gpurun_out/r02af_compute_sanitizer_racecheck_hostpack.log
