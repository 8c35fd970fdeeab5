mkdir -p gpurun_out
RVDD_FUSE_MIN_ROWS=1 python tools/memcheck_case.py > gpurun_out/plain.log 2>&1 && RVDD_FUSE_MIN_ROWS=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/memcheck_case.py > gpurun_out/memcheck_r02.log 2>&1
echo rc=$?; tail -8 gpurun_out/memcheck_r02.log
