#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (memcheck on everything small, racecheck + synccheck on
# the shared-memory kernels).  gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
O=gpurun_out
mkdir -p $O
SEL='not config3 and not config4 and not full_size and not compiled_host'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/memcheck.log python -m pytest tests -m gpu -x -q -k "$SEL" > $O/memcheck_pytest.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck_pytest.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $O/racecheck.log python -m pytest tests -m gpu -x -q -k "paths and not config and not full_size" > $O/racecheck_pytest.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck_pytest.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $O/synccheck.log python -m pytest tests -m gpu -x -q -k "paths and not config and not full_size" > $O/synccheck_pytest.log 2>&1; echo "synccheck rc=$?" >> $O/synccheck_pytest.log
for f in memcheck racecheck synccheck; do tail -n 3 $O/$f.log; done
