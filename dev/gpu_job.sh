#!/bin/bash
# One GPU-box job that regenerates every artefact of a profile tag:  bash dev/gpu_job.sh r02k
TAG=${1:-r02k}
python -m pytest tests -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/${TAG}_tests.log
bash profiles/make_profiles.sh ${TAG} 512 > gpurun_out/${TAG}_make_profiles.log 2>&1; echo profiles rc=$?; cat gpurun_out/${TAG}_summary.txt
cp gpurun_out/${TAG}_kernels.json profiles/ 2>/dev/null        # so that the bench below reads this tag's DRAM traffic
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench rc=$?; cut -c1-200 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo ref rc=$?
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_multi_signal.py tests/test_gpu_parity.py tests/test_sweep_and_limits.py -q -m gpu -x -k "two_signals or everywhere_star or m2_batched_pipeline or m3_cst or omission or handed_back" > gpurun_out/${TAG}_memcheck.log 2>&1; echo memcheck rc=$?)
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log | tail -3
(timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_multi_signal.py -q -m gpu -x -k "two_signals" > gpurun_out/${TAG}_racecheck.log 2>&1; echo racecheck rc=$?)
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_racecheck.log | tail -3
(timeout 600 compute-sanitizer --tool initcheck --print-limit 300 --error-exitcode 9 python -m pytest tests/test_multi_signal.py tests/test_gpu_parity.py tests/test_sweep_and_limits.py -q -m gpu -x -k "two_signals or everywhere_star or m2_batched_pipeline or m3_cst or omission or handed_back" > gpurun_out/${TAG}_initcheck.log 2>&1; echo initcheck rc=$?)
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_initcheck.log | tail -3
(timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_multi_signal.py -q -m gpu -x -k "two_signals" > gpurun_out/${TAG}_synccheck.log 2>&1; echo synccheck rc=$?)
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_synccheck.log | tail -3
python dev/latency.py 2>&1 | tail -6 > gpurun_out/${TAG}_latency.log; cat gpurun_out/${TAG}_latency.log
