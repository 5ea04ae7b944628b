python -m pytest tests -q -m gpu > gpurun_out/r02i_tests.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r02i_tests.log
python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo bench rc=$?; cut -c1-200 gpurun_out/r02i_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02i_bench_reference.json 2> gpurun_out/r02i_bench_reference.err; echo ref rc=$?; cut -c1-300 gpurun_out/r02i_bench_reference.json
bash profiles/make_profiles.sh r02i 512 > gpurun_out/r02i_make_profiles.log 2>&1; echo profiles rc=$?; cat gpurun_out/r02i_summary.txt
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_multi_signal.py tests/test_gpu_parity.py -q -m gpu -x -k "two_signals or everywhere_star or m2_batched_pipeline or m3_cst or omission" > gpurun_out/r02i_memcheck.log 2>&1; echo memcheck rc=$?) 
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02i_memcheck.log | tail -3
(timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_multi_signal.py -q -m gpu -x -k "two_signals" > gpurun_out/r02i_racecheck.log 2>&1; echo racecheck rc=$?)
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02i_racecheck.log | tail -3
