timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
tail -3 gpurun_out/bench_d.err
python -c "
import json;d=json.load(open('gpurun_out/bench_d.json'));print(d['value'],d['e2e']['value'],d['roofline']['stage_ms'],d['parity'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_azinv_flux -s 3 -c 1 -f -o gpurun_out/prof_flux_d python bench.py --batch 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_flux_d.ncu-rep 60 > gpurun_out/flux_d.txt 2>&1
grep -E "time_duration|bank_conflicts|wavefronts_mem_shared.sum |registers_per_thread |warps_active|fp64|shared_mem_per_block |dram__bytes_read.sum " gpurun_out/flux_d.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k 'c1_integrate or m2_batched or m4_batched or theta' 2>&1 | tail -4
