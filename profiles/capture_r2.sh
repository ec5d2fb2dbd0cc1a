#!/bin/bash
# End-of-round evidence, one gpurun call (1 GPU):  gpurun --timeout 2400 -- bash profiles/capture_r2.sh
# Everything lands in gpurun_out/; profiles/summarize_ncu.py and profiles/line_profile.py condense it afterwards.
set -x
mkdir -p gpurun_out
# 1. sanitizer on the smoke path (K0 spec/verify, K1a/c/d with the acquire/release upsert, K2, K3)
compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_memcheck.txt 2>&1
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_racecheck.txt 2>&1
tail -4 gpurun_out/r2_memcheck.txt gpurun_out/r2_racecheck.txt
# 2. launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
# 3. full captures: one launch of every hot kernel on the 25 M-read workload (first full step after one warm-up step)
ncu --set full --clock-control none --import-source on -k regex:'k0_parse_spec|k1_trim_collapse|k1_general_search|k1_slow_search' -s 4 -c 4 -o gpurun_out/r2_k0k1 -f python profiles/kernel_variants.py 25000000 0 > gpurun_out/r2_k0k1_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k2_round' -s 9 -c 9 -o gpurun_out/r2_k2 -f python profiles/kernel_variants.py 25000000 "" 0 > gpurun_out/r2_k2_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
# 3b. DRAM bytes per family from those captures: bench.py reports them as roofline.traffic
python profiles/make_traffic.py gpurun_out/r2_k0k1.ncu-rep gpurun_out/r2_k2.ncu-rep && cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
# 4. the bench line itself (not under a profiler)
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_final.err
tail -c 600 gpurun_out/r2_bench_final.json
