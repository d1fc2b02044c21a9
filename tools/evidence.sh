set -x
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_bf16_full.json 2> gpurun_out/bench_bf16_full.err
tail -c 600 gpurun_out/bench_bf16_full.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bf16.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
# fine-level field kernel = 2nd field_tc launch of a step; skip warm-up step's two
ncu --set full --clock-control none --import-source on -k regex:field_tc -s 3 -c 1 -f -o gpurun_out/prof_tc7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
STEP_LO=0 STEP_HI=40 timeout 200 python tools/trace_tc.py > gpurun_out/tc_timeline.txt 2>&1
head -8 gpurun_out/tc_timeline.txt
