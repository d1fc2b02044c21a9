# Round evidence set (run on the GPU box): bench line with CPU baseline, ncu launch
# list, one ncu --set full capture of the fine-level field kernel, block-0 timeline.
set -x
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_bf16_full.json 2> gpurun_out/bench_bf16_full.err
tail -c 700 gpurun_out/bench_bf16_full.json
NFB_DEBUG=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/benchline.py > gpurun_out/bench_skeleton_only.txt
cat gpurun_out/bench_skeleton_only.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bf16.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:field_tc -s 3 -c 1 -f -o gpurun_out/prof_tc8 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log | cut -c1-200
STEP_LO=0 STEP_HI=40 timeout 200 python tools/trace_tc.py > gpurun_out/tc_timeline.txt 2>&1
head -4 gpurun_out/tc_timeline.txt
