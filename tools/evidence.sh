# Round evidence set (run on the GPU box): default bench line (fp16x3 headline + bf16 + eval frame
# + CPU baseline), the other BASELINE workloads, the train_step workloads, ncu launch list of one
# step, ncu --set full captures of the two tensor-core field kernels and of the camera kernel,
# block-0 timeline.  Every command reads /dev/null (nothing may wait on stdin on the box).
set -x
R=r02
exec </dev/null
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_default_1gpu.json 2> gpurun_out/${R}_bench_default_1gpu.err
tail -c 600 gpurun_out/${R}_bench_default_1gpu.json
# power / clock under a sustained load (5 s): nvidia-smi's reading needs more than a 10-step run
python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-also --no-parity > gpurun_out/${R}_bench_power_60steps.json 2>/dev/null
for w in quarterhd-train vrig-train fullhd-train; do
  python bench.py --workload $w --steps 5 --no-cpu-baseline > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err
done
for w in quarterhd-trainstep vrig-trainstep; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 1 > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_fp16x3_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-also --no-parity > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:field_x3 -s 3 -c 1 -f -o gpurun_out/${R}_prof_x3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-also --no-parity > gpurun_out/ncu_x3.log 2>&1
tail -2 gpurun_out/ncu_x3.log | cut -c1-200
# gpurun copies back at most 64 MiB: keep the text pages, not the reports
export_rep() {  # name [source]
  ncu -i gpurun_out/${R}_prof_$1.ncu-rep --page details > gpurun_out/${R}_$2_ncu_details.txt 2>/dev/null
  ncu -i gpurun_out/${R}_prof_$1.ncu-rep --page raw --csv > gpurun_out/${R}_$2_ncu_raw.csv 2>/dev/null
  if [ -n "$3" ]; then ncu -i gpurun_out/${R}_prof_$1.ncu-rep --page source --csv > gpurun_out/${R}_$2_ncu_source.csv 2>/dev/null; fi
  rm -f gpurun_out/${R}_prof_$1.ncu-rep
}
export_rep x3 field_x3 src
ncu --set full --clock-control none --import-source on -k regex:field_tc -s 3 -c 1 -f -o gpurun_out/${R}_prof_tc python bench.py --precision bf16 --steps 1 --warmup 1 --no-cpu-baseline --no-also --no-parity > gpurun_out/ncu_tc.log 2>&1
export_rep tc field_tc_bf16
ncu --set full --clock-control none -k regex:camera_rays -s 6 -c 2 -f -o gpurun_out/${R}_prof_camera python tools/bench_camera.py > gpurun_out/ncu_cam.log 2>&1
export_rep camera camera
python tools/bench_camera.py > gpurun_out/${R}_bench_camera.json 2>/dev/null
python tools/microbench_tmem_a.py > gpurun_out/${R}_microbench_tmem_a.json 2>/dev/null
[ -f nerfies_b200/_variants/libnfb_trace.so ] || python tools/build_variant.py trace -DNFB_TRACE >/dev/null 2>&1 || true
PREC=fp16x3 STEP_LO=0 STEP_HI=40 timeout 200 python tools/trace_tc.py > gpurun_out/${R}_x3_timeline.txt 2>&1
head -24 gpurun_out/${R}_x3_timeline.txt
