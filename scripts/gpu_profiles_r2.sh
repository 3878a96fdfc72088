#!/bin/bash
# Round-2 profile evidence (one B200): ncu --set full of one mid-volume launch group's seven kernels, the generic-power
# K1, DRAM traffic of all K1 launches of a volume, and the launch list of a short bench run.
R='sample_grids|vertex_kernel|quad_kernel|classify_count|span_scan|emit_lists|fixup'
ncu --set full --clock-control none --import-source on -k regex:"$R" -s 91 -c 7 -o gpurun_out/prof_r2_group \
    python scripts/gpu_step_target.py 2 fast serial > gpurun_out/prof_r2_group.log 2>&1
ncu -i gpurun_out/prof_r2_group.ncu-rep --page raw --csv > gpurun_out/prof_r2_group.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2_group.ncu-rep --page source --csv -k regex:sample_grids > gpurun_out/prof_r2_group.k1.source.csv 2>/dev/null
ncu -i gpurun_out/prof_r2_group.ncu-rep --page source --csv -k regex:vertex_kernel > gpurun_out/prof_r2_group.e3.source.csv 2>/dev/null
# generic power (config 4: P = 4, 32 iterations): one mid-volume K1 launch
ncu --set full --clock-control none --import-source on -k regex:sample_grids -s 13 -c 1 -o gpurun_out/prof_r2_generic \
    python scripts/gpu_step_target.py 2 fast serial 4 32 > gpurun_out/prof_r2_generic.log 2>&1
ncu -i gpurun_out/prof_r2_generic.ncu-rep --page raw --csv > gpurun_out/prof_r2_generic.raw.csv 2>/dev/null
# DRAM traffic of the nine K1 (+ nine fixup) launches of one volume
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"sample_grids|fixup" -s 18 -c 18 --csv \
    --log-file gpurun_out/k1_traffic_r2.csv python scripts/gpu_step_target.py 2 fast serial > /dev/null 2>&1
# launch list of a short bench run
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-strong --no-other > gpurun_out/launches_bench_r2.log 2>&1
for f in gpurun_out/prof_r2_group.ncu-rep gpurun_out/prof_r2_generic.ncu-rep; do
  sz=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$sz" -gt 20000000 ]; then rm -f $f; fi; done
du -sh gpurun_out; ls gpurun_out | tail -20
