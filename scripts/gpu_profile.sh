#!/bin/bash
# ncu --set full of one launch group's seven kernels (second volume of scripts/gpu_step_target.py), exported to
# CSV on the GPU box; the .ncu-rep is kept only if it fits gpurun_out's size limit.
# usage: scripts/gpu_profile.sh <tag> [extra args of gpu_step_target.py after the rep count]
tag=${1:-prof}; shift
out=gpurun_out/$tag
ncu --set full --clock-control none --import-source on \
    -k regex:"sample_grids|vertex_kernel|quad_kernel|classify|apply_prefix|fixup|scan_chunks|span_scan" -s ${SKIP:-91} -c 7 \
    -o $out python scripts/gpu_step_target.py 2 "$@" > $out.log 2>&1
tail -2 $out.log
ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
ncu -i $out.ncu-rep --page source --csv -k regex:sample_grids > $out.k1.source.csv 2>/dev/null
ncu -i $out.ncu-rep --page source --csv -k regex:vertex_kernel > $out.e3.source.csv 2>/dev/null
ls -la gpurun_out | tail -8
sz=$(stat -c %s $out.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm -f $out.ncu-rep; fi
