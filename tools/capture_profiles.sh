#!/bin/bash
# ncu evidence for profiles/ (B200_PROFILING.md recipe): run on the GPU box, e.g.  gpurun --timeout 1800 -- 'tools/capture_profiles.sh r2'
# Every capture is taken only after the same command exited 0 without ncu.  Outputs land in gpurun_out/.
tag=${1:-r2}
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-aux --sustain-s 0"
$B > gpurun_out/pre_$tag.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:avse_ -c 400 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:avse_forward4 -c 1 -f -o gpurun_out/fwd_$tag $B --no-inverse > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"avse_inverse8|avse_mel_to_coef" -c 2 -f -o gpurun_out/inv_$tag $B > gpurun_out/ncu_i.log 2>&1
ncu --set full --clock-control none -k regex:"avse_snr_factor|avse_floor_inplace" -c 2 -f -o gpurun_out/aux_$tag $B --no-inverse > gpurun_out/ncu_a.log 2>&1
ls -la gpurun_out/*_$tag.ncu-rep gpurun_out/launches_$tag.csv
