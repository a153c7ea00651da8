#!/bin/bash
# ncu captures of the round's top kernels (run on a GPU box; each command only after the plain bench ran clean):
#   tools/ncu_round2.sh c2|c4|c5   ->  gpurun_out/r2n_<w>.ncu-rep   (read with tools/ncu_raw.py + tools/ncu_src.py)
set -x
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -c 1"
for w in "$@"; do
  case $w in
    c2) k=vb_voxelwise_white;;
    c3) k=vb_voxelwise_white;;
    c4) k=vb_voxelwise_ar_kernel;;
    c5) k=sp_noise_kernel;;
  esac
  timeout 900 $NCU -k regex:$k -o gpurun_out/r2n_$w python bench.py --workload $w --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1
  tail -2 gpurun_out/ncu_$w.log
done
if [[ " $* " == *" c5 "* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/r2n_launches_c5.csv python bench.py --workload c5 --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c5l.log 2>&1
fi
