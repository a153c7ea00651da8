#!/bin/bash
# compute-sanitizer over the spatial / slab / two-echo kernels (run on a GPU box): logs into gpurun_out/
# usage: tools/sanitize.sh        (about 10 minutes; every leg is bounded by its own timeout)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { # name tool timeout pytest-args...
  name=$1; tool=$2; to=$3; shift 3
  timeout $to $S --tool $tool --error-exitcode 86 --log-file gpurun_out/sanitizer_${name}.log \
    python -m pytest "$@" -x -q -m gpu -p no:cacheprovider > gpurun_out/sanitizer_${name}.pytest.log 2>&1
  echo "$name: exit $? ; $(grep -c 'ERROR SUMMARY' gpurun_out/sanitizer_${name}.log) summaries; $(tail -1 gpurun_out/sanitizer_${name}.log)"
  tail -2 gpurun_out/sanitizer_${name}.pytest.log
}
run memcheck_spatial memcheck 600 tests/test_gpu_spatial.py -k "golden or irregular or struck"
run racecheck_spatial racecheck 600 tests/test_gpu_spatial.py -k "golden or irregular"
run memcheck_slabs memcheck 600 tests/test_gpu_spatial_multi.py -k "one_part or mrf_slabs"
run racecheck_slabs racecheck 600 tests/test_gpu_spatial_multi.py -k "one_part or mrf_slabs"
run memcheck_ar2 memcheck 400 tests/test_gpu_ar2.py -k "poly or even_series"
run memcheck_voxelwise memcheck 400 tests/test_gpu_voxelwise.py -k "golden or poly_ar1 or tile_the_volume"
