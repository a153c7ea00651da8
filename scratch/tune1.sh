for rep in 1 2; do
for v in mb1 mb3 mb4; do
  for w in c2 c3; do
    FABBER_CUDA_LIB=scratch/lib_$v.so python bench.py --workload $w --voxels 4194304 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v $w value %.4g frac %.3f ms %.3f its %.2f'%(d['value'], d['roofline']['frac'], d['ms_per_step'], d['iterations_per_voxel']))"
  done
done
done
FABBER_CUDA_LIB=scratch/lib_mb4.so python -m pytest tests -m gpu -q -k "c2 or c3 or c1_linear" 2>&1 | tail -3
