rm -f gpurun_out/parity_r1.jsonl
FABBER_PARITY_REPORT=gpurun_out/parity_r1.jsonl python -m pytest tests -m gpu -q 2>&1 | grep -E "AssertionError|passed|failed|Error|error" | cut -c1-800 > gpurun_out/pytest8.log; cat gpurun_out/pytest8.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
for w in c2 c3 c4 c5; do python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2>gpurun_out/bench_$w.err; python -c "
import json
d=json.load(open('gpurun_out/bench_$w.json')); print('$w value %.4g e2e %.4g inner %.4g frac %.3f ms %.3f cpu %.4g launches %d bad %d'%(d['value'], d['e2e']['value'], d['e2e']['inner_abi']['value'], d['roofline']['frac'], d['ms_per_step'], d['cpu_baseline']['value'], d['gpu_launches'], d['bad_voxels']))"; done
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
CMD3="python bench.py --workload c3 --voxels 2097152 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD3 > gpurun_out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vb_voxelwise -s 3 -c 1 -o gpurun_out/prof_c3b $CMD3 > gpurun_out/ncu_c3.log 2>&1
CMD4="python bench.py --workload c4 --voxels 2097152 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD4 > gpurun_out/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vb_voxelwise_ar -s 3 -c 1 -o gpurun_out/prof_c4 $CMD4 > gpurun_out/ncu_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
