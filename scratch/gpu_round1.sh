set -x
rm -f gpurun_out/parity_r1.jsonl
FABBER_PARITY_REPORT=gpurun_out/parity_r1.jsonl python -m pytest tests -m gpu -q 2>&1 | grep -E "AssertionError|passed|failed|Error" | cut -c1-900 > gpurun_out/pytest5.log; cat gpurun_out/pytest5.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_c2.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_c2.log 2>&1
CMD3="python bench.py --workload c3 --voxels 2097152 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD3 > gpurun_out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vb_voxelwise -s 3 -c 1 -o gpurun_out/prof_c3 $CMD3 > gpurun_out/ncu_c3.log 2>&1
$CMD > gpurun_out/plain_c2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vb_voxelwise -s 3 -c 1 -o gpurun_out/prof_c2 $CMD > gpurun_out/ncu_c2b.log 2>&1
ls -la gpurun_out
