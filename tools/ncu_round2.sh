set -x
python bench.py > gpurun_out/r2n_bench_1gpu.json 2> gpurun_out/r2n_bench_1gpu.err; tail -c 600 gpurun_out/r2n_bench_1gpu.json
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -c 1"
timeout 600 $NCU -k regex:vb_voxelwise_white -o gpurun_out/r2n_c2 python bench.py --workload c2 --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1; tail -2 gpurun_out/ncu_c2.log
timeout 900 $NCU -k regex:vb_voxelwise_ar_kernel -o gpurun_out/r2n_c4 python bench.py --workload c4 --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
timeout 900 $NCU -k regex:sp_noise_kernel -o gpurun_out/r2n_c5_noise python bench.py --workload c5 --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c5n.log 2>&1; tail -2 gpurun_out/ncu_c5n.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/r2n_launches_c5.csv python bench.py --workload c5 --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c5l.log 2>&1; tail -2 gpurun_out/ncu_c5l.log
ls -la gpurun_out/*.ncu-rep
