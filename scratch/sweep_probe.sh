for c in 1 2; do
CMD="python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline"
FABBER_SWEEP_CTAS_PER_SM=$c FABBER_CUDA_LIB=scratch/lib_sub.so ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread --clock-control none -c 24 --csv --log-file gpurun_out/launches_sweep_$c.csv $CMD > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_sweep_$c.csv')))
hdr=None
for r in rows:
    if r and r[0]=="ID": hdr=r; ix={h:i for i,h in enumerate(hdr)}; continue
    if hdr and len(r)==len(hdr) and r[ix["Metric Name"]]=="gpu__time_duration.sum":
        print("ctas $c", r[ix["Kernel Name"]][:40], r[ix["Grid Size"]], r[ix["Block Size"]], float(r[ix["Metric Value"]])/1e6, "ms")
PY
done
