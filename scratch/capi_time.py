import sys, time; sys.path.insert(0,".")
import numpy as np, ctypes as C
from fabber_core_b200 import fabber as fab, synth
side=128; T=64; n=side**3
y = synth.poly_volume(n, T, 3, seed=1).numpy().reshape(-1)
f = fab.Fabber()
f._set_options({"model":"poly","degree":3,"noise":"white","method":"vb","save-mean":True,"save-std":True,"save-noise-mean":True})
f._trycall(f.clib.fabber_set_extent, f.handle, side, side, side, np.ones(n, dtype=np.int32), f.errbuf)
out = np.empty(n, dtype=np.float32)
for rep in range(3):
    t0=time.perf_counter()
    f._trycall(f.clib.fabber_set_data, f.handle, b"data", T, y, f.errbuf); t1=time.perf_counter()
    f._trycall(f.clib.fabber_dorun, f.handle, len(f.outbuf), f.outbuf, f.errbuf, f.progress_cb_type(0)); t2=time.perf_counter()
    for k in ["mean_c0","mean_c1","mean_c2","mean_c3","std_c0","std_c1","std_c2","std_c3","noise_means"]:
        f._trycall(f.clib.fabber_get_data, f.handle, k.encode(), out, f.errbuf)
    t3=time.perf_counter()
    print("rep", rep, "set_data %.1f ms dorun %.1f ms get_data %.1f ms" % ((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3))
    print("\n".join(l for l in f.outbuf.value.decode().splitlines() if "timing" in l))
