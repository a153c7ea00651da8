"""Summarise an ncu --page source --csv export: executed instruction mix by opcode, stall reasons,
and the hottest instructions. usage: python scratch/ncu_src.py file.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
by_op = collections.Counter(); samp_op = collections.Counter()
for r in data:
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[0] if toks and not toks[0].startswith("@") else (toks[1] if len(toks) > 1 else "?")
    op = op.split(".")[0]
    by_op[op] += f(r, "Instructions Executed"); samp_op[op] += f(r, "# Samples")
print("total warp-instructions executed %.4g, samples %d" % (tot_inst, tot_samp))
print("opcode mix (share of executed warp instructions | share of stall samples):")
for op, n in by_op.most_common(18):
    print("  %-10s %6.2f%% | %6.2f%%" % (op, 100 * n / tot_inst, 100 * samp_op[op] / max(tot_samp, 1)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(f(r, s) for r in data) for s in stalls}
allst = sum(tot.values())
print("stall reasons:", ", ".join("%s %.1f%%" % (s[6:], 100 * v / allst) for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
print("hottest instructions (samples):")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
    print("  %6d  %s" % (f(r, "# Samples"), r[ix["Source"]].strip()[:100]))
