"""Top stall sites from `ncu -i rep --page source --csv --kernel-name regex:X` (SASS view).
usage: python tools/ncu_source_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
tot_inst = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[idx["# Samples"]] or 0)
        ie = int(r[idx["Instructions Executed"]] or 0)
    except ValueError:
        continue
    tot += s
    tot_inst += ie
    data.append((s, ie, r))
agg = {h: 0 for h in stalls}
for s, ie, r in data:
    for h in stalls:
        try:
            agg[h] += int(r[idx[h]] or 0)
        except ValueError:
            pass
print("total samples %d, warp-instructions %d" % (tot, tot_inst))
print("stall mix: " + ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(tot, 1)) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for i, (s, ie, r) in enumerate(data):
    r.append(i)
for s, ie, r in sorted(data, key=lambda t: -t[0])[:N]:
    top = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print("%5.1f%% line %4d inst %8d  %-70s %s" % (100.0 * s / max(tot, 1), r[-1], ie, r[idx["Source"]][:70],
                                                 " ".join("%s:%d" % (n, v) for v, n in top if v)))
