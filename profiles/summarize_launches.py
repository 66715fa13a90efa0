"""Summarise an `ncu --metrics ... --csv --log-file X.csv` launch list: per kernel (name, grid) count, time share and,
when present, DRAM / L2 bytes and tensor-pipe activity.   python profiles/summarize_launches.py X.csv [skip_first_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = L.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    d[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
ids = list(L.keys())[skip:]
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TS = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
agg = collections.OrderedDict()
tot_t = tot_b = tot_l2 = 0.0
for k in ids:
    d = L[k]
    key = (d["name"][:46], d["grid"])
    v, u = d["gpu__time_duration.sum"]
    t = v * TS.get(u, 1e-3)
    b = sum(d[m][0] * SC.get(d[m][1], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum") if m in d)
    l2 = d["lts__t_bytes.sum"][0] * SC.get(d["lts__t_bytes.sum"][1], 1) if "lts__t_bytes.sum" in d else 0.0
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", (0.0, ""))[0]
    a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += b; a[3] += tp; a[4] += l2
    tot_t += t; tot_b += b; tot_l2 += l2
print(f"# {sys.argv[1]}: {len(ids)} launches (first {skip} skipped), total {tot_t:.1f} us, DRAM {tot_b / 1e6:.1f} MB, L2 {tot_l2 / 1e6:.1f} MB")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:46s} grid={k[1]:15s} n={a[0]:5d} total={a[1]:9.1f}us avg={a[1] / a[0]:8.2f}us {100 * a[1] / tot_t:5.1f}%"
          f"  dram={a[2] / a[0] / 1e6:7.2f}MB  l2={a[4] / a[0] / 1e6:7.2f}MB  tensor_pipe={a[3] / a[0]:5.1f}%")
