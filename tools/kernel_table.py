"""Turn an `ncu --csv --metrics ...` log into a per-kernel table: launches, time, DRAM GB/s, tensor-pipe %.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv \\
        --log-file gpurun_out/k.csv python <workload>
    python tools/kernel_table.py gpurun_out/k.csv [skip_first_n_launches]
"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]
ki, mi, vi, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    per.setdefault(int(r[ii]), {"name": r[ki]})[r[mi]] = v
agg = collections.OrderedDict()
for lid, d in per.items():
    if lid < skip:
        continue
    name = d["name"].split("(")[0].replace("void ", "").replace("up::", "")
    a = agg.setdefault(name, {"n": 0, "ns": 0.0, "bytes": 0.0, "tc_ns": 0.0})
    ns = d.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1
    a["ns"] += ns
    a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["tc_ns"] += ns * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                             d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)) / 100.0
tot = sum(a["ns"] for a in agg.values())
print("%-44s %6s %10s %7s %10s %9s" % ("kernel", "n", "time us", "share", "DRAM GB/s", "tensor %"))
for name, a in sorted(agg.items(), key=lambda x: -x[1]["ns"]):
    print("%-44s %6d %10.1f %6.1f%% %10.0f %9.1f" % (name[:44], a["n"], a["ns"] / 1e3, 100 * a["ns"] / tot,
                                                  a["bytes"] / max(a["ns"], 1.0), 100 * a["tc_ns"] / max(a["ns"], 1.0)))
print("total %.1f us over %d launches (serialised, cold L2 per launch: shares, not absolutes)" % (tot / 1e3, sum(a["n"] for a in agg.values())))
