#!/usr/bin/env python
"""Per-kernel table of one FCT step from an ncu launch list:

    FCT_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/kernels_step.csv python tools/kprof_step.py
    python tools/kernels_table.py gpurun_out/kernels_step.csv profiles/r2_kernels.json [peak GB/s]

(read here on the CPU box).  Writes {source, peak_GBs, kernels: {name: {launches, ms_per_launch, dram_bytes_per_launch, GBs,
pct_of_peak, share_of_time}}} -- bench.py embeds it as the `kernels` table and takes `roofline.traffic` from it.  ncu
serialises launches and starts each one with a cold cache, so GB/s here is a lower bound for the kernel inside a step;
skipped (already converged) Jacobi launches are listed separately."""
import collections
import csv
import json
import os
import re
import sys


def main(path, out, peak):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(float))      # (launch id) -> metric -> value
    names = {}
    for row in csv.DictReader(lines):
        mid = row["ID"]
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").strip()
        name = re.sub(r"<.*", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        m = row["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v * 1e3 if u in ("s", "second") else v
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            v *= mult
        per[mid][m] = v
        names[mid] = name
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for mid, mets in per.items():
        name = names[mid]
        ms = mets.get("gpu__time_duration.sum", 0.0)
        if ((name.startswith("k_jacobi") and "sweep" in name) or name == "k_tile") and ms < 0.02:
            name += " (skipped: converged)"
        a = agg[name]
        a[0] += 1
        a[1] += ms
        a[2] += mets.get("dram__bytes_read.sum", 0.0) + mets.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    table = {}
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        ms, by = a[1] / a[0], a[2] / a[0]
        gbs = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        table[k] = {"launches": a[0], "ms_per_launch": round(ms, 5), "dram_bytes_per_launch": int(by), "GBs": round(gbs, 1),
                    "pct_of_peak": round(100 * gbs / peak, 1), "share_of_time": round(100 * a[1] / tot, 1)}
    res = {"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                     f"FCT_NO_GRAPH=1, tools/kprof_step.py (one state + one adjoint FCT step + 2 gradient slices + cost, 4097^2 DoF); "
                     f"launches are serialised and cold-cache under ncu: compare shares, GB/s is a lower bound",
           "peak_GBs": peak, "total_kernel_ms": round(tot, 3), "kernels": table}
    json.dump(res, open(out, "w"), indent=1)
    print(f"total kernel time {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches")
    for k, r in table.items():
        print(f"{k:44s} n={r['launches']:4d} {r['ms_per_launch']:8.4f} ms {r['dram_bytes_per_launch'] / 1e9:7.3f} GB "
              f"{r['GBs']:7.0f} GB/s {r['pct_of_peak']:5.1f}% of peak  share {r['share_of_time']:5.1f}%")


if __name__ == "__main__":
    pk = float(sys.argv[3]) if len(sys.argv) > 3 else None
    if pk is None:
        try:
            pk = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                   "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:  # noqa: BLE001
            pk = 6650.0
    main(sys.argv[1], sys.argv[2], pk)
