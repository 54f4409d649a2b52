"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total and share of GPU
time, library (cuDNN / cuBLAS / ATen) vs libwhisperq.  usage: python scripts/launch_summary.py launches.csv out.txt "<command>" """
import collections
import csv
import sys


def main(path, out, cmd):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = r[col["Kernel Name"]]
        agg[name][0] += v
        agg[name][1] += 1
    tot = sum(v[0] for v in agg.values())
    own = sum(v[0] for k, v in agg.items() if "anonymous namespace" in k or "k_" in k.split("(")[0].split("::")[-1][:2])
    lines = [cmd, f"{sum(v[1] for v in agg.values())} launches, {tot / 1e3:.2f} ms of GPU time (ncu: serialised, cold clocks/caches -- "
             f"shares matter, not absolutes); libwhisperq kernels: {100 * own / tot:.1f} % of it", ""]
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        lines.append(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}% n={n:6d} avg={t / n:8.1f}us  {k[:120]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
