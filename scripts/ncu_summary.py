"""Summarise `ncu -i <report> --page raw --csv` into a short per-launch text table (kernel, grid, duration, DRAM bytes,
DRAM / tensor-pipe / issue utilisation, L2 hit rate, registers, shared memory) and, with --traffic-json, the DRAM bytes
per launch of the encoder-shaped GEMMs keyed MxNxK (read by bench.py for roofline.traffic).
usage: python scripts/ncu_summary.py raw.csv out.txt "<command>" [--traffic-json out.json --M 384000]"""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def main(argv):
    path, out_txt, cmd = argv[1], argv[2], argv[3]
    tj = argv[argv.index("--traffic-json") + 1] if "--traffic-json" in argv else None
    M = int(argv[argv.index("--M") + 1]) if "--M" in argv else 0
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    text = [cmd, f"{len(data)} launches", ""]
    per = {}
    shapes = [(N, K) for K in (384, 512, 768, 1024, 1280, 1536, 2048, 3072, 4096, 5120)
              for N in (384, 512, 768, 1024, 1152, 1280, 1536, 2048, 2304, 2560, 3072, 3840, 4096, 5120)]
    for r in data:
        name = r[col["Kernel Name"]]
        text.append(f"{name[:110]}   grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for k in KEEP:
            if k in col:
                text.append(f"    {k:72s} {units[col[k]]:14s} {r[col[k]]}")
        if tj and "k_gemm_tc" in name and "dram__bytes_read.sum" in col:
            rd = num(r[col["dram__bytes_read.sum"]]) * UNIT.get(units[col["dram__bytes_read.sum"]], 1.0)
            wr = num(r[col["dram__bytes_write.sum"]]) * UNIT.get(units[col["dram__bytes_write.sum"]], 1.0)
            if wr > 0.5 * M * 384 * 2:           # encoder-shaped: writes a [M, N] fp16 matrix
                # reads: the int8 A matrix, plus the fp16 residual [M, N] when the launch fuses the residual add (fc2)
                cands = [(N, K, res) for (N, K) in shapes for res in (0, 1)]
                N, K, res = min(cands, key=lambda s: abs(M * s[0] * 2 - wr) / (M * s[0] * 2)
                                + abs(M * s[1] + s[2] * M * s[0] * 2 - rd) / (M * s[1] + s[2] * M * s[0] * 2))
                per.setdefault(f"{M}x{N}x{K}", []).append(rd + wr)
                text.append(f"    -> recognised as {M}x{N}x{K}" + (" (+ fused residual read)" if res else ""))
        text.append("")
    open(out_txt, "w").write("\n".join(text) + "\n")
    if tj:
        json.dump({"source": cmd + "; dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches of a shape",
                   "kind": "llmint8", "traffic_bytes_per_launch": {k: sum(v) / len(v) for k, v in sorted(per.items())},
                   "launches": {k: len(v) for k, v in sorted(per.items())}}, open(tj, "w"), indent=1)
    print(f"{len(data)} launches summarised into {out_txt}")


if __name__ == "__main__":
    main(sys.argv)
