"""Turn `ncu -i <report> --page raw --csv` of the encoder-shaped GEMM launches of one timed bench step into
(a) profiles/<tag>_gemm_traffic.json -- DRAM bytes per launch and shape, read by bench.py for roofline.traffic --
and (b) a per-launch text summary.  Shapes are recognised from the bytes each launch moves (Y = M x N fp16 written,
A = M x K int8 read), M = utterances x 1500.
usage: python scripts/ncu_gemm_traffic.py raw.csv M out.json out.txt "<command line that produced the capture>" """
import csv
import json
import sys


def num(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")


def to_bytes(v, unit):
    return num(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(path, M, out_json, out_txt, cmd):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    shapes = [(N, K) for K in (384, 512, 768, 1024, 1280, 1536, 2048, 3072, 4096, 5120)
              for N in (384, 512, 768, 1024, 1152, 1280, 1536, 2048, 2304, 2560, 3072, 3840, 4096, 5120)]
    keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed"]
    per, text = {}, [cmd, f"{len(data)} launches; shapes recognised from bytes moved (M = {M})", ""]
    for r in data:
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        N, K = min(shapes, key=lambda s: abs(M * s[0] * 2 - wr) / (M * s[0] * 2) + abs(M * s[1] - rd) / (M * s[1]))
        key = f"{M}x{N}x{K}"
        per.setdefault(key, []).append(rd + wr)
        text.append(f"{r[col['Kernel Name']][:96]}   grid {r[col['Grid Size']]}   -> {key}")
        for k in keep:
            if k in col:
                text.append(f"    {k:78s} {units[col[k]]:16s} {r[col[k]]}")
        text.append("")
    out = {"source": cmd + "; dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches of a shape",
           "kind": "llmint8", "traffic_bytes_per_launch": {k: sum(v) / len(v) for k, v in sorted(per.items())},
           "launches": {k: len(v) for k, v in sorted(per.items())}}
    json.dump(out, open(out_json, "w"), indent=1)
    open(out_txt, "w").write("\n".join(text) + "\n")
    print(json.dumps(out["traffic_bytes_per_launch"], indent=1), out["launches"])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5])
