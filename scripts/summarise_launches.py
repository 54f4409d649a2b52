"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot = collections.Counter()
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].replace('<unnamed>::', '').replace('void ', '')
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1.0)
        m = re.match(r'([A-Za-z0-9_:]+)', name)
        base = m.group(1) if m else name[:40]
        if base.startswith('at::'):
            f2 = re.search(r'(\w+Functor|\w+_kernel_cuda|\w+Op)\b', name)
            base = base + ':' + (f2.group(1) if f2 else '')
        if base.startswith('k_gemm_tc'):
            t = re.search(r'k_gemm_tc<([^>]*)>', name)
            base = 'k_gemm_tc<' + (t.group(1) if t else '') + '>'
        tot[base] += v
        cnt[base] += 1
    T = sum(tot.values())
    ours = sum(v for k, v in tot.items() if k.startswith('k_'))
    print(f"kernels {sum(cnt.values())}  GPU time {T / 1e6:.2f} ms  libwhisperq {ours / 1e6:.2f} ms ({100 * ours / T:.1f} %)")
    for k, v in tot.most_common(40):
        print(f"{v / 1e6:9.3f} ms {100 * v / T:5.1f}% n={cnt[k]:6d} avg={v / cnt[k] / 1e3:8.2f}us  {k}")


if __name__ == '__main__':
    main(sys.argv[1])
