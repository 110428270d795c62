"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and (optionally) the first N launches."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("grove::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v if unit == "us" else v * 1e3 if unit == "ms" else v * 1e6
        out.append((name, v, row["Grid Size"]))
    return out


if __name__ == "__main__":
    rows = load(sys.argv[1])
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for name, v, _ in rows:
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T / 1e3:.3f} ms over {len(rows)} launches")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v / 1e3:8.3f} ms {100 * v / T:5.1f}% x{cnt[k]:3d}  {k[:100]}")
    for i, (name, v, g) in enumerate(rows[:n]):
        print(i, f"{v:9.1f} us", name[:60], g)
