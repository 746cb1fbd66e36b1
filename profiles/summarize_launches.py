"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, mean us, share."""
import collections
import csv
import re
import sys


def summarize(path, min_share=0.003):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        per[name].append(v)
    tot = sum(sum(v) for v in per.values())
    out = [f"{'kernel':60s} {'n':>4s} {'mean_us':>9s} {'share':>6s}"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        if sum(v) / tot >= min_share:
            out.append(f"{k[:60]:60s} {len(v):4d} {sum(v) / len(v):9.1f} {sum(v) / tot:6.3f}")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
