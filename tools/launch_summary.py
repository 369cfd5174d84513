"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel shares and one step's sequence."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        rows.append((re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1], v))
    return rows


if __name__ == "__main__":
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.0f} us, {len(rows)} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
        print(f"{100 * v[1] / tot:5.1f}% {v[1]:9.1f} us n={v[0]:3d} avg {v[1] / v[0]:8.1f}  {k[:70]}")
    if len(sys.argv) > 3:
        idx = [i for i, (k, v) in enumerate(rows) if k == "class_slots_kernel"]
        print([(k[:14], round(v)) for k, v in rows[idx[0]:idx[0] + 19] if v > 20])
        j = [i for i, (k, v) in enumerate(rows) if k == "attn_bwd_prep_kernel"]
        print([(k[:14], round(v)) for k, v in rows[j[0]:j[0] + 26] if v > 20])
