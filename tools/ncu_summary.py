"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit, val = r.get("Metric Unit", "ns"), float(r["Metric Value"].replace(",", ""))
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns))
    agg = OrderedDict()
    for n, ns in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    lines = [f"launches: {len(rows)}   total device time: {tot / 1e6:.3f} ms (cold-cache, serialised: compare SHARES)", "",
             "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for n, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{n}` | {cnt} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
