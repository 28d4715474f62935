"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys


def summarize(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<unnamed>::|void ", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"{'ms':>10} {'share':>6} {'launches':>8}  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v[1]/1e3:10.3f} {100*v[1]/tot:5.1f}% {v[0]:8d}  {k[:100]}")
    out.append(f"{tot/1e3:10.3f} 100.0% {sum(v[0] for v in agg.values()):8d}  TOTAL")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
