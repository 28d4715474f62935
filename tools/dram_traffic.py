"""Per-kernel DRAM traffic from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list
-> profiles/r02_dram_traffic.json (bench.py reads it for roofline.traffic):  python tools/dram_traffic.py launches.csv out.json"""
import csv, json, re, sys, collections


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    return re.sub(r"[<(].*", "", name)


def main(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    idx = {k: i for i, k in enumerate(rows[h])}
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for r in rows[h + 1:]:
        if len(r) < len(rows[h]):
            continue
        k = short(r[idx["Kernel Name"]])
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        per[k][r[idx["Metric Name"]]] += v * scale
        ids[k].add(r[idx["ID"]])
    res = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over " + path.split("/")[-1]}
    for k, m in per.items():
        n = len(ids[k])
        res[k] = {"launches": n, "dram_bytes_per_launch": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / n,
                  "dram_read_bytes_per_launch": m["dram__bytes_read.sum"] / n, "dram_write_bytes_per_launch": m["dram__bytes_write.sum"] / n,
                  "avg_ns": m["gpu__time_duration.sum"] / n}
    json.dump(res, open(out, "w"), indent=1)
    for k, v in sorted(res.items(), key=lambda kv: -(kv[1]["avg_ns"] * kv[1]["launches"]) if isinstance(kv[1], dict) else 0)[:12]:
        if isinstance(v, dict):
            print(f"{k:44s} n={v['launches']:5d} {v['dram_bytes_per_launch'] / 1e6:9.1f} MB/launch {v['avg_ns'] / 1e3:9.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
