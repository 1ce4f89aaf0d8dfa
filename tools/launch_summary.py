"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) -> per-kernel summary CSV.
usage: python tools/launch_summary.py gpurun_out/launches.csv profiles/rNN_bench_launches_summary.csv"""
import collections, csv, io, sys


def main(src, dst):
    txt = open(src).read().splitlines()
    start = [i for i, l in enumerate(txt) if l.startswith('"ID"')][0]
    agg = collections.OrderedDict()
    for x in csv.DictReader(txt[start:]):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v, u = float(x["Metric Value"].replace(",", "")), x["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        agg.setdefault((x["Kernel Name"], x["Grid Size"] if "mapper" in x["Kernel Name"] or "knn2_tc" in x["Kernel Name"] else ""), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out = io.StringIO()
    w = csv.writer(out)
    w.writerow(["kernel", "grid", "launches", "mean_us", "min_us", "max_us", "total_us", "share"])
    for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        if len(v) < 2 and sum(v) / tot < 0.002:
            continue
        w.writerow([k[:90], g, len(v), round(sum(v) / len(v), 2), round(min(v), 2), round(max(v), 2), round(sum(v), 1), round(sum(v) / tot, 3)])
    open(dst, "w").write(out.getvalue())
    print(out.getvalue())


if __name__ == "__main__":
    main(*sys.argv[1:3])
