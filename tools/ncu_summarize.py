"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid)."""
import csv, re, sys, collections

def main(path, header=""):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*$", "", name)
        name = name.replace("(anonymous namespace)::", "").replace("tc::", "")
        rows.append((name, r["Grid Size"], float(r["Metric Value"].replace(",", "")) / 1e3))
    agg = collections.OrderedDict()
    for n, g, us in rows:
        k = (n, g)
        a = agg.setdefault(k, [0.0, 0])
        a[0] += us; a[1] += 1
    tot = sum(a[0] for a in agg.values())
    print(header, end="")
    print(f"# total kernel time {tot/1e3:.2f} ms over {len(rows)} launches (cold-cache, serialised: compare SHARES)")
    print(f"{'us':>10s} {'share':>6s} {'n':>5s} {'avg us':>8s}  kernel grid")
    for (n, g), (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{us:10.1f} {100*us/tot:5.1f}% {c:5d} {us/c:8.2f}  {n} {g}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2].encode().decode("unicode_escape") if len(sys.argv) > 2 else "")
