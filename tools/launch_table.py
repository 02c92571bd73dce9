"""Per-kernel table (launches, time, share) of an ncu launch list (--metrics gpu__time_duration.sum --csv), all launches.

    python tools/launch_table.py gpurun_out/x.csv "title" > profiles/r02_x_launches.txt
"""
import collections, csv, re, sys


def short(n):
    n = n.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    n = re.sub(r"GemmCfg<[^>]*>, ", "", n).replace("(bool)", "").replace("(int)", "")
    return re.sub(r"\(.*", "", n)


src, title = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
agg = collections.OrderedDict()
tot = 0.0
for x in csv.DictReader(lines):
    if x["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(x["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(x["Metric Unit"], 1e-6)
    a = agg.setdefault(short(x["Kernel Name"]), [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(title)
print("serialised kernel time %.2f ms (cold-cache, one launch at a time: read the shares)" % tot)
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-52s %4d launches %9.3f ms %5.1f%%" % (n[:52], c, t, 100 * t / tot))
