"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_*] --csv) of
tools/one_eval.py: per-kernel launches / time / DRAM traffic of the SECOND full-GP evaluation.

    python tools/launch_summary.py gpurun_out/launches3.csv profiles/r01_launch_summary_N10000.json
"""
import collections
import csv
import json
import re
import sys


def short(n):
    n = n.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    n = re.sub(r"GemmCfg<[^>]*>, ", "", n)
    n = n.replace("(bool)", "").replace("(int)", "")
    return re.sub(r"\(.*", "", n)


def main(src, dst):
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    by = collections.OrderedDict()
    for x in csv.DictReader(lines):
        d = by.setdefault(x["ID"], {"name": x["Kernel Name"], "grid": x["Grid Size"], "stream": x["Stream"]})
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        if x["Metric Name"].startswith("dram"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        else:
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)  # -> us
        d[x["Metric Name"]] = v
    L = list(by.values())
    idx = [i for i, d in enumerate(L) if "gram_sym" in d["name"]]
    seg = L[idx[1]: idx[1] + (idx[1] - idx[0])] if len(idx) > 1 else L[idx[0]:]
    agg = collections.OrderedDict()
    tot = 0.0
    for d in seg:
        n = short(d["name"])
        t = d["gpu__time_duration.sum"]
        a = agg.setdefault(n, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        tot += t
    gem = [d for d in seg if "gemm_tile" in d["name"] or "gemm_tma" in d["name"]]
    out = {
        "source": "ncu launch list of tools/one_eval.py (second full-GP evaluation); times are cold-cache and "
                  "serialised: compare shares, not absolutes",
        "all_kernels_ms_serialized": tot / 1e3,
        "gemm_launches": len(gem),
        "gemm_time_ms_serialized": sum(d["gpu__time_duration.sum"] for d in gem) / 1e3,
        "gemm_share_of_eval": sum(d["gpu__time_duration.sum"] for d in gem) / tot,
        "gemm_dram_bytes_per_eval": sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in gem),
        "per_kernel": {k: {"launches": v[0], "ms": v[1] / 1e3, "share": v[1] / tot, "dram_MB": v[2] / 1e6}
                       for k, v in agg.items()},
    }
    for k, v in out["per_kernel"].items():
        print("%-44s n=%4d %9.3f ms %5.1f%%  dram %9.1f MB" % (k[:44], v["launches"], v["ms"], 100 * v["share"], v["dram_MB"]))
    print("gemm: %d launches, %.2f ms (%.1f%% of %.2f ms), DRAM %.2f GB" % (
        out["gemm_launches"], out["gemm_time_ms_serialized"], 100 * out["gemm_share_of_eval"],
        out["all_kernels_ms_serialized"], out["gemm_dram_bytes_per_eval"] / 1e9))
    if dst:
        json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
