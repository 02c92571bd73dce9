"""Brief text summary of an .ncu-rep (run where ncu is installed; no GPU needed):

    python tools/ncu_brief.py gpurun_out/x.ncu-rep profiles/r02_ncu_x.txt

One block per profiled launch: duration, DRAM traffic and rate, pipe utilisation, occupancy limits, and the
warp-stall mix aggregated from the per-instruction samples of the source page."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit: registers"), ("launch__occupancy_limit_shared_mem", "CTAs/SM limit: smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "FP64+DMMA shared pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "  of which DFMA-class (fp64 pipe) %"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "  of which DMMA sub-pipe %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    stalls, cur, h = [], None, None
    names = []
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {}
            stalls.append(cur)
            names.append(r[1])
        elif r and r[0] == "Address":
            h = r
        elif cur is not None and h is not None and len(r) > 10:
            for i, c in enumerate(h):
                if c.startswith("stall_") and "Not Issued" not in c and r[i]:
                    cur[c] = cur.get(c, 0) + int(r[i])
    # the source page may list a launch more than once (several views): match the blocks to the raw rows by kernel
    # name, in order, skipping repeated views of the same launch
    def norm(n):
        return "".join(ch for ch in n.replace("(int)", "").replace("(bool)", "") if not ch.isspace())
    by_name = {}
    for n, st in zip(names, stalls):
        lst = by_name.setdefault(norm(n), [])
        if not lst or lst[-1] != st:
            lst.append(st)
    counts = {}
    for r in data:
        counts[norm(r[hdr.index("Kernel Name")])] = counts.get(norm(r[hdr.index("Kernel Name")]), 0) + 1
    matched = []
    used = {}
    for r in data:
        n = norm(r[hdr.index("Kernel Name")])
        lst = by_name.get(n, [])
        step = max(1, len(lst) // max(1, counts[n]))
        i = used.get(n, 0)
        matched.append(lst[i * step] if i * step < len(lst) else {})
        used[n] = i + 1
    stalls = matched
    with open(out, "w") as f:
        f.write("source: %s (ncu --set full --clock-control none; one block per profiled launch)\n" % rep)
        for k, r in enumerate(data):
            f.write("\n== %s\n" % r[hdr.index("Kernel Name")])
            dur = None
            tot_bytes = 0.0
            for key, label in WANT:
                if key not in hdr:
                    continue
                v, u = r[hdr.index(key)], units[hdr.index(key)]
                f.write("   %-42s %s %s\n" % (label, v, u))
                try:
                    if key.startswith("gpu__time"):
                        dur = float(v) * SCALE.get(u, 1)
                    if key.startswith("dram__bytes"):
                        tot_bytes += float(v) * SCALE.get(u, 1)
                except ValueError:
                    pass
            if dur and tot_bytes:
                f.write("   %-42s %.0f GB/s\n" % ("DRAM read+write rate", tot_bytes / dur / 1e9))
            if k < len(stalls) and stalls[k]:
                tot = float(sum(stalls[k].values()))
                top = sorted(stalls[k].items(), key=lambda x: -x[1])[:6]
                f.write("   warp-stall samples: " + ", ".join("%s %.0f%%" % (a.replace("stall_", ""), 100 * b / tot) for a, b in top) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
