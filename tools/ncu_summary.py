"""Summaries of ncu output for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py launches <launch list csv> <out.txt> "<command>" "<workload>"
    python tools/ncu_summary.py kernel <file.ncu-rep> <out.json> "<command>" "<workload>"
"""
import collections, csv, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic_bytes", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "second": 1.0,
         "Ghz": 1e9, "Mhz": 1e6}


def launches(path, out, command, workload):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for x in csv.DictReader(lines):
        if x.get("Metric Name") == "gpu__time_duration.sum":
            v = float(x["Metric Value"].replace(",", "")) * SCALE.get(x["Metric Unit"], 1.0) * 1e6
            rows.append((x["Kernel Name"].split("(")[0], x["Grid Size"], x["Block Size"], v))
    tot = sum(r[3] for r in rows)
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r[0], [0, 0.0])
        a[0] += 1
        a[1] += r[3]
    with open(out, "w") as f:
        f.write(command + "\n" + workload + "\n")
        f.write("Per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes.  %d launches, %.1f us in total.\n\n" % (len(rows), tot))
        f.write("%-60s %7s %12s %9s %7s\n" % ("kernel", "count", "total_us", "mean_us", "share"))
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %7d %12.1f %9.2f %7.3f\n" % (k[:60], c, t, t / c, t / tot))
        # the last decode step: from the last decode_mega launch to the end
        idx = [i for i, r in enumerate(rows) if "decode_mega" in r[0]]
        if idx:
            step = rows[idx[-1]:]
            st = sum(r[3] for r in step)
            f.write("\none decode step = the last %d launches (one CUDA-graph replay): %.1f us serialised\n" % (len(step), st))
            for r in step:
                f.write("  %-56s grid %-14s block %-14s %10.2f us  share %.3f\n" % (r[0][:56], r[1], r[2], r[3], r[3] / st))
    print(open(out).read())


def kernel(path, out, command, workload):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, units = r[0], r[1]
    res = []
    for row in r[2:]:
        d = {"kernel": row[h.index("Kernel Name")].split("(")[0], "grid": row[h.index("Grid Size")]}
        for k in KEEP:
            if k in h:
                i = h.index(k)
                try:
                    d[k] = float(row[i].replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    pass
        res.append(d)
    j = {"command": command, "workload": workload, "launches": res}
    with open(out, "w") as f:
        json.dump(j, f, indent=1)
    print(json.dumps(j, indent=1))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](*sys.argv[2:6])
