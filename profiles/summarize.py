"""Turn raw ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv profiles/r01_launches_<tag>.txt
    python profiles/summarize.py full gpurun_out/prof_gemm.ncu-rep profiles/r01_ncu_<tag>.txt
    python profiles/summarize.py traffic gpurun_out/prof_gemm.ncu-rep profiles/r01_traffic.json
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        key = (r[ki][:90], r[gi])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s ; %d launches, %.1f us total\n" % (src, sum(v[0] for v in agg.values()), tot))
        for (k, g), v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-92s grid %-14s n=%4d total %9.1f us share %5.1f%% avg %8.2f us\n" % (k, g, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; source: %s\n" % src)
        for r in rows[2:]:
            f.write("\n%s  grid %s block %s\n" % (r[idx["Kernel Name"]][:100], r[idx["Grid Size"]], r[idx["Block Size"]]))
            for m in METRICS:
                if m in idx:
                    f.write("  %-78s %s %s\n" % (m, r[idx[m]], units[idx[m]]))


def traffic(src, dst):
    """Per-kernel DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) -> JSON read by bench.py.
    ncu flushes the caches before every replay, so these are COLD-cache figures (upper bounds for the in-pipeline run)."""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        b = sum(float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += b
    res = {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0]} for k, v in agg.items()}
    gem = [v for k, v in agg.items() if "gemm" in k]
    res["_gemm_class_avg_dram_bytes_per_launch"] = sum(v[1] for v in gem) / max(1, sum(v[0] for v in gem))
    res["_source"] = "%s (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum; cold cache)" % src
    json.dump(res, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
