#!/usr/bin/env python3
"""Turn ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_rN.csv  profiles/rN_launches.md
    python tools/ncu_summary.py kernels  gpurun_out/prof_rN.ncu-rep  profiles/rN_kernels.md [profiles/dominant_kernel_traffic.json [preset]]
"""
import collections
import csv
import json
import re
import subprocess
import sys

STAGE_OF = {"k_scan_emit": "scan_emit", "k_os_pass": "radix_sort_one_pass", "k_os_hist": "radix_sort_hist", "k_scan_reads": "scan_reads", "k_emit_pairs": "emit_pairs", "k_rs_hist": "radix_sort", "k_rs_scatter": "radix_sort",
            "k_reduce1": "reduce1", "k_match": "match", "k_reduce2": "reduce2", "k_finalize": "finalize", "k_entropy_sum": "entropy",
            "k_entropy_compact": "entropy", "k_seg_heads": "segments", "k_seg_ids": "segments", "k_junc_init": "reduce1", "k_flag_scan": "flag_scans"}


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("ns", "nsecond"):
            v *= 1e-3
        agg.setdefault(short(row["Kernel Name"]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("Source: `%s`. Per-launch times are cold-cache and serialised by the profiler: compare SHARES, not absolutes.\n\n" % src)
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
    print("wrote", dst)


WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes/inst"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("inst_executed", "warp inst")]


PEAK_GBS = 6540          # MEASURED_PEAKS.json hbm_gbs on this pool's B200s


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def kernels(src, dst, traffic_json=None, preset=None):
    if src.endswith(".csv"):                       # already exported on the GPU box (`ncu -i rep --page raw --csv`): the .ncu-rep files are too big to bring back
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    seen = collections.OrderedDict()
    for r in rows[2:]:
        seen.setdefault(short(r[col["Kernel Name"]]), []).append(r)
    traffic = {}
    with open(dst, "w") as f:
        f.write("# ncu `--set full --clock-control none` summary\n\nSource: `%s` (kept out of git; regenerate with the command in DESIGN.md).\n\n" % src)
        f.write("Last column: DRAM bytes moved / kernel time, and its share of the measured HBM peak (MEASURED_PEAKS.json hbm_gbs).\n\n")
        f.write("| kernel | " + " | ".join(n for _, n in WANT) + " | achieved DRAM |\n|---|" + "---:|" * (len(WANT) + 1) + "\n")
        for k, rs in seen.items():
            r = rs[-1]
            cells = []
            for m, _ in WANT:
                cells.append(("%s %s" % (r[col[m]], units[col[m]])).strip() if m in col else "-")
            gbs = ""
            if "dram__bytes_read.sum" in col and "gpu__time_duration.sum" in col:
                tot = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
                tu = units[col["gpu__time_duration.sum"]]
                secs = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3}.get(tu, 1.0)
                if secs > 0:
                    gbs = "%.0f GB/s (%.0f%% of %d)" % (tot / secs / 1e9, 100 * tot / secs / 1e9 / PEAK_GBS, PEAK_GBS)
            f.write("| `%s` | " % k + " | ".join(cells) + " | " + gbs + " |\n")
            if "dram__bytes_read.sum" in col:
                t = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
                    to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
                st = STAGE_OF.get(k.split("<")[0])
                if st:
                    traffic[st] = traffic.get(st, 0) + t
    if traffic_json:
        # {preset: {stage: bytes per launch}}: bench.py copies the dominant stage's value of its preset into roofline.traffic
        allp = {}
        try:
            with open(traffic_json) as f:
                allp = json.load(f)
            if allp and not all(isinstance(v, dict) for v in allp.values()):
                allp = {}
        except Exception:
            pass
        allp[preset or "c2"] = {k: int(v) for k, v in traffic.items()}
        with open(traffic_json, "w") as f:
            json.dump(allp, f, indent=1, sort_keys=True)
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernels(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None, sys.argv[5] if len(sys.argv) > 5 else None)
