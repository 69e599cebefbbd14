#!/usr/bin/env python3
"""profiles/rN_stage_table.md from bench lines: per preset, every stage with its time, SURVEY 8(d) bytes, GB/s and fraction of the HBM peak.

    python tools/stage_table.py profiles/r2_stage_table.md c2=profiles/r2_bench_c2.json c3=profiles/r2_bench_c3.json ...
"""
import json
import sys


def main():
    dst = sys.argv[1]
    with open(dst, "w") as f:
        f.write("# Per-stage roofline table (bench.py `roofline.stages`; CUDA events on the library's compute stream)\n\n")
        f.write("Bytes = SURVEY 8(d) taken literally (see DESIGN.md §3); peak = MEASURED_PEAKS.json hbm_gbs.  The sort and the segmentation are overhead by 8(d):\n"
                "their fraction is against one read + one write of what they permute / label, and they are not part of the pipeline numerator.\n\n")
        for arg in sys.argv[2:]:
            name, path = arg.split("=", 1)
            d = json.loads(open(path).read().strip().split("\n")[-1])
            r = d["roofline"]
            ws = d.get("workload_stats", {})
            f.write("## %s — %s\n\n" % (name, d["config"]["workload"]))
            f.write("%d GPU(s); %s records, %s read-junction pairs, %s junctions; device pipeline %.3f ms per step; value %.3g spliced alignments/s; "
                    "pipeline %.0f GB/s = **%.3f** of %.0f GB/s; dominant stage `%s` at **%.3f**.\n\n"
                    % (d["n_gpus"], ws.get("records"), ws.get("read_junction_pairs"), ws.get("junctions"), d["device_ms_per_step"], d["value"],
                       r["pipeline_achieved_gbs"], r["pipeline_frac"], r["peak"], r["kernel"], r["frac"]))
            f.write("| stage | ms | share | 8(d) bytes | GB/s | fraction of peak |\n|---|---:|---:|---:|---:|---:|\n")
            tot = sum(v["ms"] for v in r["stages"].values())
            for k, v in r["stages"].items():
                f.write("| %s | %.3f | %.0f %% | %.3g | %.0f | %.3f |\n" % (k, v["ms"], 100 * v["ms"] / tot, v["alg_bytes"], v["gbs"], v["frac"]))
            f.write("\n")
            if d.get("e2e"):
                f.write("e2e (C ABI, pinned host columns): %.3g spliced alignments/s, %.2f ms per step, %.0f MB H2D + %.0f MB D2H per step.\n\n"
                        % (d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"] / 1e6, d["e2e"]["d2h_bytes_per_step"] / 1e6))
            if d.get("e2e_bam"):
                b = d["e2e_bam"]
                f.write("e2e_bam (BAM file -> output files): %.3g spliced alignments/s, %.2f s; breakdown (rank 0) %s; parity %s.\n\n"
                        % (b["value"], b["seconds"], json.dumps(b.get("breakdown_s_rank0")), json.dumps(b.get("parity", {}).get("equals_reference_md5"))))
            if d.get("cpu_baseline"):
                f.write("cpu_baseline: %s\n\n" % json.dumps(d["cpu_baseline"]))
    print("wrote", dst)


if __name__ == "__main__":
    main()
