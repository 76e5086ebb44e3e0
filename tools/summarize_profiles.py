"""Turns ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r1.csv profiles/r1_launches_bench.md
    python tools/summarize_profiles.py full gpurun_out/prof_rec.ncu-rep profiles/r1_rec_kernels_full.md
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        v = v / 1e3 if row[ui] == "ns" else v * 1e3 if row[ui] == "ms" else v
        agg[row[ki]][0] += 1
        agg[row[ki]][1] += v
        tot += v
    with open(dst, "w") as out:
        out.write("# ncu launch list (gpu__time_duration.sum, --clock-control none): cold-cache, serialised —\n"
                  "# compare SHARES, not absolutes.  Source: `%s`\n\n" % src)
        out.write(f"total {tot/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n\n")
        out.write("| share | total us | launches | avg us | kernel |\n|---:|---:|---:|---:|---|\n")
        mine = 0.0
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            if k.startswith("mrg::") or "mrg::" in k.split("(")[0]:
                mine += t
            out.write(f"| {100*t/tot:.1f}% | {t:.1f} | {n} | {t/n:.1f} | `{k[:110]}` |\n")
        out.write(f"\nkernels of this library (`mrg::*`): {100*mine/tot:.1f}% of the listed GPU time\n")
    print("wrote", dst)


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write("# ncu --set full (--clock-control none), selected raw metrics per captured launch.  Source: `%s`\n\n" % src)
        for r in rows[2:]:
            out.write("## %s\n\n| metric | value | unit |\n|---|---:|---|\n" % r[hdr.index("Kernel Name")])
            for k in KEYS:
                if k in hdr:
                    out.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            stalls = [(h, float(r[i])) for i, h in enumerate(hdr)
                      if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and r[i]]
            out.write("\nwarp stall reasons (warps per issue-active cycle):\n\n")
            for h, v in sorted(stalls, key=lambda x: -x[1])[:8]:
                name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
                out.write(f"* {name}: {v:.3f}\n")
            out.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
