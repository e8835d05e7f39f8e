"""Turn the ncu captures in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py r01
writes profiles/<tag>_launches.csv (kernel, count, avg us, share of the step),
profiles/<tag>_<kernel>_full.txt (key metrics of the --set full capture) and copies the
bench JSON lines.  Needs the ncu CLI (reads .ncu-rep files, no GPU).
"""

import collections
import csv
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_static", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def launches(tag):
    path = os.path.join(OUT, tag + "_launches.csv")
    if not os.path.exists(path):
        path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) > mi:
            try:
                agg.setdefault(r[ki].split("(")[0], []).append(float(r[mi].replace(",", "")))
            except ValueError:
                pass
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(DST, tag + "_launches.csv"), "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 20 --warmup 5 (device sections)\n")
        fh.write("# per-launch times under ncu are cold-cache and serialised: compare the SHARES\n")
        fh.write("kernel,launches,avg_us,min_us,max_us,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write("%s,%d,%.2f,%.2f,%.2f,%.1f\n" % (k, len(v), sum(v) / len(v) / 1e3, min(v) / 1e3, max(v) / 1e3, 100 * sum(v) / total))


def full(tag, rep, label):
    path = os.path.join(OUT, rep)
    if not os.path.exists(path):
        return
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    with open(os.path.join(DST, "%s_%s_full.txt" % (tag, label)), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on (from gpurun_out/%s)\n" % rep)
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            fh.write("\n== %s\n" % name)
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    fh.write("%-85s %s %s\n" % (k, r[i], units[i]))
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")) if "dram__bytes_read.sum" in hdr else 0
            fh.write("# traffic = dram read + write per launch, unit as above\n")


def steady_dram(tag):
    """profiles/<tag>_fill_steady_dram.txt from the single-purpose DRAM captures of scripts/ncu_round2.sh
    (application replay, no cache control, launches 240.. of a rotating chain)."""
    lines = []
    for dt in ("FP32", "FP16"):
        path = os.path.join(OUT, "%s_fill_steady_dram_%s.csv" % (tag, dt))
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        hdr = rows[start]
        ki, mn, mv, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
        per = collections.OrderedDict()
        for r in rows[start + 1:]:
            if len(r) > mv:
                per.setdefault(r[idi], {"kernel": r[ki].split("(")[0]})[r[mn]] = r[mv].replace(",", "")
        lines.append("== %s: launches 240..247 of 256 back-to-back %s fills of 38,535,168 B rotating over 4 region sets (154 MB)" % (
            next(iter(per.values()))["kernel"], "64 x FP32[3,224,224]" if dt == "FP32" else "1 x FP16[128,3,224,224]"))
        lines.append("%-8s %16s %16s %14s %14s" % ("launch", "dram_write_B", "dram_read_B", "duration_ns", "warps_active_%"))
        tw = 0.0
        for k, v in per.items():
            w = float(v.get("dram__bytes_write.sum", 0))
            tw += w
            lines.append("%-8s %16d %16d %14s %14s" % (k, w, float(v.get("dram__bytes_read.sum", 0)), v.get("gpu__time_duration.sum", "?"),
                                                    v.get("sm__warps_active.avg.pct_of_peak_sustained_active", "?")))
        lines.append("mean dram write per launch: %.0f B = %.3f of the 38,535,168 algorithmic bytes (the rest is overwritten in L2 before eviction)" % (tw / len(per), tw / len(per) / 38535168))
        lines.append("")
    if lines:
        with open(os.path.join(DST, tag + "_fill_steady_dram.txt"), "w") as fh:
            fh.write("# ncu --replay-mode application --cache-control none --clock-control none -k regex:fill_uniform_kernel --launch-skip 240 --launch-count 8\n")
            fh.write("# (scripts/ncu_round2.sh, scripts/fill_steady.py): steady-state HBM traffic of the fill launch; durations under ncu are serialised launches\n\n")
            fh.write("\n".join(lines))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(DST, exist_ok=True)
    launches(tag)
    steady_dram(tag)
    full(tag, tag + "_prof_fill_uniform.ncu-rep", "fill_uniform_kernel")
    if tag != "r01":  # the captures below are round 1's; later rounds name their files <tag>_*
        full(tag, tag + "_prof_resize.ncu-rep", "resize_pack_kernel")
        rep = os.path.join(OUT, tag + "_prof_resize.ncu-rep")
        dst = os.path.join(DST, tag + "_resize_pack_kernel_full.txt")
        if os.path.exists(rep) and os.path.exists(dst):  # where the instructions go: the SASS cut at its barriers
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), rep, "3136"], capture_output=True, text=True)
            with open(dst, "a") as fh:
                fh.write("\n# scripts/ncu_regions.py (unit = one 32x32 output tile, 3136 per launch): setup | row copies + tables | wait | horizontal pass | vertical pass\n")
                fh.write(r.stdout)
        print(sorted(os.listdir(DST)))
        return
    full(tag, "prof_fill.ncu-rep", "fill_kernel")
    full(tag, "prof_pack.ncu-rep", "pack_image_kernel")
    full(tag, "prof_check.ncu-rep", "check_kernel")
    full(tag, "prof_resize.ncu-rep", "resize_pack_kernel")
    full(tag, "prof_deflate.ncu-rep", "deflate_chunk_kernel")
    full(tag, "prof_fill_unaligned.ncu-rep", "fill_unaligned_kernel")
    for src, dst in (("bench.json", "bench_b200.json"), ("bench_ref.json", "bench_reference.json"),
                     ("fill_sweep.txt", "fill_sweep.txt"), ("store_bench.txt", "store_bench.txt"),
                     ("nvidia_smi.txt", "nvidia_smi.txt")):
        p = os.path.join(OUT, src)
        if os.path.exists(p) and os.path.getsize(p):
            shutil.copy(p, os.path.join(DST, tag + "_" + dst))
    print(sorted(os.listdir(DST)))


if __name__ == "__main__":
    main()
