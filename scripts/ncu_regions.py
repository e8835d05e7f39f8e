"""Per-region instruction counts of one kernel from an ncu capture taken with --import-source on:
the SASS of the kernel is cut at its barriers (BAR.SYNC, mbarrier try_wait) and each piece reports its share of
the executed warp instructions and of the stall samples, plus the opcode mix.

  python scripts/ncu_regions.py gpurun_out/r02_prof_resize.ncu-rep [units]      # units: divide counts (e.g. tiles)
"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print("# " + rows[0][1][:110])
    hdr, data = rows[1], rows[2:]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][ia], 16)
    tot = sum(int(r[iex]) for r in data)
    tots = sum(int(r[ismp]) for r in data) or 1
    print("warp instructions executed: %d (%.0f per unit), stall samples: %d" % (tot, tot / units, tots))
    marks = [0] + [int(r[ia], 16) - base for r in data if "BAR.SYNC" in r[isrc] or "SYNCS.PHASECHK" in r[isrc]] + [1 << 30]
    for lo, hi in zip(marks[:-1], marks[1:]):
        sel = [r for r in data if lo <= int(r[ia], 16) - base < hi]
        e, sm = sum(int(r[iex]) for r in sel), sum(int(r[ismp]) for r in sel)
        if e:
            print("  SASS +%05x..+%05x  instructions %5.1f %% (%7.0f per unit)  samples %5.1f %%" % (lo, min(hi, 0xFFFFF), 100.0 * e / tot, e / units, 100.0 * sm / tots))
    mix = collections.Counter()
    for r in data:
        t = r[isrc].split()
        op = t[0] if not t[0].startswith("@") else t[1]
        mix[op.split(".")[0]] += int(r[iex])
    print("opcode mix: " + ", ".join("%s %.1f %%" % (k, 100.0 * v / tot) for k, v in mix.most_common(12)))


if __name__ == "__main__":
    main()
