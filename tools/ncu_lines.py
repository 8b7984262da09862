"""Aggregate an .ncu-rep's source page (needs -lineinfo and --import-source on) into executed
warp instructions and stall samples per source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top=50] [index of the launch in the report]
"""
import csv
import subprocess
import sys


def main(rep, top=50, skip=None):
    sel = [] if skip is None else ["--launch-skip", str(skip), "--launch-count", "1"]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"] + sel,
                         capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, []
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur = r[1]
            continue
        if r[0] == "Line No":
            hdr = {}
            for i, h in enumerate(r):
                hdr.setdefault(h, i)
            continue
        if hdr and r[0].isdigit():
            try:
                agg.append((int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]]), cur.split("/")[-1],
                            int(r[0]), r[1].strip()[:100]))
            except (ValueError, KeyError, IndexError):
                continue
    tot = sum(a[0] for a in agg) or 1
    ts = sum(a[1] for a in agg) or 1
    print("total warp instructions", tot, "stall samples", ts)
    byfile = {}
    for a in agg:
        byfile[a[2]] = byfile.get(a[2], 0) + a[0]
    print({k: round(v / tot, 3) for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])})
    for a in sorted(agg, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / ts, a[2], a[3], a[4]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50, int(sys.argv[3]) if len(sys.argv) > 3 else None)
