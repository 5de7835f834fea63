"""Per-source-line summary of an ncu report's source page: python tools/ncu_lines.py REPORT.ncu-rep [TOP]
(needs `ncu` on PATH; the kernel must have been compiled with -lineinfo and captured with --import-source on)."""
import collections
import csv
import subprocess
import sys


def f(x):
    try:
        return float(x)
    except (TypeError, ValueError):
        return 0.0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout.splitlines()
    lines = collections.OrderedDict()
    cur, hdr = None, None
    for r in csv.reader(txt):
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "":
            continue
        d = dict(zip(hdr[4:], r[4:]))
        L = lines.setdefault((cur, r[0]), [r[1], 0, 0, 0])
        L[1] += f(d.get("# Samples")); L[2] += f(d.get("Instructions Executed")); L[3] += f(d.get("Thread Instructions Executed"))
    ts = sum(v[1] for v in lines.values()) or 1
    ti = sum(v[2] for v in lines.values()) or 1
    print(f"samples {ts:.0f}  warp instructions {ti:.0f}")
    pf, pfi = collections.Counter(), collections.Counter()
    for (c, _), v in lines.items():
        pf[c] += v[1]; pfi[c] += v[2]
    for k in pf:
        print(f"  {k:28s} samples {pf[k] / ts * 100:5.1f}%  instructions {pfi[k] / ti * 100:5.1f}%")
    for (c, l), v in sorted(lines.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{c:24s} {l:>5s} s={v[1] / ts * 100:5.2f}% i={v[2] / ti * 100:5.2f}% thr={v[3] / max(v[2], 1):5.1f} {v[0].strip()[:110]}")


if __name__ == "__main__":
    main()
