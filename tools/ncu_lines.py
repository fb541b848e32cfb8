#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source=cuda,sass` dump by source line.

    ncu -i prof.ncu-rep --page source --csv --print-source=cuda,sass > cs.csv
    python tools/ncu_lines.py cs.csv <warp-iterations per launch> [top]
"""
import csv
import sys


def main():
    path, per = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
    rows = list(csv.reader(open(path)))
    agg, tot, tot_s, seen, cur = [], 0, 0, set(), None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            if cur in seen:  # second captured launch: stop
                break
            seen.add(cur)
            continue
        if r[0] == "Line No":
            hdr = r
            ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= ci or not r[0].isdigit():
            continue
        try:
            n, s = int(r[ci]), int(r[cs])
        except ValueError:
            continue
        if n > 0 or s > 0:
            agg.append((n, s, cur, int(r[0]), r[1].strip()[:100]))
        tot += n
        tot_s += s
    print(f"total warp-instructions {tot}  per warp-iteration {tot / per:.0f}  samples {tot_s}")
    print("by instructions:")
    for n, s, f, l, src in sorted(agg, reverse=True)[:top]:
        print(f"{n / per:8.1f} {100.0 * s / max(tot_s, 1):5.1f}% {f}:{l}  {src}")
    print("by stall samples:")
    for n, s, f, l, src in sorted(agg, key=lambda x: -x[1])[:25]:
        print(f"{n / per:8.1f} {100.0 * s / max(tot_s, 1):5.1f}% {f}:{l}  {src}")


if __name__ == "__main__":
    main()
