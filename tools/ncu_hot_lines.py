#!/usr/bin/env python
"""Where a kernel's issue slots go, by CUDA source line: joins the per-instruction counters of an .ncu-rep (captured with
`--set full --import-source on`) with the line table of the SAME build of the library (`-lineinfo`; nvdisasm -g on the cubin
extracted from the .so) and prints (1) the dynamic opcode histogram with the share of warp-stall samples and (2) the source
lines ranked by executed warp instructions.  This is how the round-2 instruction cuts of the Flat step kernel (row builder,
store_state) and of the MLP kernel (weight staging) were found.

    python tools/ncu_hot_lines.py prof.ncu-rep <kernel name substring, mangled> <items, e.g. groups of 32 envs> [--lib path.so] [--top 40]
"""
import argparse
import csv
import io
import os
import re
import subprocess
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(lib, key):
    """[(sass text, (file, line))] of the first function whose mangled name contains `key`, in address order."""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
        for cubin in sorted(os.listdir(d)):
            txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
            lines = txt.split("\n")
            starts = [i for i, l in enumerate(lines) if l.startswith(".text.") and key in l]
            if not starts:
                continue
            seq, cur, i = [], None, starts[0] + 1
            while i < len(lines) and not lines[i].startswith("//-----"):
                m = re.search(r'//## File "([^"]+)", line (\d+)', lines[i])
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", lines[i])
                if m2:
                    seq.append((m2.group(2), cur))
                i += 1
            return seq
    raise SystemExit(f"no function matching {key!r} in {lib}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("items", type=float)
    ap.add_argument("--lib", default=os.path.join(ROOT, "sus_net_b200", "libsusnet_b200.so"))
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    seq = line_table(a.lib, a.kernel)
    if len(seq) != len(data):
        raise SystemExit(f"the report holds {len(data)} instructions, the library's function {len(seq)}: not the same build")

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0

    ops, by_line = defaultdict(lambda: [0.0, 0.0]), defaultdict(lambda: [0.0, 0.0])
    for (txt, loc), r in zip(seq, data):
        p = txt.split()
        op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
        n, s = f(r, "Instructions Executed"), f(r, "# Samples")
        ops[op][0] += n; ops[op][1] += s
        by_line[loc or ("?", 0)][0] += n; by_line[loc or ("?", 0)][1] += s
    tot = sum(v[0] for v in ops.values()); tots = sum(v[1] for v in ops.values()) or 1.0
    print(f"{tot / 1e6:.2f} M warp instructions executed, {tot / a.items:.1f} per item, {int(tots)} warp-stall samples")
    print("opcode        per item   share   samples")
    for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:24]:
        print(f"{op:12s} {n / a.items:9.1f}  {100 * n / tot:5.1f} %  {100 * s / tots:5.1f} %")
    print("source line                    per item   share   samples")
    for (fn, ln), (n, s) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{fn + ':' + str(ln):28s} {n / a.items:9.1f}  {100 * n / tot:5.1f} %  {100 * s / tots:5.1f} %")


if __name__ == "__main__":
    main()
