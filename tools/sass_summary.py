#!/usr/bin/env python
"""Per-kernel SASS instruction census of the built library (`cuobjdump -sass`): the mnemonics that prove the TMA bulk-store /
mbarrier / warp-reduction paths are what the compiler emitted, plus the memory instructions, for profiles/*_sass_summary.txt.

    python tools/sass_summary.py [path/to/libsusnet_b200.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# mnemonic prefixes worth counting; B200_PROFILING.md: UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops,
# REDUX = warp reduce, UTMA* = tensor-map TMA (not used here: the outputs are plain row blocks)
GROUPS = ["UBLKCP", "UTMA", "SYNCS", "REDUX", "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "ATOMS", "SHFL", "VOTE", "MATCH",
          "BAR", "MEMBAR", "FENCE", "LDGDEPBAR", "DEPBAR", "POPC", "FLO", "IMAD", "LOP3", "PRMT", "DADD", "DMUL", "DFMA", "FFMA",
          "HMMA", "UTC", "STL", "LDL"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sus_net_b200", "libsusnet_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
        if cur and m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            kernels[cur][op] += 1
    demangle = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: instruction census per kernel (static counts)")
    print("# UBLKCP = cp.async.bulk (TMA bulk copy engine), SYNCS = mbarrier arrive/try_wait, REDUX = redux.sync, STL/LDL = local-memory spills")
    for (name, c), pretty in zip(kernels.items(), demangle):
        short = (pretty.split(">(")[0] + ">") if ">(" in pretty else re.sub(r"\(.*", "", pretty)
        short = short.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        print(f"\n{short}   [{c['_total']} instructions]")
        by_group = collections.OrderedDict()
        for g in GROUPS:
            ops = {op: n for op, n in c.items() if op != "_total" and (op == g or op.startswith(g + ".") or (g in ("UBLKCP", "UTMA", "SYNCS", "UTC") and op.startswith(g)))}
            if ops:
                by_group[g] = ops
        for g, ops in by_group.items():
            detail = ", ".join(f"{op} x{n}" for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
            print(f"  {g:<10s} {sum(ops.values()):5d}   {detail}")


if __name__ == "__main__":
    main()
