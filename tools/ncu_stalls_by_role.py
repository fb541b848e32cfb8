#!/usr/bin/env python
"""Split the warp-stall samples of the warp-specialised fused kernel (k_step_ws) by WARP ROLE, from the per-instruction source
page of an `ncu --set full --import-source on` capture: an instruction belongs to the emitter warp when its source line lies
inside emit_sub / emitter_loop / emitter_dispatch (susnet_api.cu) -- the emitter executes nothing else after the prologue --
and to the compute warps otherwise.  (`ncu` has no per-warp-role breakdown; source lines are the role's signature.)

    python tools/ncu_stalls_by_role.py prof.ncu-rep > profiles/r02_stalls_by_role.txt
"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def emitter_line_ranges():
    """Line ranges of the emitter-only functions in the current susnet_api.cu."""
    src = open(os.path.join(ROOT, "sus_net_b200", "csrc", "susnet_api.cu")).read().splitlines()
    ranges, start, depth, seen_brace = [], None, 0, False
    for i, line in enumerate(src, 1):
        if start is None and re.search(r"__device__ __forceinline__ void (emit_sub|emitter_loop|emitter_dispatch)\(", line):
            start, depth, seen_brace = i, 0, False
        if start is not None:
            depth += line.count("{") - line.count("}")
            seen_brace = seen_brace or "{" in line
            if seen_brace and depth == 0:
                ranges.append((start, i))
                start = None
    return ranges


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    ranges = emitter_line_ranges()
    cur_file, hdr, cur_line = None, None, -1
    sass = {}  # address -> (file, line, stalls dict, instructions executed); one entry per SASS instruction
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file, hdr = row[1], None
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            hdr = row
            continue
        if hdr is None or cur_file is None:
            continue
        i_addr = hdr.index("Address")
        addr = row[i_addr]
        if not re.fullmatch(r"0x[0-9a-fA-F]+", addr or ""):
            try:
                cur_line = int(row[0])  # a source-line summary row: the SASS rows that follow belong to it
            except ValueError:
                pass
            continue
        ln = cur_line
        stalls = {}
        for k, v in zip(hdr, row):
            if k.startswith("stall_") and "(Not Issued)" not in k:
                try:
                    stalls[k] = int(float(v or 0))
                except ValueError:
                    pass
        try:
            n_exec = int(float(row[hdr.index("Instructions Executed")] or 0))
        except ValueError:
            n_exec = 0
        key = int(addr, 16)
        # an instruction inlined from a header is listed under the header AND under its call site: keep the call-site row of
        # susnet_api.cu for the role decision, the counters are identical
        if key not in sass or cur_file.endswith("susnet_api.cu"):
            sass[key] = (cur_file, ln, stalls, n_exec)
    # the emitter's SASS region: from the first to the last instruction whose susnet_api.cu line is inside an emitter function
    em_addrs = [a for a, (f, ln, _s, _n) in sass.items() if f.endswith("susnet_api.cu") and any(lo <= ln <= hi for lo, hi in ranges)]
    lo, hi = min(em_addrs), max(em_addrs)
    roles = {"emitter": {}, "compute": {}}
    insts = {"emitter": 0, "compute": 0}
    for a, (f, ln, stalls, n_exec) in sass.items():
        role = "emitter" if lo <= a <= hi else "compute"
        insts[role] += n_exec
        for k, v in stalls.items():
            roles[role][k] = roles[role].get(k, 0) + v
    print(f"# warp-stall samples of k_step_ws by warp role ({os.path.basename(rep)}): {len(sass)} SASS instructions; the emitter warp's code is")
    print(f"# the address range [{lo:#x}, {hi:#x}] spanned by the instructions of emit_sub / emitter_loop (susnet_api.cu lines {ranges}),")
    print("# helpers inlined into that range (bulk_wait, mbar_wait, drain) included; everything else is executed by the compute warps")
    for role in ("compute", "emitter"):
        tot = sum(roles[role].values())
        print(f"\n{role} warps: {tot} stall samples, {insts[role]} warp-instructions executed")
        for k, v in sorted(roles[role].items(), key=lambda kv: -kv[1])[:8]:
            if v:
                print(f"  {k:<26s} {v:8d}  {100.0 * v / max(tot, 1):5.1f} %")


if __name__ == "__main__":
    main()
