#!/usr/bin/env python
"""Write profiles/traffic.json from the ncu passes of tools/ncu_limiter.sh: DRAM bytes per launch of the headline kernel with
the feature tensors in L2-compressible / ordinary memory, the best SM->L2 store stream of tools/micro/store_ceiling_bench.cu per
memory kind, and the content hash of the library build the capture was made with (bench.py reports `roofline.traffic` only
while the built library has that hash).

    python tools/update_traffic.py gpurun_out/r02_limiter_metrics.csv gpurun_out/r02_limiter_metrics_plainmem.csv \
        gpurun_out/r02_store_ceiling.json gpurun_out/r02_lib_hash.txt
"""
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith('"')]
    per_launch = {}
    for r in csv.DictReader(io.StringIO("\n".join(lines))):
        if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            per_launch.setdefault(r["ID"], {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    tot = [v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in per_launch.values()]
    rd = [v["dram__bytes_read.sum"] for v in per_launch.values()]
    return sum(tot) / len(tot), sum(rd) / len(rd)


def main():
    comp_csv, plain_csv, ceiling_json, hash_txt = sys.argv[1:5]
    c_tot, c_rd = dram_bytes(comp_csv)
    p_tot, p_rd = dram_bytes(plain_csv)
    best = {}
    for r in json.load(open(ceiling_json))["results"]:
        if r["variant"] != "memset" and r["gbs"] > best.get(r["memory"], (0, ""))[0]:
            best[r["memory"]] = (r["gbs"], f'{r["variant"]}: {r["config"]}')
    out = {
        "kernel": "k_step_ws<BASE> (fused step + Global encode)", "envs_per_launch": 1048576,
        "lib_hash": open(hash_txt).read().strip(),
        "dram_bytes_per_launch": c_tot,
        "source": f"{os.path.basename(comp_csv)} (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum; feature tensors in "
                  f"L2-compressible memory: {c_rd / 1e6:.1f} MB read + {(c_tot - c_rd) / 1e6:.1f} MB written)",
        "uncompressed": {"dram_bytes_per_launch": p_tot,
                         "source": f"{os.path.basename(plain_csv)} (SUSNET_COMPRESSIBLE=0: feature tensors in cudaMalloc memory)"},
        "store_ceiling_gbs": {k: v[0] for k, v in best.items()},
        "store_ceiling_patterns": {k: v[1] for k, v in best.items()},
        "store_ceiling_source": f"{os.path.basename(ceiling_json)}: best of every pattern of tools/micro/store_ceiling_bench.cu",
    }
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
