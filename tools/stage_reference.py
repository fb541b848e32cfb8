#!/usr/bin/env python
"""Stage the UNMODIFIED reference for the GPU box: copy `/root/reference/src` (Python sources only) into the git-ignored
`baseline/_ref/src`.  `/root/reference` does not exist on the GPU box, but the working-directory snapshot `gpurun` ships
does include `baseline/_ref/`, so there the reference's own CPU implementation can be timed beside the GPU path
(`bench.py` `cpu_baseline.kind == "reference"`, `bench.py --impl reference`) and the drop-in tests can run the reference's
own `ReplayBuffer.populate` / `train()` against the CUDA env (`tests/test_gpu_reference_dropin.py`).

Nothing under `baseline/_ref/` is ever committed (`.gitignore`) and nothing in `sus_net_b200/` imports it.  The reference
needs `gymnasium` (not installed, no network); it is imported through the stand-in in `tests/_shims` (oracle/ref_harness.py).

    python tools/stage_reference.py [--source /root/reference] [--check]
"""
import argparse
import hashlib
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")


def file_hashes(root):
    out = {}
    for d, _dirs, files in os.walk(os.path.join(root, "src")):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                out[os.path.relpath(p, root)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    return out


def stage(source="/root/reference", quiet=False):
    """Copy the sources; returns the manifest, or None if `source` is absent (GPU box: the staged copy is used as is)."""
    if not os.path.isdir(os.path.join(source, "src", "environment")):
        return None
    want = file_hashes(source)
    manifest_path = os.path.join(DEST, "STAGED.json")
    if os.path.exists(manifest_path):
        try:
            if json.load(open(manifest_path)).get("files") == want and file_hashes(DEST) == want:
                return json.load(open(manifest_path))
        except Exception:  # noqa: BLE001
            pass
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    for rel in want:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(source, rel), dst)
    manifest = {"source": source, "staged_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "files": want,
                "note": "unmodified copies of the reference's Python sources; test / baseline infrastructure only"}
    json.dump(manifest, open(manifest_path, "w"), indent=1)
    if not quiet:
        print(f"staged {len(want)} files from {source} into {DEST}")
    return manifest


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default="/root/reference")
    ap.add_argument("--check", action="store_true", help="verify the staged copy against its manifest (and the source if present)")
    a = ap.parse_args()
    if a.check:
        m = json.load(open(os.path.join(DEST, "STAGED.json")))
        assert file_hashes(DEST) == m["files"], "staged files differ from the manifest"
        if os.path.isdir(a.source):
            assert file_hashes(a.source) == m["files"], "staged files differ from the source"
        print(f"ok: {len(m['files'])} staged files are unmodified copies")
        return
    if stage(a.source) is None:
        sys.exit(f"{a.source} not found")


if __name__ == "__main__":
    main()
