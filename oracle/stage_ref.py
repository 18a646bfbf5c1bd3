#!/usr/bin/env python
"""Stage the reference's own Python files of the hot path into oracle/_ref/ (build container only).

The reference is pure Python: there is nothing to compile.  What the GPU box needs in order to time the REAL reference
on its host cores (bench.py `--impl reference`, `cpu_baseline.kind == "reference"`) is the unmodified files
themselves, and /root/reference does not exist there.  This recipe packs them, byte for byte, into ONE archive
(oracle/_ref/reference_py.zip + MANIFEST.json with their sha256) that travels the same way a built .so does:
oracle/_ref/ is git-ignored (no reference source ever enters the history) but not gpurun-ignored.  At run time the
archive is unpacked into a scratch directory outside the repo (oracle/refload.py).

    mapf_primal.py, mapf_gridworld.py
    MARL-curve-main/src/envs/{marl_partial.py, multiagentenv.py}
    MARL-curve-main/src/utils/draw.py            (imported by the two MultiAgentEnv classes)

Called from __graft_entry__.build(); a no-op where /root/reference is absent (the GPU box uses the staged copy).
"""
import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MAPF_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "mapf_primal.py",
    "mapf_gridworld.py",
    "MARL-curve-main/src/envs/marl_partial.py",
    "MARL-curve-main/src/envs/multiagentenv.py",
    "MARL-curve-main/src/utils/draw.py",
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


ARCHIVE = os.path.join(DST, "reference_py.zip")


def stage(verbose=False):
    """Packs the files into oracle/_ref/reference_py.zip (stored, fixed timestamps: the archive is reproducible) and
    writes oracle/_ref/MANIFEST.json {relative path: sha256}.  Returns the manifest, or None when the reference tree
    is not present."""
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return None
    os.makedirs(DST, exist_ok=True)
    manifest = {rel: _sha(os.path.join(SRC, rel)) for rel in FILES}
    mpath = os.path.join(DST, "MANIFEST.json")
    if os.path.exists(ARCHIVE) and os.path.exists(mpath):
        try:
            if json.load(open(mpath)).get("files") == manifest:
                return manifest
        except Exception:
            pass
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_STORED) as z:
        for rel in FILES:
            info = zipfile.ZipInfo(rel, date_time=(1980, 1, 1, 0, 0, 0))
            with open(os.path.join(SRC, rel), "rb") as f:
                z.writestr(info, f.read())
            if verbose:
                print("staged", rel, manifest[rel][:12])
    with open(mpath, "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def unpack(dst_dir):
    """Extracts the staged archive into dst_dir (a scratch directory outside the repo) after checking every file
    against the manifest; returns dst_dir."""
    files = json.load(open(os.path.join(DST, "MANIFEST.json")))["files"]
    with zipfile.ZipFile(ARCHIVE) as z:
        for rel, sha in files.items():
            data = z.read(rel)
            if hashlib.sha256(data).hexdigest() != sha:
                raise RuntimeError("staged reference file %s does not match its manifest hash" % rel)
            out = os.path.join(dst_dir, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            with open(out, "wb") as f:
                f.write(data)
    return dst_dir


if __name__ == "__main__":
    m = stage(verbose=True)
    if m is None:
        print("reference tree not found at", SRC, file=sys.stderr)
        sys.exit(1)
