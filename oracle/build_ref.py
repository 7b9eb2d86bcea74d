"""TEST / BASELINE INFRASTRUCTURE ONLY — snapshots the UNMODIFIED reference's own implementation of the hot path into
oracle/_ref/ so that the STOCK code path can be executed on the GPU box (where /root/reference does not exist):

    python -m oracle.build_ref           (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (no build system, nothing to compile): "building" it = copying, byte for byte, the files the
path consists of — the vendored `CLIP/clip` package (ClipWrapper, ClipGradcam, the explainability ViT), `net.py`,
`unet3d.py`, `arm/optim/lamb.py` — from where they lie under /root/reference.  oracle/_ref/ is git-ignored (no reference
source enters the repository history) but not gpurun-ignored, so it travels to the box like our own built .so.
Consumers: `bench.py --impl reference` / `cpu_baseline` / the `cuda_eager` comparator (through oracle/ref_import.py with
SEMABS_REFERENCE_ROOT pointing here) and nothing else; the product never imports it."""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SEMABS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ["net.py", "unet3d.py", "arm/__init__.py", "arm/optim/__init__.py", "arm/optim/lamb.py", "LICENSE"]
TREES = ["CLIP/clip"]


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "CLIP", "clip", "__init__.py"))


def build(verbose: bool = True) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "CLIP", "clip")):
        return DST if available() else None  # GPU box: only the prebuilt snapshot is used
    os.makedirs(DST, exist_ok=True)
    copied = []
    for t in TREES:
        for d, _, fs in os.walk(os.path.join(SRC, t)):
            for f in fs:
                if f.endswith((".py", ".gz", ".txt")):
                    copied.append(os.path.relpath(os.path.join(d, f), SRC))
    for rel in FILES:
        if os.path.isfile(os.path.join(SRC, rel)):
            copied.append(rel)
    h = hashlib.sha256()
    for rel in sorted(copied):
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        h.update(rel.encode() + open(dst, "rb").read())
    for pkg in ("CLIP", "arm", "arm/optim"):
        init = os.path.join(DST, pkg, "__init__.py")
        if not os.path.exists(init) and os.path.isdir(os.path.join(DST, pkg)):
            src_init = os.path.join(SRC, pkg, "__init__.py")
            if os.path.exists(src_init):
                shutil.copyfile(src_init, init)
    open(os.path.join(DST, "SNAPSHOT.txt"), "w").write(f"byte-for-byte copy of {len(copied)} files from {SRC}\nsha256 {h.hexdigest()}\n" + "\n".join(sorted(copied)) + "\n")
    if verbose:
        print(f"oracle/_ref: {len(copied)} reference files snapshotted (sha256 {h.hexdigest()[:16]})")
    return DST


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
