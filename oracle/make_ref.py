"""TEST / BASELINE INFRASTRUCTURE -- recipe that ships the reference's own hot-path Python files to the GPU box.

The reference is pure Python (no setup.py / pyproject at its root, so `pip install /root/reference` cannot work):
"installing" it is a directory copy.  This script copies the two package trees the hot paths import --
`flowdiffusion/` and `diffuser/`, *.py only, ~1 MB -- from /root/reference into `baseline/_ref/`, which is
git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so it travels with the snapshot.
`oracle/ref_import.py` then imports the UNMODIFIED modules from there behind its inert third-party stubs, and

  * `bench.py --impl reference`            times the reference's own CPU path     (cpu_baseline.kind = "reference"),
  * the `gpu_eager` block of the bench line times the same modules on cuda:0     (E32 / E16),

instead of the oracle port.  Run by `__graft_entry__.build()` whenever /root/reference is mounted:
    python oracle/make_ref.py
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = ("flowdiffusion", "diffuser")


def make(src: str = SRC, dst: str = DST) -> int:
    if not os.path.isdir(os.path.join(src, "flowdiffusion", "flowdiffusion")):
        return 0
    n = 0
    for tree in TREES:
        for d, _, files in os.walk(os.path.join(src, tree)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                rel = os.path.relpath(os.path.join(d, f), src)
                out = os.path.join(dst, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(d, f), out)
                n += 1
    with open(os.path.join(dst, "PROVENANCE.txt"), "w") as fh:
        fh.write(f"{n} unmodified *.py files copied from {src} ({', '.join(TREES)}) by oracle/make_ref.py\n")
    return n


if __name__ == "__main__":
    k = make()
    print(f"baseline/_ref: {k} files" if k else f"{SRC} not mounted: nothing copied", file=sys.stderr)
