#!/usr/bin/env python
"""Recipe: copy the UNMODIFIED reference sources the hot path needs into git-ignored baseline/_ref/.

    python baseline/make_ref.py            (build container only: needs /root/reference)

baseline/_ref/ is listed in .gitignore (never committed) but not in .gpurunignore, so it travels to the GPU
box with the snapshot, where /root/reference does not exist.  Nothing is edited: files are byte-for-byte
copies (checked below); `baseline/ref_harness.py` imports them through an `__init__`-bypass shim because the
package `__init__`s pull tensordict / torchrl, which are not in the image.
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SD_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = ["world_model", "utils", "ablations", "configs"]
KEEP = (".py", ".yaml")


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} not present: baseline/_ref left as it is")
        return 0
    n = 0
    for tree in TREES:
        for dirpath, _, files in os.walk(os.path.join(SRC, tree)):
            for f in files:
                if not f.endswith(KEEP):
                    continue
                s = os.path.join(dirpath, f)
                d = os.path.join(DST, os.path.relpath(s, SRC))
                os.makedirs(os.path.dirname(d), exist_ok=True)
                if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                    shutil.copyfile(s, d)
                assert filecmp.cmp(s, d, shallow=False)
                n += 1
    print(f"baseline/_ref: {n} reference files in place (byte-identical copies of {SRC})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
