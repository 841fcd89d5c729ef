"""Recipe for oracle/_ref (TEST INFRASTRUCTURE, git-ignored, travels to the GPU box with the snapshot).

The reference is pure Python, so "building" it means making its importable package available where
/root/reference does not exist: this script copies the reference's `nerfstudio/` Python package (only *.py, unmodified)
from $NERFSTUDIO_REFERENCE (default /root/reference) into oracle/_ref/nerfstudio/.  Nothing under oracle/_ref is ever
committed, imported by the product package, or timed as the product; `tests/` use it to run the UNMODIFIED reference
models behind `nerfstudio_b200.integration.install()` on the GPU, and `bench.py --impl reference` uses it as the CPU arm.

    python oracle/make_ref.py            # no-op (exit 0) when the reference tree is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def main() -> int:
    src_root = os.environ.get("NERFSTUDIO_REFERENCE", "/root/reference")
    src = os.path.join(src_root, "nerfstudio")
    if not os.path.isdir(src):
        print(f"make_ref: {src} not present; keeping whatever oracle/_ref already holds")
        return 0
    dst = os.path.join(DST, "nerfstudio")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, src)
        for f in files:
            if f.endswith(".py"):
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel, f))
                n += 1
    with open(os.path.join(DST, "SOURCE"), "w") as fh:
        fh.write(f"copied from {src} by oracle/make_ref.py ({n} files); unmodified; not part of the product\n")
    print(f"make_ref: {n} files -> {dst}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
