#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref/ (git-ignored; travels to the GPU box with gpurun).

dome272/Paella ships no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` has nothing to
build: the install is a verbatim file copy of the reference's importable Python (src/, utils/) — byte for byte, checked
below.  Nothing under baseline/_ref is product source and nothing in paella_b200/ imports it; only
`bench.py --impl reference` / `--impl reference-cuda` (the baseline arms) and tests that pin the oracle load it.

    python baseline/install_ref.py            # no-op (exit 0) when /root/reference is absent and _ref exists
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PAELLA_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["src/modules.py", "src/utils.py", "src/vqgan.py", "utils/modules.py", "utils/alter_attention.py",
         "src_distributed/utils.py", "LICENSE"]


def install() -> bool:
    if not os.path.isdir(os.path.join(REF, "src")):
        return os.path.isdir(DST)
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
        assert filecmp.cmp(src, dst, shallow=False), rel
    return True


if __name__ == "__main__":
    ok = install()
    print(f"baseline/_ref {'ready' if ok else 'NOT available (no reference tree here)'}")
    sys.exit(0)
