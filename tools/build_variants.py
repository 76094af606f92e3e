"""Build A/B variants of libmppi_b200.so into build/variants/ (git-ignored; travels to the GPU box).
    python tools/build_variants.py name=flag,flag ...      e.g.  v1=-DMPPI_KAHAN_MASK=3 v2=-DMPPI_KAHAN_MASK=3,-DMPPI_CW2
Prints registers / spills of the Philox certified rollout kernels of each variant."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mppi_robotarm_b200 import build as b  # noqa: E402


def main():
    out = os.path.join(ROOT, "build", "variants")
    os.makedirs(out, exist_ok=True)
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition("=")
        flags = [f for f in flags.split(",") if f]
        lib = os.path.join(out, name + ".so")
        cmd = [b.find_nvcc(), "-Xptxas=-v", *b.NVCC_FLAGS, *flags, "-o", lib, *b.SOURCES]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            print(res.stderr[-3000:])
            sys.exit(f"variant {name} failed")
        info = []
        blocks = res.stderr.split("Compiling entry function")
        for blk in blocks:
            m = re.search(r"mppi_rollout_sm100aILi0ELb0ELi([1-4])ELi0ELb1E", blk)
            if m:
                regs = re.search(r"Used (\d+) registers", blk).group(1)
                sp = re.search(r"(\d+) bytes spill stores", blk).group(1)
                info.append(f"NS={m.group(1)}: {regs} regs, {sp} B spills")
        print(f"{name}: {' '.join(flags) or '(default)'} -> {'; '.join(sorted(info))}")


if __name__ == "__main__":
    main()
