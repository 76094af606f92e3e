"""Write the four reference-trajectory files (trajectory.txt, trajectory1.txt, xydq.txt, xydq_circle.txt) from
the exact float64 arrays in tests/golden/ref_paths.npz, in the text format the reference ships them in
(np.savetxt default '%.18e', one row per line) — so `np.loadtxt('xydq_circle.txt')` (run.py:18) works in a
fresh checkout.  tests/test_oracle_vs_reference.py compares the output byte for byte with /root/reference.

    python tools/export_ref_paths.py [directory]        (default: the repository root)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("trajectory", "trajectory1", "xydq", "xydq_circle")


def export(directory: str = ROOT) -> list:
    os.makedirs(directory, exist_ok=True)
    out = []
    with np.load(os.path.join(ROOT, "tests", "golden", "ref_paths.npz")) as z:
        for name in NAMES:
            path = os.path.join(directory, name + ".txt")
            np.savetxt(path, z[name])
            out.append(path)
    return out


if __name__ == "__main__":
    for p in export(sys.argv[1] if len(sys.argv) > 1 else ROOT):
        print(p)
