"""(Round-1 tool: it recognises the loop shape of the round-1 kernels — two 30-candidate search blocks that a certified
warp jumps over.  The round-2 kernels are described by the executed-instruction mix of the ncu capture instead,
profiles/r2_rollout_c4_instruction_mix.txt.)

Static view of the rollout kernel's horizon loop from `cuobjdump -sass` (no GPU needed): size of the loop body,
the two search blocks the certified lookups jump over, and the opcode mix of what remains (the path a fully
certified warp executes; the Philox block in it runs every second iteration).

    python tools/sass_loop_mix.py [path/to/libmppi_b200.so] [mangled-name-prefix]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mppi_robotarm_b200", "libmppi_b200.so")
KEY = sys.argv[2] if len(sys.argv) > 2 else "_ZN4mppi19mppi_rollout_sm100aILi0ELb1ELi2ELi0ELb1E"   # philox, const window, NS=2, _F, certified

CLASSES = [("FMA pipe (FFMA, FFMA2, FMUL, FADD)", {"FFMA", "FFMA2", "FMUL", "FADD"}),
           ("compare / select / min", {"FSETP", "FSEL", "FSET", "FMNMX", "SEL", "ISETP"}),
           ("integer (Philox, indices)", {"IMAD", "LOP3", "IADD3", "SHF", "I2FP", "LEA", "VIADD"}),
           ("MUFU", {"MUFU"}),
           ("moves / constant loads", {"MOV", "HFMA2", "LDC", "LDCU", "UMOV", "S2R"}),
           ("shared memory", {"LDS", "STS"}),
           ("control flow / votes", {"BRA", "BSSY", "BSYNC", "VOTE", "WARPSYNC", "NOP"})]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn = [f for f in sass.split("Function : ") if f.startswith(KEY)]
    if not fn:
        sys.exit(f"no function starting with {KEY}")
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4})\*/\s+(.*?);", fn[0])]
    loop = None
    for a, t in ins:                                  # innermost long backward branch = the horizon loop
        m = re.search(r"BRA\s+0x([0-9a-f]+)", t)
        if m and 0x2000 < a - int(m.group(1), 16) < 0x3400 and (loop is None or a - int(m.group(1), 16) < loop[1] - loop[0]):
            loop = (int(m.group(1), 16), a)
    lo, hi = loop
    skips = []
    for a, t in ins:                                  # forward branches over >= 100 instructions inside the loop
        m = re.search(r"^@!?P\d BRA\s+0x([0-9a-f]+)", t)
        if m and lo < a < hi and int(m.group(1), 16) - a >= 0x640:
            skips.append((a, int(m.group(1), 16)))
    search = [s for s in skips if (s[1] - s[0]) // 16 < 130]          # the two 30-candidate searches
    print(f"{fn[0].split()[0]}")
    print(f"horizon loop: {(hi - lo) // 16 + 1} SASS instructions per iteration (two samples per thread)")
    print(f"search blocks skipped by a certified warp: {[(t - a) // 16 - 1 for a, t in search]}")
    other = [s for s in skips if s not in search]
    if other:
        print(f"other forward skips (Philox draw on odd steps): {[(t - a) // 16 - 1 for a, t in other]}")
    ops = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi and not any(s0 < a < s1 for s0, s1 in search):
            op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
            ops[op] += 1
    total = sum(ops.values())
    print(f"certified path: {total} instructions per iteration = {total / 2:.0f} per sample-step (static)")
    rest = dict(ops)
    for name, members in CLASSES:
        n = sum(rest.pop(m, 0) for m in list(members))
        print(f"  {name:38s}{n:5d}  {n / 2:6.1f} per sample-step")
    print(f"  {'other: ' + ', '.join(sorted(rest)):38s}{sum(rest.values()):5d}")


if __name__ == "__main__":
    main()
