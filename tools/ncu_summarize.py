"""Turn ncu outputs brought back in gpurun_out/ into the text summaries kept under profiles/.

    python tools/ncu_summarize.py rep   gpurun_out/prof.ncu-rep  profiles/rX_rollout_ncu_summary.txt  "header line"
    python tools/ncu_summarize.py mix   gpurun_out/prof.ncu-rep  profiles/rX_rollout_instruction_mix.txt  K T
    python tools/ncu_summarize.py list  gpurun_out/launches.csv  profiles/rX_launch_list_summary.txt  "header line"
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum(\.per_second)?|dram__cycles_active\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|launch__occupancy_limit_\w+|"
    r"launch__registers_per_thread|launch__grid_size|launch__block_size|sm__cycles_elapsed\.max|"
    r"sm__inst_executed\.sum(\.per_cycle_elapsed)?|smsp__inst_executed\.sum|"
    r"sm__inst_executed_pipe_(fma|fmaheavy|fmalite|alu|xu|lsu|fp64|adu|uniform)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_(alu|fma|fmaheavy)_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.(per_cycle_active|pct_of_peak_sustained_active)|"
    r"smsp__issue_active\.avg\.(per_cycle_active|pct_of_peak_sustained_active)|smsp__warps_eligible\.avg\.per_cycle_active|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__thread_inst_executed_per_inst_executed\.ratio|"
    r"sm__inst_issued\.avg\.per_cycle_active|l1tex__data_bank_conflicts_pipe_lsu\.sum|lts__t_bytes\.sum)$")

CLASSES = [
    ("FP32 arithmetic, FMA pipe (FFMA, FFMA2, FMUL, FADD)", {"FFMA", "FFMA2", "FMUL", "FADD", "FMUL2", "FADD2"}),
    ("compare / select / min (FSETP, FSEL, FSET, FMNMX, SEL)", {"FSETP", "FSEL", "FSET", "FMNMX", "SEL", "FMNMX3"}),
    ("Philox + index integer work (IMAD, LOP3, IADD3, SHF, I2FP, ISETP, LEA)", {"IMAD", "LOP3", "IADD3", "SHF", "I2FP", "ISETP", "LEA", "VIADD", "IADD", "PRMT", "F2I", "I2F"}),
    ("MUFU (rcp, lg2, sqrt, sin, cos)", {"MUFU"}),
    ("moves / constants (MOV, HFMA2, LDC, LDCU, UMOV, ...)", {"MOV", "HFMA2", "LDC", "LDCU", "UMOV", "S2R", "CS2R", "S2UR", "R2UR", "UIMAD", "UISETP", "UPRMT", "USHF", "UIADD3", "ULEA", "ULOP3", "R2P", "P2R", "PLOP3", "UPLOP3"}),
    ("shared / global / local memory (LDS, LDG, STG, STS, LDL, STL)", {"LDS", "LDG", "STG", "STS", "LDL", "STL", "ATOMG", "RED", "LD", "ST"}),
    ("control flow / sync / votes (BRA, BSSY, BSYNC, BAR, VOTE, ...)", {"BRA", "BSSY", "BSYNC", "BAR", "SYNCS", "VOTE", "WARPSYNC", "EXIT", "NOP", "CALL", "RET", "ELECT", "UBLKCP", "FENCE", "MEMBAR", "ERRBAR", "DEPBAR", "BREAK", "SHFL", "REDUX"}),
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def do_rep(rep, dst, header):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    lines = [header, ""]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        lines.append(f"{'Kernel Name':95s} {d.get('Kernel Name', '')}")
        lines.append(f"{'Block Size':95s} {d.get('Block Size', '')}")
        lines.append(f"{'Grid Size':95s} {d.get('Grid Size', '')}")
        for h, u, v in sorted(zip(hdr, units, vals)):
            if KEEP.match(h) and v != "":
                lines.append(f"{h:95s} {v:>18s} {u}")
        lines.append("")
    open(dst, "w").write("\n".join(lines))


def do_mix(rep, dst, K, T):
    rows = ncu_csv(rep, "source")
    first = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, rows = rows[first], rows[first:]
    i_src = hdr.index("Source")
    i_exe = next(i for i, h in enumerate(hdr) if h.strip() in ("Instructions Executed", "# Instructions Executed", "Warp Instructions Executed"))
    ops = collections.Counter()
    for r in rows[1:]:
        try:
            n = int(float(r[i_exe]))
        except (ValueError, IndexError):
            continue
        t = r[i_src].strip()
        if t.startswith("@"):
            t = t.split(None, 1)[1] if " " in t else t
        op = t.split()[0].split(".")[0] if t else "?"
        ops[op] += n
    total = sum(ops.values())
    per = K * T / 32.0
    lines = [f"Executed warp-instructions of the rollout kernel, K={K}, T={T} (ncu --set full, source page of {rep}).", "",
             f"{'group':72s}{'executed':>14s}{'per sample-step':>17s}{'share':>8s}"]
    rest = dict(ops)
    for name, members in CLASSES:
        n = sum(rest.pop(m, 0) for m in list(members))
        lines.append(f"{name:72s}{n:14d}{n / per:17.1f}{100.0 * n / total:7.1f}%")
    n = sum(rest.values())
    lines.append(f"{'other: ' + ', '.join(sorted(rest)):72s}{n:14d}{n / per:17.1f}{100.0 * n / total:7.1f}%")
    lines.append(f"{'total':72s}{total:14d}{total / per:17.1f}{100.0:7.1f}%")
    lines += ["", "note: a warp-instruction covers 32 samples; 'per sample-step' = executed warp-instructions / (K*T/32).", "", "top opcodes:"]
    for op, n in ops.most_common(18):
        lines.append(f"  {op:10s}{n:14d}{n / per:10.1f}")
    open(dst, "w").write("\n".join(lines) + "\n")


def do_list(src, dst, header):
    txt = open(src).read()
    start = txt.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        elif r.get("Metric Unit") in ("ms", "msecond"):
            v *= 1e6
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    step = {k: v for k, v in agg.items() if k.startswith(("mppi::mppi_prepare", "void mppi::mppi_rollout", "mppi::mppi_softmin",
                                                          "mppi::mppi_finalize", "void mppi::mppi_wsum", "mppi::mppi_reduce"))}
    tot = sum(v[1] / v[0] for v in step.values())
    lines = [header, "", f"{'kernel':62s}{'launches':>9s}{'mean ns':>13s}{'share of one step':>19s}"]
    for k, (n, t) in step.items():
        lines.append(f"{k:62s}{n:9d}{t / n:13.0f}{100.0 * (t / n) / tot:18.2f}%")
    lines += ["", "other launches in the run (not part of a step):"]
    for k, (n, t) in agg.items():
        if k not in step:
            lines.append(f"{k[:100]:100s}{n:9d}{t / n:13.0f}")
    open(dst, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    kind = sys.argv[1]
    if kind == "rep":
        do_rep(sys.argv[2], sys.argv[3], sys.argv[4])
    elif kind == "mix":
        do_mix(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]))
    elif kind == "list":
        do_list(sys.argv[2], sys.argv[3], sys.argv[4])
