#!/usr/bin/env python
"""Evidence listing from the SHIPPED library: for every kernel in rogue-gym_b200/librogue_b200.so, the architecture,
registers / shared memory / stack from the ELF, the SASS instruction total and the counts of the opcodes that show how
the code is built (bulk-copy engine = 1-D TMA, mbarrier, warp votes / shuffles / reductions, bit tricks, atomics).

    python tools/sass_opcodes.py > profiles/sass_opcodes_r2.txt

Needs cuobjdump (CUDA toolkit); no GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rogue-gym_b200", "librogue_b200.so")
GROUPS = collections.OrderedDict([
    ("UBLKCP (cp.async.bulk: TMA 1-D)", r"^UBLKCP"), ("SYNCS (mbarrier)", r"^SYNCS"), ("UTMA* (tensor maps)", r"^UTMA"),
    ("VOTE / VOTEU", r"^VOTEU?"), ("SHFL", r"^SHFL"), ("REDUX (warp reduce)", r"^REDUX"), ("MATCH", r"^MATCH"),
    ("POPC / BREV / FLO", r"^(POPC|BREV|FLO)"), ("PRMT (byte permute)", r"^PRMT"), ("LOP3", r"^LOP3"),
    ("SHF (funnel shift)", r"^SHF"), ("IMAD / IADD3", r"^(IMAD|IADD3)"), ("LDG / STG", r"^(LDG|STG)"), ("LDS / STS", r"^(LDS|STS)"),
    ("LDL / STL (local: spills)", r"^(LDL|STL)"), ("ATOMG / RED / ATOMS", r"^(ATOMG|ATOM|RED|ATOMS)"),
    ("MEMBAR / FENCE / ERRBAR", r"^(MEMBAR|FENCE|ERRBAR)"), ("BAR / WARPSYNC", r"^(BAR|WARPSYNC)"), ("CALL / RET", r"^(CALL|RET)"),
    ("HMMA / IMMA / UTC*MMA (tensor cores)", r"^(HMMA|IMMA|DMMA|UTCHMMA|UTCIMMA|UTCMMA)"),
])


def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True, check=True).stdout


def main():
    elf = run("cuobjdump", "-lelf", LIB)
    print("library:", os.path.relpath(LIB, ROOT))
    print("cubins:", ", ".join(sorted(set(re.findall(r"ELF file\s+\d+:\s+(\S+)", elf)))))
    print("architectures:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", elf)))))
    res = run("cuobjdump", "-res-usage", LIB)
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(m.group(i)) for i in range(2, 6))
    sass = run("cuobjdump", "-sass", LIB)
    kernels = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = sorted(kernels, key=lambda k: -sum(kernels[k].values()))
    print("\n%-34s %6s %5s %6s %6s %8s" % ("kernel (demangled stem)", "instr", "regs", "stack", "smem", "local"))
    for k in names:
        stem = re.sub(r"^_ZN2rg\d+", "", k)
        stem = re.split(r"ENS_|EPK|EP|Ei|Ev", stem)[0]
        u = usage.get(k, (0, 0, 0, 0))
        print("%-34s %6d %5d %6d %6d %8d" % (stem[:34], sum(kernels[k].values()), u[0], u[1], u[2], u[3]))
    print("\nopcode groups per kernel (static SASS counts)")
    hdr = ["kernel"] + [g.split(" ")[0] for g in GROUPS]
    print(" ".join("%-12s" % h[:12] for h in hdr))
    for k in names:
        stem = re.split(r"ENS_|EPK|EP|Ei|Ev", re.sub(r"^_ZN2rg\d+", "", k))[0]
        row = [stem[:12]]
        for g, pat in GROUPS.items():
            row.append(str(sum(v for op, v in kernels[k].items() if re.match(pat, op))))
        print(" ".join("%-12s" % c for c in row))
    print("\nlegend:")
    for g in GROUPS:
        print("  %-12s %s" % (g.split(" ")[0], g))
    tot = collections.Counter()
    for k in kernels:
        tot.update(kernels[k])
    tc = sum(v for op, v in tot.items() if re.match(GROUPS["HMMA / IMMA / UTC*MMA (tensor cores)"], op))
    print("\ntensor-core instructions in the library: %d (integer grid work: none expected)" % tc)


if __name__ == "__main__":
    main()
