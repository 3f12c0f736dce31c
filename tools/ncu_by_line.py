#!/usr/bin/env python
"""Attributes an ncu `--page source --csv` SASS dump to CUDA source lines and functions.

  ncu -i prof.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all librogue_b200.so; nvdisasm -g -c rg_kernels.sm_100a.cubin > k.sass
  python tools/ncu_by_line.py sass.csv k.sass <mangled kernel name> [top_n] [substring of the instance's kernel name]

Joins on instruction offset (nvdisasm `/*0040*/` vs ncu address - first address); prints the
share of warp-stall samples and executed instructions per source line and per enclosing
function (functions found by scanning the source files for `name(` definitions at column 0)."""
import csv
import re
import sys
from collections import defaultdict


def parse_sass(path, kernel):
    off2line = {}
    cur = None
    active = False
    for ln in open(path, errors="replace"):
        if ln.startswith(".text."):
            active = ln.strip().rstrip(":") == ".text." + kernel
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            off2line[int(m.group(1), 16)] = cur
    return off2line


def function_table(path):
    starts = []
    try:
        lines = open(path, errors="replace").read().split("\n")
    except OSError:
        return starts
    for i, ln in enumerate(lines, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|RG_DEV|static|inline|__noinline__|\s)*[\w:<>\*&\s]+?\b(\w+)\s*\([^;]*$", ln)
        if m and not ln.startswith((" ", "\t", "//", "#", "}")) and ("__device__" in ln or "RG_DEV" in ln or "__global__" in ln):
            starts.append((i, m.group(1)))
    return starts


def main():
    sass_csv, disasm, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    want = sys.argv[5] if len(sys.argv) > 5 else None  # substring of the "Kernel Name" row of the instance to take
    off2line = parse_sass(disasm, kernel)
    rows = list(csv.reader(open(sass_csv)))
    # several kernel instances may be concatenated: take the first whose name matches
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    name_i = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    by_line = defaultdict(lambda: [0, 0, defaultdict(int)])
    total_s = total_i = 0
    stall_cols = None
    for k, h in enumerate(hdr_i):
        end = name_i[name_i.index(h - 1) + 1] if (h - 1) in name_i and name_i.index(h - 1) + 1 < len(name_i) else len(rows)
        if want and not any(want in c for c in rows[h - 1]):
            continue
        hdr = rows[h]
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(j, c) for j, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        base = None
        for r in rows[h + 1:end]:
            if len(r) <= ii or not r[0].startswith("0x"):
                continue
            a = int(r[0], 16)
            if base is None:
                base = a
            key = off2line.get(a - base)
            s, n = int(r[si] or 0), int(r[ii] or 0)
            by_line[key][0] += s
            by_line[key][1] += n
            for j, c in stall_cols:
                v = int(r[j] or 0)
                if v:
                    by_line[key][2][c] += v
            total_s += s
            total_i += n
        break  # first instance only
    tables = {}
    by_fn = defaultdict(lambda: [0, 0, defaultdict(int)])
    for key, (s, n, st) in by_line.items():
        fn = "?"
        if key:
            if key[0] not in tables:
                tables[key[0]] = function_table(key[0])
            for ln, name in tables[key[0]]:
                if ln <= key[1]:
                    fn = name
        by_fn[fn][0] += s
        by_fn[fn][1] += n
        for c, v in st.items():
            by_fn[fn][2][c] += v
    print("kernel %s: %d samples, %d warp instructions" % (kernel, total_s, total_i))
    print("\n== by function (samples%, instr%, top stalls)")
    for fn, (s, n, st) in sorted(by_fn.items(), key=lambda kv: -kv[1][0])[:top]:
        tops = ", ".join("%s %.0f%%" % (c[6:], 100.0 * v / max(s, 1)) for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%-28s %6.2f%% %6.2f%%   %s" % (fn, 100.0 * s / max(total_s, 1), 100.0 * n / max(total_i, 1), tops))
    print("\n== by line")
    for key, (s, n, st) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        tops = ", ".join("%s %.0f%%" % (c[6:], 100.0 * v / max(s, 1)) for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        loc = "%s:%d" % (key[0].split("/")[-1], key[1]) if key else "?"
        print("%-24s %6.2f%% %6.2f%%   %s" % (loc, 100.0 * s / max(total_s, 1), 100.0 * n / max(total_i, 1), tops))


if __name__ == "__main__":
    main()
