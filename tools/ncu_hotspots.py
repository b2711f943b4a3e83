"""Per-source-line stall attribution of one `ncu --set full --import-source on` capture.

    python tools/ncu_hotspots.py gpurun_out/shade.ncu-rep strelka_b200/libstrelka_b200.so [--top 30]

The ncu source page lists stall samples per SASS instruction; `nvdisasm -g` on the cubin inside the library
gives the source file:line of every SASS instruction (the library must be the build that was profiled; compile
with -lineinfo).  Both list the instructions of the kernel in the same order, so they are joined by position
(and checked by opcode).  Output: the source lines with the most stall samples, with the dominant stall reasons.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

STALLS = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_branch_resolving", "stall_barrier", "stall_lg", "stall_mio",
          "stall_not_selected", "stall_selected", "stall_no_inst", "stall_dispatch", "stall_membar", "stall_tex", "stall_drain", "stall_sleep",
          "stall_misc"]


def ncu_rows(rep):
    # a saved `ncu -i x.ncu-rep --page source --csv` page works as well as the report itself
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernel = rows[0][1]
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    return kernel, ix, rows[2:]


def disasm_sections(lib):
    with tempfile.TemporaryDirectory(dir=os.path.dirname(os.path.abspath(lib))) as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True, check=True)
        sections = {}
        for f in os.listdir(tmp):
            if not f.endswith(".cubin"):
                continue
            txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
            name, cur = None, None
            for line in txt.splitlines():
                m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
                if m:
                    name = m.group(1)
                    sections[name] = []
                    cur = None
                    continue
                m = re.search(r'//## File "(.*)", line (\d+)', line)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
                if m and name:
                    sections[name].append((m.group(1).strip(), cur))
        return sections


def opcode(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s.strip())
    return s.split()[0].split(".")[0] if s else ""


def main():
    rep, lib = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    kernel, ix, rows = ncu_rows(rep)
    sections = disasm_sections(lib)
    best = None
    for name, ins in sections.items():
        if len(ins) != len(rows):
            continue
        same = sum(opcode(a[0]) == opcode(r[ix["Source"]]) for a, r in zip(ins, rows))
        if best is None or same > best[1]:
            best = (name, same)
    if best is None or best[1] < 0.98 * len(rows):
        sys.exit(f"no section of {lib} matches the {len(rows)} instructions of {kernel}: not the profiled build?")
    ins = sections[best[0]]
    per_line = defaultdict(Counter)
    total = Counter()
    for (text, loc), r in zip(ins, rows):
        c = per_line[loc]
        c["samples"] += int(r[ix["# Samples"]])
        c["inst"] += int(r[ix["Instructions Executed"]])
        c["thread_inst"] += int(r[ix["Thread Instructions Executed"]])
        for s in STALLS:
            if s in ix:
                c[s] += int(r[ix[s]])
    for c in per_line.values():
        total.update(c)
    print(f"# {rep}: {kernel}")
    print(f"# section {best[0]}: {len(rows)} SASS instructions, {total['samples']} samples, {total['inst']} warp instructions, "
          f"{total['thread_inst'] / max(total['inst'], 1):.1f} threads/instruction")
    print("# stall totals: " + ", ".join(f"{s[6:]} {100.0 * total[s] / max(total['samples'], 1):.1f}%" for s in STALLS if total[s] * 100 >= total["samples"]))
    print(f"{'samples':>8} {'%':>6} {'warp-inst%':>10} {'thr/inst':>8}  location                         dominant stalls")
    for loc, c in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        dom = sorted(((c[s], s[6:]) for s in STALLS if c[s]), reverse=True)[:3]
        where = f"{loc[0]}:{loc[1]}" if loc else "?"
        print(f"{c['samples']:8d} {100.0 * c['samples'] / max(total['samples'], 1):6.1f} {100.0 * c['inst'] / max(total['inst'], 1):10.1f} "
              f"{c['thread_inst'] / max(c['inst'], 1):8.1f}  {where:32s} " + ", ".join(f"{n} {100.0 * v / max(c['samples'], 1):.0f}%" for v, n in dom))


if __name__ == "__main__":
    main()
