#!/usr/bin/env python
"""Map the per-instruction stall samples of an `ncu --set full` capture back to source lines of the .cuh kernel bodies.

ncu's CSV source page only prints the .cu file that was compiled, not the headers the kernel bodies live in, so this joins
  ncu -i <rep> --page source --csv --print-source sass     (address, SASS text, samples, executed instructions, stall columns)
with
  nvdisasm -gi -c <cubin>                                   (offset, SASS text, "//## File ..., line N" annotations)
of the SAME build of the library, by instruction offset inside the kernel's section, and aggregates by function (line ranges of
gusto.jl_b200/csrc/ipm.cuh found by a small parser) and by line.

  python tools/ncu_lines.py <report.ncu-rep> <libgusto_b200.so> <kernel mangled-name substring> [--kernel <regex for a multi-kernel report>] [--top N] [--md out.md]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def disasm(lib, kernel_sub):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    out, inside, cur = {}, False, None
    pending = []
    for line in txt.splitlines():
        if line.startswith(".text."):
            inside = kernel_sub in line
            pending = []
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            if pending:
                cur = pending
                pending = []
            out[int(m.group(1), 16)] = (m.group(2).strip(), cur)
        elif line.startswith("\t.section") or line.startswith("//-----"):
            if out:
                inside = False
    return out


def ncu_rows(rep, kfilter=None):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
    if kfilter:
        cmd += ["-k", "regex:" + kfilter, "-c", "1"]
    raw = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    out = []
    for r in rows[hi + 1:]:
        if r and r[0] in ("Kernel Name", "Address"):      # the next launch of a multi-launch report
            break
        if len(r) == len(hdr):
            out.append(r)
    return hdr, out


def function_ranges(path):
    """(first line, name) of every function-like definition of a header, in order."""
    fr = []
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:GDEV_NOINLINE|GDEV|GHD|__device__ __noinline__|__device__ __forceinline__|__global__)\b.*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(")
    prev = ""
    for i, line in enumerate(open(path), 1):
        s = line.strip()
        m = pat.match(s) or (pat.match(prev + " " + s) if prev.startswith("template") else None)
        if m and not s.endswith(";"):
            fr.append((i, m.group(1)))
        prev = s if s.startswith("template") and "(" not in s else ""
    return fr


def main():
    rep, lib, ksub = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    md = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
    dis = disasm(lib, ksub)
    kf = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else None
    hdr, rows = ncu_rows(rep, kf)
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(rows[0][ia], 16)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    franges = {}
    for f in (x for x in os.listdir(os.path.join(root, "gusto.jl_b200", "csrc")) if x.endswith((".cuh", ".cu"))):
        franges[f] = function_ranges(os.path.join(root, "gusto.jl_b200", "csrc", f))

    def fn_of(fname, ln):
        best = "?"
        for l0, name in franges.get(fname, []):
            if l0 <= ln:
                best = name
            else:
                break
        return best

    by_fn = collections.defaultdict(lambda: collections.Counter())
    by_line = collections.defaultdict(lambda: collections.Counter())
    mism = 0
    tot = collections.Counter()
    for r in rows:
        off = int(r[ia], 16) - base
        ent = dis.get(off)
        samp, ex = int(r[isamp] or 0), int(r[iex] or 0)
        if ent is None or ent[1] is None:
            key_fn, key_line = "(unmapped)", ("?", 0)
            mism += 1
        else:
            chain = ent[1]
            # innermost frame first; attribute to the outermost frame that is still inside a kernel-body header (.cuh), so that
            # small inlined helpers (g_rcp, tile jobs ...) count towards the phase function that called them
            body = [c for c in chain if c[0].endswith(".cuh") and c[0] not in ("common.cuh",)]
            pick = body[-1] if body else chain[0]
            inner = chain[0]
            key_fn = f"{pick[0]}:{fn_of(*pick)}"
            key_line = (pick[0], pick[1])
        for c, s in ((by_fn[key_fn], 1), (by_line[key_line], 1)):
            c["samples"] += samp; c["inst"] += ex
            for i, h in stall_cols:
                v = int(r[i] or 0)
                if v:
                    c[h] += v
        tot["samples"] += samp; tot["inst"] += ex
    lines = []
    lines.append(f"kernel section: {len(dis)} SASS instructions, {len(rows)} profiled rows, {mism} unmapped; total samples {tot['samples']}, "
                 f"warp instructions executed {tot['inst']}")
    lines.append("")
    lines.append("| function | samples % | warp-inst % | top stalls (share of the function's samples) |")
    lines.append("|---|---|---|---|")
    for k, c in sorted(by_fn.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((h, v) for h, v in c.items() if h.startswith("stall_")), key=lambda hv: -hv[1])[:4]
        lines.append(f"| {k} | {100 * c['samples'] / max(1, tot['samples']):.1f} | {100 * c['inst'] / max(1, tot['inst']):.1f} | "
                     + ", ".join(f"{h[6:]} {100 * v / max(1, c['samples']):.0f}%" for h, v in st) + " |")
    lines.append("")
    lines.append("| line | samples % | warp-inst % | top stalls |")
    lines.append("|---|---|---|---|")
    for k, c in sorted(by_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((h, v) for h, v in c.items() if h.startswith("stall_")), key=lambda hv: -hv[1])[:3]
        lines.append(f"| {k[0]}:{k[1]} | {100 * c['samples'] / max(1, tot['samples']):.2f} | {100 * c['inst'] / max(1, tot['inst']):.2f} | "
                     + ", ".join(f"{h[6:]} {100 * v / max(1, c['samples']):.0f}%" for h, v in st) + " |")
    txt = "\n".join(lines)
    print(txt)
    if md:
        open(md, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
