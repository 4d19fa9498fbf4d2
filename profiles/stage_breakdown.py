"""Per-stage share of executed warp instructions and stall samples of bf_refine_kernel from an ncu report
(`--set full --import-source on`), joined with `nvdisasm -g -c` line info of the cubin that was profiled.

    python profiles/stage_breakdown.py X.ncu-rep boxfusion_b200/lib/bf_refine.o <mangled-kernel-substring> [launch-index]
Stages = the functions of csrc/bf_refine_eval.cuh (innermost inlined location) + the kernel's own phases."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def function_ranges(path):
    out, cur = [], None
    for i, ln in enumerate(open(path), 1):
        m = re.match(r"(?:BF_HD|BF_HD_NOINLINE)\s+\S+\s+(bf_\w+)\(", ln)
        if m:
            out.append([m.group(1), i, 10 ** 9])
            if len(out) > 1:
                out[-2][2] = i - 1
    return out


def main(rep, obj, kern, launch=0):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    addr2line, cur, inside = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith(".text.") and ln.endswith(":"):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S.*);", ln)
        if m:
            addr2line[int(m.group(1), 16)] = cur
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    blocks, cur_rows = [], None
    for r in csv.reader(sass.splitlines()):
        if r and r[0] == "Kernel Name":
            cur_rows = []
            blocks.append(cur_rows)
        elif cur_rows is not None:
            cur_rows.append(r)
    rows = blocks[launch]
    hdr = rows[0]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    fr = function_ranges(os.path.join(ROOT, "boxfusion_b200", "csrc", "bf_refine_eval.cuh"))
    g, gs, tot, tots, base = collections.Counter(), collections.Counter(), 0.0, 0.0, None
    for r in rows[1:]:
        try:
            a = int(r[ia], 16)
        except ValueError:
            continue
        base = a if base is None else base
        n, s = float(r[ii] or 0), float(r[isamp] or 0)
        tot += n; tots += s
        f, l = addr2line.get(a - base) or ("?", 0)
        key = f
        if f == "bf_refine_eval.cuh":
            key = next((nm for nm, lo, hi in fr if lo <= l <= hi), "eval:other")
        elif f == "bf_refine.cu":
            key = "kernel (item loop, leader phase, init)"
        elif "sm_90_rt" in f or "cooperative" in f:
            key = "cluster barrier (cooperative_groups)"
        g[key] += n; gs[key] += s
    print(f"launch {launch}: {tot:.3e} warp instructions, {tots:.0f} stall samples")
    for nm, n in g.most_common(20):
        print(f"  {nm:42s} inst {100 * n / tot:6.2f}%   samples {100 * gs[nm] / max(tots, 1):6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
