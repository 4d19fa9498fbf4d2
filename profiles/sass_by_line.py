"""Join an ncu SASS-page CSV (`ncu -i X.ncu-rep --page source --csv --print-source sass`) with `nvdisasm -g -c`
line info of the same cubin: per source line -> executed warp instructions and stall samples."""
import collections
import csv
import re
import sys


def main(sass_csv, disasm, kernel_mangled, top=40):
    # address -> source line from nvdisasm
    addr2line, cur, inside = {}, None, False
    for ln in open(disasm):
        if ln.startswith(".text." + kernel_mangled + ":"):
            inside = True
            continue
        if inside and ln.startswith("//--------------------- .text.") and kernel_mangled not in ln:
            break
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S.*);", ln)
        if m:
            addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    per_line = collections.defaultdict(lambda: [0.0, 0.0])
    per_op = collections.defaultdict(lambda: [0.0, 0.0])
    tot_i = tot_s = 0.0
    for r in rows[2:]:
        try:
            a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        except ValueError:
            continue
        if base is None:
            base = a
        off = a - base
        n, s = float(r[ii] or 0), float(r[isamp] or 0)
        tot_i += n; tot_s += s
        line, op = addr2line.get(off, ((None, 0, False), "?"))
        key = (line[0], line[1]) if line else ("?", 0)
        per_line[key][0] += n; per_line[key][1] += s
        opname = op.split()[0] if not op.startswith("@") else op.split()[1]
        per_op[opname.split(".")[0]][0] += n; per_op[opname.split(".")[0]][1] += s
    print(f"total warp instructions {tot_i:.3e}, stall samples {tot_s:.0f}")
    print("---- by source line (innermost inlined location)")
    for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{str(k[0])[:18]:18s}:{k[1]:4d}  inst {100 * v[0] / tot_i:6.2f}%  samples {100 * v[1] / tot_s:6.2f}%")
    print("---- by opcode")
    for k, v in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:30]:
        print(f"{k:12s} inst {100 * v[0] / tot_i:6.2f}%  samples {100 * v[1] / tot_s:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
