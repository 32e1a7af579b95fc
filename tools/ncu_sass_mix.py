"""Opcode mix of one kernel from an .ncu-rep captured with --import-source on: executed warp
instructions and stall samples per SASS opcode (which pipe the issue slots go to).
    python tools/ncu_sass_mix.py gpurun_out/prof.ncu-rep [top_n]
"""
import collections
import csv
import io
import subprocess
import sys


def main(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name = rows[0][1]
    hdr = rows[1]
    i_src, i_ex, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ex, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= i_ex:
            continue
        toks = r[i_src].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = ".".join(op.split(".")[:2])
        ex[op] += int(r[i_ex] or 0)
        smp[op] += int(r[i_smp] or 0)
    tot, tots = sum(ex.values()), sum(smp.values())
    print("# %s\n# %s : %d warp instructions, %d stall samples" % (path, name, tot, tots))
    print("%-22s %14s %7s %9s %7s" % ("opcode", "warp_instr", "%", "samples", "%"))
    for op, n in ex.most_common(top):
        print("%-22s %14d %6.1f%% %9d %6.1f%%" % (op, n, 100.0 * n / tot, smp[op], 100.0 * smp[op] / max(tots, 1)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
