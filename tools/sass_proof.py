"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel takes, from `cuobjdump -sass` of the
shipped library (runs without a GPU).
    python tools/sass_proof.py [lib] > profiles/rNN_sass_proof.txt
A function's block runs from its "Function : <mangled>" line to the next one; every instruction line of the block is counted
(the round-1 generator stopped at the first blank line, which truncated long kernels and reported zeros for them)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UBLKCP", "HMMA", "MUFU.EX2", "LDGSTS", "LDSM", "SYNCS", "FFMA2"]


def main(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    counts, order, cur, k = {}, [], None, 0
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[k].strip() if k < len(names) else m.group(1)
            k += 1
            cur = re.sub(r"\(.*$", "", cur).replace("void lwb::", "").replace("lwb::", "").replace("(anonymous namespace)::", "")
            cur = re.sub(r"\((int|bool|unsigned int)\)", "", cur)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for c in COLS:
            if op == c or op.startswith(c + "."):
                counts[cur][c] += 1
    print("# cuobjdump -sass %s (sm_100a): per-kernel counts of the SASS mnemonics that prove the path" % os.path.relpath(lib, ROOT))
    print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load,")
    print("# UBLKCP = cp.async.bulk (1-D bulk copy), HMMA = legacy mma.sync, LDGSTS = cp.async, FFMA2 = packed fp32x2 FMA")
    w = max(len(n) for n in order) + 2
    print("%-*s" % (w, "kernel") + "".join("%11s" % c for c in COLS))
    for n in sorted(order):
        print("%-*s" % (w, n) + "".join("%11d" % counts[n][c] for c in COLS))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lw-detr_b200", "lib", "liblwdetr_b200.so"))
