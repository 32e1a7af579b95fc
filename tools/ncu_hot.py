"""Summarise an ncu --set full capture (--import-source on): per kernel, the opcode mix (instructions executed), the stall
samples by reason, and the hottest SASS lines.  usage: python tools/ncu_hot.py rep.ncu-rep [kernel-index] [top-n]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# split per kernel
blocks, cur = [], None
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
b = blocks[kidx]
print(b[0][:200])
rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
def f(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in data); tot_s = sum(f(r, "# Samples") for r in data)
print("total warp instructions %.3g, samples %d" % (tot_inst, tot_s))
ops = collections.Counter(); ops_s = collections.Counter()
for r in data:
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2])
    ops[op] += f(r, "Instructions Executed"); ops_s[op] += f(r, "# Samples")
print("--- opcode mix (share of executed warp instr | share of samples)")
for op, n in ops.most_common(28):
    print("%-22s %6.2f%%  %6.2f%%" % (op, 100 * n / tot_inst, 100 * ops_s[op] / max(1, tot_s)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("--- stall reasons (all samples)")
st = {h: sum(f(r, h) for r in data) for h in stalls}
for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
    print("%-26s %6.2f%%" % (h, 100 * v / max(1, tot_s)))
print("--- hottest lines by samples")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
    top = sorted(((f(r, h), h) for h in stalls), reverse=True)[:2]
    print("%5.2f%%  exec %9d  %-70s %s" % (100 * f(r, "# Samples") / max(1, tot_s), f(r, "Instructions Executed"), r[ix["Source"]].strip()[:70], " ".join("%s=%d" % (h[6:], v) for v, h in top if v)))
