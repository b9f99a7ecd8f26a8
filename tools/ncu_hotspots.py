"""Developer script: where the warp-stall samples of an ncu capture fall in the CUDA source.

    cuobjdump -xelf all conflict_rez_b200/libobca_b200.so          # -> obca_api.sm_100a.cubin
    nvdisasm -gi -c obca_api.sm_100a.cubin > disasm_gi.txt         # SASS with (inlined) line info
    ncu -i rep.ncu-rep --page source --csv > src.csv               # per-instruction samples
    python tools/ncu_hotspots.py src.csv disasm_gi.txt [kernel_symbol] [top]

The library must be the build the capture was taken from (the instruction offsets are matched one to one).  Samples are
summed per innermost source line, per outermost line inside the kernel's call tree (phase level) and per function file range.
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, disasm, sym = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "_Z7k_solve9SolveArgs")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# ---- offset -> inline chain [(file, line) innermost first]
chains = {}
cur = []
inside = False
pend = []
for ln in open(disasm, errors="replace"):
    if ln.startswith(sym + ":"):
        inside = True
        continue
    if not inside:
        continue
    if ln.startswith("\t.section") or ln.startswith("//---------------------"):
        if chains:
            break
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2)), m.group(3).split("/")[-1] if m.group(3) else None, int(m.group(4)) if m.group(4) else None))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        if pend:
            ch = [(pend[0][0], pend[0][1])]
            for p in pend:
                if p[2] is not None:
                    ch.append((p[2], p[3]))
            cur = ch
            pend = []
        chains[int(m.group(1), 16)] = (cur, m.group(2))

rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address")
H = rows[hi]
col = {h: i for i, h in enumerate(H)}
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
base = None
tot = 0
inner, outer, per_stall = defaultdict(float), defaultdict(float), defaultdict(lambda: defaultdict(float))
insts = defaultdict(float)
for r in rows[hi + 1:]:
    if len(r) < len(H):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    off = addr - base
    n = float(r[col["# Samples"]] or 0)
    ex = float(r[col["Instructions Executed"]] or 0)
    ch = chains.get(off, ([("?", 0)], r[1]))[0]
    k_in, k_out = ch[0], ch[-1] if len(ch) == 1 else ch[-2] if ch[-1][0] == "obca_api.cu" and len(ch) > 1 else ch[-1]
    tot += n
    inner[k_in] += n
    outer[k_out] += n
    insts[k_in] += ex
    for s in stalls:
        v = float(r[col[s]] or 0)
        if v:
            per_stall[k_in][s] += v
print("total samples %d, instructions matched %d" % (tot, len(chains)))


def show(title, d):
    print("\n== %s" % title)
    for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:top]:
        st = per_stall.get(k, {})
        s3 = ", ".join("%s %.0f%%" % (a.replace("stall_", ""), 100 * b / max(v, 1)) for a, b in sorted(st.items(), key=lambda kv: -kv[1])[:3]) if d is inner else ""
        print("%6.2f %%  %-22s line %-5d  %s" % (100 * v / tot, k[0], k[1], s3))


show("innermost source line", inner)
show("call-site level (line of the enclosing phase function)", outer)
byfile = defaultdict(float)
for k, v in inner.items():
    byfile[k[0]] += v
print("\n== per file")
for k, v in sorted(byfile.items(), key=lambda kv: -kv[1]):
    print("%6.2f %%  %s" % (100 * v / tot, k))
