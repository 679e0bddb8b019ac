"""Join the SASS-level stall samples of an ncu report with nvdisasm line info of the built object (development aid).
usage: python tools/ncu_lines.py <report.ncu-rep> <object.o> <mangled-kernel-substring> [topN]"""
import csv, collections, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(obj)} > /dev/null && nvdisasm -g *.cubin > dis.txt", shell=True, check=True)
lines = open(os.path.join(tmp, "dis.txt")).read().split("\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
# several instantiations can match the name (e.g. mh_lanes_kernel<0> and <1>): take the one whose instruction count is the report's
insts = None
for start in [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l]:
    cand = []; cur = None
    for l in lines[start + 1:]:
        if l.startswith("//---------------------"): break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l): cand.append(cur)
    if len(cand) == len(data): insts = cand; break
assert insts is not None, "no instantiation with %d instructions" % len(data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
agg = collections.Counter(); st = collections.defaultdict(collections.Counter); byfile = collections.Counter()
for k in range(len(data)):
    n = int(data[k][ci["# Samples"]]); agg[insts[k]] += n
    for h in stalls: st[insts[k]][h] += int(data[k][ci[h]])
tot = sum(agg.values())
srcdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "causalgpslc.jl_b200", "csrc")
src = {f: open(os.path.join(srcdir, f)).read().split("\n") for f in os.listdir(srcdir)}
print("total samples", tot)
for key, c in agg.most_common(topn):
    if key is None: continue
    f, ln = key
    text = src[f][ln - 1].strip()[:88] if f in src and ln - 1 < len(src[f]) else ""
    top = ", ".join(f"{k[6:]}:{v}" for k, v in st[key].most_common(3))
    print(f"{100*c/tot:5.1f}% {f}:{ln:4d} {text}   [{top}]")
