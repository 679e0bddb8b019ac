"""Development aid: list the DMMA clusters of a kernel's SASS with the local-memory (spill) traffic inside each.
usage: python tools/sass_loops.py <object.o> <mangled-kernel-name>"""
import subprocess, sys
obj, fun = sys.argv[1:3]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout.split("\n")
idx = [i for i, l in enumerate(sass) if "DMMA" in l]
clusters = []; cur = [idx[0]]
for i in idx[1:]:
    if i - cur[-1] < 60: cur.append(i)
    else: clusters.append(cur); cur = [i]
clusters.append(cur)
for c in clusters:
    seg = sass[max(c[0] - 80, 0): c[-1] + 40]
    cnt = lambda k: sum(k in l for l in seg)
    print(f"lines {c[0]}-{c[-1]}: DMMA {len(c)}, LDL {cnt('LDL')}, STL {cnt('STL')}, LDS {cnt('LDS')}, SYNCS {cnt('SYNCS')}, SHFL {cnt('SHFL')}")
