"""sass_hist.py -- instruction histogram of libkmg.so's kernels (cuobjdump -sass), for profiles/."""
import collections
import re
import subprocess
import sys

so = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", ln)
    if m and cur:
        hist[cur][m.group(1)] += 1
names = list(hist)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
BW = ("UTCIMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UCGABAR", "UTCATOMSWS", "UTMACCTL", "UTMAPF")
for f, d in zip(names, dem):
    h = hist[f]
    tot = sum(h.values())
    d = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", d)
    d = re.sub(r"^void ", "", d)
    d = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", d)
    fam = collections.Counter()
    for op, c in h.items():
        fam[op.split(".")[0]] += c
    sp = ", ".join("%s x%d" % (op, c) for op, c in sorted(h.items()) if op.split(".")[0] in BW)
    print("%s  [%d instructions]" % (d[:120], tot))
    if sp:
        print("    blackwell: " + sp)
    print("    top: " + ", ".join("%s %d" % (o, c) for o, c in fam.most_common(9)))
