"""Where are the local-memory (spill / stack) instructions of a kernel?  usage: sass_spills.py all.sass <function substring>
all.sass = nvdisasm --print-line-info <cubin>."""
import re, sys
from collections import Counter
path, pat = sys.argv[1], sys.argv[2]
inside = False
cur = None
c, tot = Counter(), Counter()
for line in open(path):
    if line.startswith(".text."):
        inside = pat in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.search(r"^\s+/\*[0-9a-f]{4,}\*/", line):
        tot[cur] += 1
        if re.search(r"\b(LDL|STL)\b", line):
            c[cur] += 1
print("instructions", sum(tot.values()), "local ld/st", sum(c.values()))
for k, v in c.most_common(30):
    print(k, v)
