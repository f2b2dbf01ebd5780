"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel share table used in profiles/*.md."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    tot[name] += float(r[14]) / 1e6; cnt[name] += 1
total = sum(tot.values())
print("%d launches, %.2f ms in total\n" % (len(rows), total))
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (k, cnt[k], v, 100 * v / total))
