"""Instruction mix and hottest SASS regions of one kernel from `ncu -i X.ncu-rep --page source --csv` (all kernels) output.
Usage: sass_mix.py all_src.csv <substring of the kernel name>"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
sec, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; sec.append(cur); continue
    if cur is not None:
        cur["rows"].append(r)
for s in sec:
    if sys.argv[2] not in s["name"] or len(s["rows"]) < 2:
        continue
    hdr = s["rows"][0]
    if "Instructions Executed" not in hdr:
        continue
    ii, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    data = [(int(r[ii]), r[isrc].strip(), int(r[ist]) if r[ist].isdigit() else 0) for r in s["rows"][1:] if len(r) > ii and r[ii].isdigit()]
    tot, stot = sum(d[0] for d in data), sum(d[2] for d in data)
    print(s["name"][:70], "| SASS instrs", len(data), "| executed", tot, "| stall samples", stot)
    ops, st = collections.Counter(), collections.Counter()
    for n, src, sm in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0]
        ops[op] += n; st[op] += sm
    for op, n in ops.most_common(26):
        print("  %-10s %5.1f%% instr   %5.1f%% stall samples" % (op, 100 * n / tot, 100 * st[op] / max(stot, 1)))
    w = 50
    best = sorted(((sum(d[0] for d in data[i:i + w]), sum(d[2] for d in data[i:i + w]), i) for i in range(0, len(data), w)), reverse=True)[:8]
    for n, sm, i in best:
        print("  region sass[%d:%d] %5.1f%% instr %5.1f%% stalls | %s" % (i, i + w, 100 * n / tot, 100 * sm / max(stot, 1), data[i][1][:50]))
    break
