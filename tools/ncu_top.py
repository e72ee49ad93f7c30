"""Top stall lines per kernel from `ncu -i X.ncu-rep --page source --csv [--kernel-name regex:K]` output (SASS view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
seen = set()
for b in blocks:
    if (b["name"], len(b["data"])) in seen or not b["data"]:      # ncu repeats the kernel header line
        continue
    seen.add((b["name"], len(b["data"])))
    hdr = b["hdr"]; ci = {h: i for i, h in enumerate(hdr)}; data = b["data"]; s = ci["# Samples"]
    tot = sum(float(r[s] or 0) for r in data)
    print("==", b["name"][:100], "| samples", tot, "| instructions", len(data))
    for k, r in enumerate(data):
        r.append(k)
    for r in sorted(data, key=lambda r: -float(r[s] or 0))[:top]:
        stalls = {h[6:]: int(r[ci[h]]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h and int(r[ci[h]] or 0) > 0}
        print(f"{r[-1]:5d} {r[s]:>5} {r[ci['Instructions Executed']]:>8} {r[ci['Source']].strip()[:70]:70s} {stalls}")
