"""Distil an `ncu --page source --csv` export into an opcode mix + the SASS lines with the most stall samples.
    python tools/ncu_hotspots.py gpurun_out/r2_ncu_source_message_tc_fwd.csv [kernel-substring] [top]"""
import collections, csv, sys
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 18
csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(path)))
# the export holds one block per kernel instance: a "Kernel Name" row, a header row, then the SASS lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "lines": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r:
        cur["lines"].append(r)
seen = set()
for b in blocks:
    if want not in b["name"] or b["name"] in seen or not b["lines"]:
        continue
    seen.add(b["name"])
    h = b["hdr"]
    i_src, i_smp, i_exec = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    num = lambda s: float(s.replace(",", "")) if s not in ("", "-") else 0.0
    tot_exec = sum(num(r[i_exec]) for r in b["lines"])
    tot_smp = sum(num(r[i_smp]) for r in b["lines"])
    mix = collections.defaultdict(lambda: [0.0, 0.0])
    for r in b["lines"]:
        toks = r[i_src].split()
        op = next((t for t in toks if not t.startswith("@")), "?").split(".")[0]
        mix[op][0] += num(r[i_exec]); mix[op][1] += num(r[i_smp])
    print("== %s" % b["name"][:150])
    print("   warp instructions executed %d, stall samples %d" % (tot_exec, tot_smp))
    print("   opcode mix (share of executed warp instructions | share of stall samples):")
    for op, (e, s) in sorted(mix.items(), key=lambda kv: -kv[1][0])[:14]:
        print("     %-10s %5.1f %% | %5.1f %%" % (op, 100 * e / max(tot_exec, 1), 100 * s / max(tot_smp, 1)))
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    agg = collections.defaultdict(float)
    for r in b["lines"]:
        for i in stall_cols:
            agg[h[i]] += num(r[i])
    tot = sum(agg.values())
    print("   stall reasons: " + ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    print("   SASS lines with the most stall samples:")
    for r in sorted(b["lines"], key=lambda r: -num(r[i_smp]))[:top]:
        print("     %6d samples %10d exec   %s" % (num(r[i_smp]), num(r[i_exec]), r[i_src][:100]))
    print()
