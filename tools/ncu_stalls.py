"""Top stall sites per kernel from an `ncu --page source --csv` export (SASS view with warp-stall samples).
usage: python tools/ncu_stalls.py source.csv [top_n]"""
import csv, sys

top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
sections, cur = [], None
for r in rows:
    if r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(r)
for si, s in enumerate(sections):
    h = s["hdr"]
    isrc, ismp, iexe = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[ismp] or 0) for r in s["rows"])
    print("== kernel %d: %s   total samples %d" % (si, s["name"][:70], tot))
    order = sorted(range(len(s["rows"])), key=lambda i: -int(s["rows"][i][ismp] or 0))[:top_n]
    for i in order:
        r = s["rows"][i]
        n = int(r[ismp] or 0)
        st = sorted(((int(r[j] or 0), c) for j, c in stall_cols), reverse=True)[:2]
        print("  %5.1f%%  line %5d  exec %8s  %-58s %s" % (100.0 * n / max(tot, 1), i, r[iexe], r[isrc].strip()[:58],
                                                          ", ".join("%s=%d" % (c[6:], v) for v, c in st if v)))
