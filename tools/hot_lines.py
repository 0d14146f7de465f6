"""Aggregate `ncu --page source --print-source cuda,sass --csv` (gzip ok) per CUDA source line.
usage: python tools/hot_lines.py <source.csv[.gz]> <out.txt>"""
import collections, csv, gzip, sys

csv.field_size_limit(10**9)
src, dst = sys.argv[1], sys.argv[2]
fh = gzip.open(src, "rt") if src.endswith(".gz") else open(src)
cur = None; agg = collections.Counter(); txt = {}; tot = 0; last = None; hdr = None
stall = collections.Counter()
for r in csv.reader(fh):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) >= 5 and r[0] == "Line No":
        hdr = r; continue
    if len(r) > 6 and cur:
        if r[0]:
            try:
                last = (cur, int(r[0])); txt[last] = r[1]
            except ValueError:
                continue
        try:
            s = int(r[4] or 0)
        except ValueError:
            s = 0
        if last:
            agg[last] += s; tot += s
        if hdr and len(r) == len(hdr):
            for i, h in enumerate(hdr):
                if h.startswith("stall_") and "(Not Issued)" not in h:
                    try:
                        stall[h] += int(r[i] or 0)
                    except ValueError:
                        pass
with open(dst, "w") as out:
    out.write("# ncu --set full --import-source on, lm_kernel, warp-stall samples aggregated per CUDA source line (top 40)\n")
    out.write("total samples %d\n" % tot)
    for k, v in agg.most_common(40):
        out.write("%8d %5.1f%%  %s:%d  %s\n" % (v, 100 * v / max(tot, 1), k[0], k[1], txt[k].strip()[:110]))
    out.write("\n# stall reasons (all samples)\n")
    s = sum(stall.values())
    for k, v in stall.most_common(10):
        out.write("%-26s %5.1f%%\n" % (k, 100 * v / max(s, 1)))
print(open(dst).read()[:3500])
