"""Per-instruction listing + stall summary of one kernel from an .ncu-rep (source page, SASS view).

    python tools/ncu_regions.py rep.ncu-rep [out_listing.txt]
"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iSm = hdr.index('# Samples')
iT = hdr.index('Avg. Predicated-On Threads Executed')
warps = int(data[0][iE])
tot = sum(int(r[iE]) for r in data); ts = sum(int(r[iSm]) for r in data)
lines = []
for n, r in enumerate(data):
    lines.append(f'{n:5d} {int(r[iE])/warps:6.2f} {float(r[iT]):5.1f} {r[iSm]:>6s}  {r[iS].strip()}')
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write("\n".join(lines) + "\n")
print('warps', warps, 'instr/warp %.1f' % (tot / warps), 'samples', ts)
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[hdr.index(h)] or 0) for r in data) for h in stall}
print(' '.join(f'{k[6:]}={v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
