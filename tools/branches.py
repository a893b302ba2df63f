"""Hot taken branches of a step-kernel capture: python tools/branches.py sass.csv src.csv  (ncu --page source --csv, --print-source=sass and =cuda,sass)"""
import csv, sys
sass, srcf = sys.argv[1:3]
rows = list(csv.reader(open(srcf)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
h = rows[hi]; iA = h.index('Address')
addr2line = {}; cur = None; curfile = None
for i, r in enumerate(rows):
    if r and r[0] == 'File Path': curfile = r[1]
    if i <= hi: continue
    if r[0].isdigit() and int(r[0]) > 0: cur = (curfile, int(r[0]))
    elif r[0] == '' and len(r) > iA and r[iA].startswith('0x'): addr2line[r[iA]] = cur
src = open('/root/repo/hlynr_intercept_b200/csrc/hlynr_device.cuh').read().split('\n')
s = list(csv.reader(open(sass)))
hh = s[1]; iE = hh.index('Instructions Executed'); iN = hh.index('stall_no_inst'); iS = hh.index('# Samples')
d = []
for r in s[2:]:
    try: d.append((r[0], r[1].strip(), int(r[iE]), int(r[iN]), int(r[iS])))
    except Exception: pass
W = max(x[2] for x in d[:50])
tot = sum(x[4] for x in d); ni = sum(x[3] for x in d)
print(f"warps {W}, SASS {len(d)}, inst/warp {sum(x[2] for x in d)/W:.0f}, samples {tot}, no_inst {ni} ({ni/tot*100:.1f}%)")
for k, (a, t, e, n_, sm) in enumerate(d[:-1]):
    op = t.split()[1] if t.startswith('@') else t.split()[0]
    if op.startswith('BRA') and e >= 0.9 * W:
        taken = max(0, e - d[k + 1][2]) / W
        if taken > 0.1:
            tgt = t.split()[-1]
            ti = [j for j, x in enumerate(d) if x[0] == tgt]
            skip = (ti[0] - k - 1) if ti else None
            ln = addr2line.get(a)
            txt = src[ln[1] - 1].strip()[:100] if ln and ln[0] and ln[0].endswith('hlynr_device.cuh') else str(ln)
            tn = d[ti[0]][3] if ti else -1
            print(f"idx {k:5d} taken {taken:.2f} skip {skip} target_no_inst {tn} | L{ln[1] if ln else '?'}: {txt}")
