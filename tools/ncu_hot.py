"""Top stall locations of an `ncu --page source --csv` dump (SASS level)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Address' in r][0]
h = rows[hi]
c = h.index('Warp Stall Sampling (All Samples)')
ex = h.index('Instructions Executed')
stalls = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
out = []
for k, r in enumerate(rows[hi + 1:]):
  try:
    v = float(r[c])
  except (ValueError, IndexError):
    continue
  top = sorted(((float(r[i] or 0), h[i]) for i in stalls), reverse=True)[:2]
  out.append((v, k, r[1].strip()[:70], r[ex], top))
tot = sum(o[0] for o in out) or 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for v, k, s, e, top in sorted(out, reverse=True)[:n]:
  print('%5.1f%% #%-5d %-70s exec=%-8s %s' % (100 * v / tot, k, s, e, ' '.join('%s=%d' % (b, a) for a, b in top if a)))
