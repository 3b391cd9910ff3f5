"""Per-launch timeline comparison: python tools/tl_diff.py a.json b.json [filter]"""
import json, sys
a = json.load(open(sys.argv[1])); b = json.load(open(sys.argv[2]))
flt = sys.argv[3] if len(sys.argv) > 3 else ''
ta = tb = 0.0
for (n1, t1, m1), (n2, t2, m2) in zip(a, b):
    if flt in n1:
        print('%-24s %-10s %7.1f %7.1f  %+6.1f' % (n1, t1, m1 * 1e3, m2 * 1e3, (m2 - m1) * 1e3))
        ta += m1; tb += m2
print('sum %.1f %.1f' % (ta * 1e3, tb * 1e3))
