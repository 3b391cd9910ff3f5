"""Static SASS census of the built library: per kernel instantiation, the tcgen05 / TMA
instructions it contains (proof of which hardware paths a kernel uses) and the density of its
MMA issue code.   python tools/sass_summary.py [libsegb200.so] > profiles/rNN_sass_summary.md"""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'segmentation_b200', 'libsegb200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = subprocess.run(['cuobjdump', '-elf', lib], capture_output=True, text=True).stdout  # unused
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        kern[cur] = []
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(.*?);', line)
    if m and cur:
        kern[cur].append(m.group(1))


def demangle(n):
    r = subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    r = re.sub(r'\(.*', '', r).replace('segb::', '').replace('void ', '')
    return r.replace('(bool)', '').replace('(int)', '')


KEYS = ['UTCHMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UBLKRED', 'UTCATOMSWS', 'LDTM', 'ELECT']
rows = []
for n, ins in kern.items():
    c = {k: sum(1 for i in ins if re.search(r'\b%s' % k, i)) for k in KEYS}
    if not (c['UTCHMMA'] or c['UTMALDG'] or c['UBLKRED'] or c['UTMAREDG']):
        continue
    # densest issue region: fewest instructions spanning a run of consecutive UTCHMMA groups
    idx = [i for i, s in enumerate(ins) if 'UTCHMMA' in s]
    per = None
    if len(idx) >= 8:
        gaps = [b - a for a, b in zip(idx, idx[1:])]
        per = (idx[-1] - idx[0] + 1) / float(len(idx))
    mc = sum(1 for i in ins if 'MULTICAST' in i)
    rows.append((demangle(n), len(ins), c, per, mc))
rows.sort(key=lambda r: r[0])
print('# SASS census of `libsegb200.so` (`tools/sass_summary.py`, cuobjdump -sass, sm_100a)\n')
print('`UTCHMMA` = tcgen05.mma, `UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UTMALDG/STG/REDG` = '
      'TMA tensor load / store / reduce, `UBLKRED` = bulk reduce-add.  "instr / MMA" = static '
      'instructions between the first and the last `UTCHMMA` of the kernel divided by their number '
      '(the unrolled single-thread issue code; lower is denser).\n')
print('| kernel | SASS instr | UTCHMMA | UTCBAR | LDTM | UTMALDG | UTMASTG | UTMAREDG | UBLKRED | multicast | instr / MMA |')
print('|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|')
for name, n, c, per, mc in rows:
    print('| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %s |' % (
        name, n, c['UTCHMMA'], c['UTCBAR'], c['LDTM'], c['UTMALDG'], c['UTMASTG'], c['UTMAREDG'],
        c['UBLKRED'], mc, ('%.1f' % per) if per else '—'))
