"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md): tcgen05.mma -> UTC*MMA,
tcgen05.ld / st -> LDTM / STTM, TMA -> UTMALDG / UBLKCP, plus MUFU and registers, from the built library.

    python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'oriana_b200', 'lib', 'liboriana_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
keys = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'SYNCS', 'MUFU', 'HMMA', 'LDS', 'RED', 'ATOM']
cur, counts, order = None, {}, []
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); order.append(cur); continue
    if cur is None:
        continue
    m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        counts[cur]['_total'] += 1
        for k in keys:
            if op.startswith(k):
                counts[cur][k] += 1
                if k == 'UTCHMMA' and '.2CTA' in op:
                    counts[cur]['UTCHMMA.2CTA'] += 1
dem = subprocess.run(['cu++filt'] + order, capture_output=True, text=True).stdout.splitlines() if order else []
print('library: %s' % os.path.relpath(lib, ROOT))
print('%-92s %6s %8s %8s %5s %5s %8s %7s %5s %5s' % ('kernel', 'instr', 'UTCHMMA', '(.2CTA)', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'MUFU', 'HMMA'))
tot = collections.Counter()
for f, d in zip(order, dem if len(dem) == len(order) else order):
    c = counts[f]
    name = re.sub(r'\(ori::TcMaps, ori::TcArgs\)|\(bool\)|\(int\)', '', d).replace('ori::', '')
    print('%-92s %6d %8d %8d %5d %5d %8d %7d %5d %5d' % (name[:92], c['_total'], c['UTCHMMA'], c['UTCHMMA.2CTA'], c['LDTM'], c['STTM'],
                                                        c['UTMALDG'], c['UBLKCP'], c['MUFU'], c['HMMA']))
    tot.update(c)
print('%-92s %6d %8d %8d %5d %5d %8d %7d %5d %5d' % ('TOTAL', tot['_total'], tot['UTCHMMA'], tot['UTCHMMA.2CTA'], tot['LDTM'], tot['STTM'],
                                                    tot['UTMALDG'], tot['UBLKCP'], tot['MUFU'], tot['HMMA']))
