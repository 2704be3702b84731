"""Regenerates profiles/sass_grep.txt: per-kernel counts of the SASS mnemonics that prove which hardware units the
shipped library uses (tcgen05 tensor cores, TMA, TMEM), from `cuobjdump -sass` of the in-tree .so.  CPU only.

    python tools/sass_grep.py > profiles/sass_grep.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'margipose_b200', 'libmargipose_b200.so')
COLS = ['UTCHMMA.2CTA', 'UTCHMMA', 'UTMALDG', 'UBLKCP', 'LDTM', 'UTCBAR', 'SYNCS', 'HMMA', 'MUFU.EX2', 'MUFU.LG2',
        'RED.E.ADD', 'ATOMG']


def demangle(names):
    out = subprocess.run(['c++filt'] + names, capture_output=True, text=True, check=True).stdout.splitlines()
    res = []
    for n in out:
        n = n.replace('(anonymous namespace)::', '')
        n = re.sub(r'\((?:bool|int|unsigned int)\)', '', n)     # "(bool)1" -> "1" inside template arguments
        n = re.sub(r'\(.*$', '', n)                # drop the parameter list
        res.append(n)
    return res


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None or '/*' not in line:
            continue
        m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        c = counts[cur]
        if op.startswith('UTCHMMA'):
            c['UTCHMMA.2CTA' if '.2CTA' in op else 'UTCHMMA'] += 1
        elif op.startswith('HMMA'):
            c['HMMA'] += 1
        elif op.startswith('RED.E.ADD'):
            c['RED.E.ADD'] += 1
        else:
            for k in ('UTMALDG', 'UBLKCP', 'LDTM', 'UTCBAR', 'SYNCS', 'MUFU.EX2', 'MUFU.LG2', 'ATOMG'):
                if op.startswith(k):
                    c[k] += 1
    names = demangle(order)
    print('# SASS evidence: tensor-core / TMA / TMEM instructions per kernel of margipose_b200/libmargipose_b200.so')
    print('# (cuobjdump -sass, CUDA 12.9, -gencode arch=compute_100a,code=sm_100a; regenerate: python tools/sass_grep.py)')
    print('# UTCHMMA = tcgen05.mma (kind::f16, bf16 operands), .2CTA = cta_group::2; UTMALDG = cp.async.bulk.tensor (TMA tile')
    print('# load); UBLKCP = cp.async.bulk (TMA 1-D copy, BatchNorm rings); LDTM = tcgen05.ld (TMEM -> registers); UTCBAR =')
    print('# tcgen05.commit -> mbarrier; SYNCS = mbarrier operations; HMMA = legacy mma.sync (must be 0).')
    print()
    print('%-72s' % 'kernel' + ''.join('%13s' % c for c in COLS))
    total = collections.Counter()
    for raw, name in zip(order, names):
        c = counts[raw]
        total.update(c)
        print('%-72s' % name[:72] + ''.join('%13d' % c[k] for k in COLS))
    print('%-72s' % 'TOTAL (all kernels)' + ''.join('%13d' % total[k] for k in COLS))
    if total['HMMA']:
        sys.exit('legacy HMMA found')


if __name__ == '__main__':
    main()
