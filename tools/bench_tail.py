"""BASELINE.json configs[3]: flat_softmax + dsnt (+ xyz combination) fusion and the full fused tail (+ JS + Euclid),
heatmap side 32/64/128, 17 joints, batch 128, three planes -- achieved HBM GB/s vs the measured peak
(the sweep itself lives in bench.py: tail_sweep; bench.py reports it as `tail_roofline`).
    python tools/bench_tail.py [tunable=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from margipose_b200._lib import lib

for kv in sys.argv[1:]:
    k, v = kv.split('=')
    assert lib().mp_set_tunable(k.encode(), int(v)) == 0, kv
print(' '.join(sys.argv[1:]) or '(defaults)')
peak = bench.peaks()[1]
for row in bench.tail_sweep(peak):
    print('S=%3d  ' % row['heatmap'] + '   '.join('%s %7.1f us %5.0f GB/s (%.2f)' % (k, v['us'], v['GB/s'], v['frac'])
                                                   for k, v in row.items() if k != 'heatmap'))
