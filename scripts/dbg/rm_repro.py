import sys, os, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import cbl_testutil as util
import cbl_b200
variant = sys.argv[1]
def low_complexity(n, seed):
    rng = np.random.default_rng(seed)
    s = np.full(n, ord("A"), dtype=np.uint8)
    idx = rng.integers(0, n, size=n // 12)
    s[idx] = util.BASES[rng.integers(0, 4, size=len(idx))]
    return s
g = cbl_b200.CBL(25, 64, 24, False)
a = util.random_dna(40000, 11).tobytes()
b = a[10000:30000] + util.random_dna(20000, 12).tobytes()
c = low_complexity(30000, 13).tobytes()
seqs = {'a': a, 'b': b, 'c': c}
ins, rem = variant.split(':')
for ch in ins:
    g.insert_seq(seqs[ch]); print('insert', ch, g.count(), flush=True)
for ch in rem:
    g.remove_seq(seqs[ch]); print('remove', ch, g.count(), flush=True)
print('ok', variant)
