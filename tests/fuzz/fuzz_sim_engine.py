"""Long-running fuzz (not collected by pytest): random batches through the SIMULATED engine (tests/sim: the product's
engine.cu + kernels.cuh on the host) against the oracle -- random sequence sets, adapter sets, scoring schemes, API mode
(cross product / pair list / multi submit), pipeline chunk sizes and every engine option in random combination.

    python tests/fuzz/fuzz_sim_engine.py <seed> <seconds>

Round 1: seeds 21-24 x 2400 s (130 655 batches, 12.4 M alignments) before the pair profile existed, seeds 31-34 x
1500-1800 s with it (102 860 batches, 9.8 M alignments), seeds 51-53 x 1200 s with the per-group profile (47 952 batches,
4.5 M alignments): 0 mismatches."""
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), 'sim')]
import numpy as np                     # noqa: E402
import sim_engine                      # noqa: E402
from helpers import oracle_batch       # noqa: E402
from test_emulation import _gen, _mut, SCHEMES      # noqa: E402

seed, seconds = int(sys.argv[1]), float(sys.argv[2])
rng = random.Random(seed)
W = sim_engine.load()
DEFAULTS = {'h2d_pack': 0, 'tight_window': 1, 'profile': 1, 'direct_max': 512, 'chunk_tasks': 131072,
            'scratch_mb': 128, 'hbuf': 'auto'}


def rand_scheme():
    if rng.random() < 0.6:
        return list(rng.choice(SCHEMES))
    while True:
        ma = rng.randint(0, 12); mi = rng.randint(-25, ma); go = rng.randint(-30, 1); ge = rng.randint(-30, 1)
        if max(abs(ma), abs(mi), abs(go), abs(ge)) >= 1:
            return [ma, mi, go, ge]


def rand_batch():
    kind = rng.choice(['windows', 'windows', 'mixed', 'long'])
    n_ad = rng.choice([1, 1, 2, 2, 3, 5, 8])
    mmax = rng.choice([24, 28, 32, 50, 111, 200, 300 if rng.random() < 0.1 else 64])
    al = rng.choice(['ACGT', 'ACGT', 'ACGTN', 'AC'])
    ads = [''.join(rng.choice(al) for _ in range(rng.randint(0 if rng.random() < 0.03 else 1, mmax))) for _ in range(n_ad)]
    n = rng.randint(1, 60 if kind != 'long' else 10)
    seqs = []
    for _ in range(n):
        L = {'windows': lambda: rng.choice([150, 150, rng.randint(0, 150)]), 'mixed': lambda: rng.randint(0, 700),
             'long': lambda: rng.randint(400, 5000)}[kind]()
        s = ''.join(rng.choice(al) for _ in range(L))
        for _ in range(rng.randint(0, 2)):
            a = rng.choice(ads)
            if a:
                p = rng.randint(0, len(s))
                ins = _mut(rng, a, al)
                s = (s[:p] + ins + s[p:])[:max(L, len(ins))] if kind == 'windows' else s[:p] + ins + s[p:]
        if rng.random() < 0.05:
            s = s.lower()
        seqs.append(s)
    return seqs, ads


iters = checked = 0
t0 = time.time()
while time.time() - t0 < seconds:
    iters += 1
    sc = rand_scheme()
    opts = {}
    for k, vals in (('h2d_pack', [1]), ('tight_window', [0]), ('profile', [0]),
                    ('direct_max', [50, 100, 300, 100000]), ('chunk_tasks', [1, 7, 40, 300]), ('scratch_mb', [1]),
                    ('hbuf', ['global', 'smem'])):
        if rng.random() < 0.3:
            opts[k] = rng.choice(vals)
    mode = rng.choice(['cross', 'cross', 'pairs', 'multi'])
    try:
        for k, v in opts.items():
            W.set_option(k, v)
        if mode == 'multi':
            batches = []
            for _ in range(rng.randint(1, 4)):
                seqs, ads = rand_batch()
                batches.append(W.pack_sequences(seqs) + W.pack_sequences(ads, offset_dtype=np.int32))
            got = W.adapter_alignment_batch_multi(batches, sc)
            for b, g in zip(batches, got):
                exp = oracle_batch(b[0], b[1], b[2], b[3], sc)
                assert np.array_equal(g, exp), ('multi', seed, iters, sc, opts)
                checked += len(exp)
        else:
            seqs, ads = rand_batch()
            sb, so = W.pack_sequences(seqs)
            ab, ao = W.pack_sequences(ads, offset_dtype=np.int32)
            if mode == 'pairs':
                k = rng.randint(1, 3 * len(seqs))
                ps = np.array([rng.randrange(len(seqs)) for _ in range(k)], dtype=np.int32)
                pa = np.array([rng.randrange(len(ads)) for _ in range(k)], dtype=np.int32)
                got = W.adapter_alignment_batch(sb, so, ab, ao, sc, ps, pa)
                exp = oracle_batch(sb, so, ab, ao, sc, ps, pa)
            else:
                got = W.adapter_alignment_batch(sb, so, ab, ao, sc)
                exp = oracle_batch(sb, so, ab, ao, sc)
            if not np.array_equal(got, exp):
                bad = np.nonzero((got != exp).any(axis=1))[0][:3]
                print('MISMATCH', seed, iters, mode, sc, opts, 'records', bad.tolist(), got[bad].tolist(), exp[bad].tolist(), flush=True)
                print(repr(seqs), repr(ads), flush=True)
                sys.exit(1)
            checked += len(exp)
    finally:
        for k in opts:
            W.set_option(k, DEFAULTS[k])
print('seed', seed, 'iterations', iters, 'alignments', checked, 'mismatches 0', 'sec %.0f' % (time.time() - t0), flush=True)
