"""Long-running fuzz (not collected by pytest): the C restatement (oracle/liboracle.so) against the reference's own C++
(oracle/_ref/cpp_functions.so) with random scoring schemes, incl. positive gap scores and |score| up to 1000.

    python tests/fuzz/fuzz_oracle.py <seed> <iterations>

Round 1: seeds 301, 302 x 200 000 pairs: 0 mismatches."""
import sys, random, time
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import abi_string, oracle_lib, ref_lib
from test_emulation import _gen
seed = int(sys.argv[1]); iters = int(sys.argv[2])
rng = random.Random(seed)
ol, rl = oracle_lib(), ref_lib()
bad = 0; t = time.time()
for it in range(iters):
    r = rng.random()
    if r < 0.6:
        sc = [rng.randint(0, 12), rng.randint(-25, 3), rng.randint(-30, 0), rng.randint(-30, 0)]
    elif r < 0.9:
        sc = [rng.randint(-5, 20), rng.randint(-30, 20), rng.randint(-40, 5), rng.randint(-40, 5)]
    else:
        sc = [rng.randint(-1000, 1000) for _ in range(4)]
    rd, ad = _gen(rng, rng.choice([8, 30, 70, 130, 300]), 0, rng.choice([40, 200, 600]))
    a, b = abi_string(ol, rd, ad, sc), abi_string(rl, rd, ad, sc)
    if a != b and not (a.startswith('-1,') and b.startswith('-1,')):
        bad += 1
        print('MISMATCH', sc, repr(rd), repr(ad), a, b, flush=True)
        if bad > 5: break
print('seed', seed, 'iters', iters, 'bad', bad, 'sec %.0f' % (time.time() - t), flush=True)
