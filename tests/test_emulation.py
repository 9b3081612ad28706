"""CPU tier: the product's own kernel core (porechop_b200/csrc/dp_core.cuh, compiled for the host by
tests/emu/emu_group.cpp) against the golden fixtures and the oracle -- wavefront, trace packing, scout, traceback,
statistics, paired halves of different shapes, and the two-pass (score pass + bounded window) scheme."""
import random

from helpers import class_geometry, emu_slot, load_golden, oracle_record
from porechop_b200.align import record_string
from test_oracle import rebuild_fullread_inputs

SCHEMES = [[3, -6, -5, -2], [3, -6, -2, -2], [1, -1, -1, -1], [2, -3, -2, -5], [5, -4, -8, -1], [3, -6, -5, -5], [1, 0, -1, -1],
           [10, -20, -15, -7]]


def test_emu_golden_random_single_pass():
    cases = [c for c in load_golden('golden_random.json') if len(c[1]) <= 256]
    for k in range(0, len(cases) - 1, 2):
        (ra, aa, sca, ea), (rb, ab, scb, eb) = cases[k], cases[k + 1]
        G, R = class_geometry(max(len(aa), len(ab), 1))
        if sca == scb:
            st, recA, recB = emu_slot((ra, aa), (rb, ab), G, R, 0, sca)
            assert st == 0 and record_string(recA) == ea and record_string(recB) == eb
        else:
            for rd, ad, sc, exp in (cases[k], cases[k + 1]):
                st, rec, _ = emu_slot((rd, ad), None, G, R, 0, sc)
                assert st == 0 and record_string(rec) == exp


def test_emu_golden_windows_pairs():
    g = load_golden('golden_windows.json')
    reads = load_golden('fixture_reads.json')
    res = g['results']
    for k in range(0, len(res) - 1, 2):
        pair = []
        for ri, kind, ai, exp in (res[k], res[k + 1]):
            seq = reads[ri]['seq']
            pair.append(((seq[:150] if kind == 'start' else seq[-150:]), g['panel'][ai][1], exp))
        G, R = class_geometry(max(len(pair[0][1]), len(pair[1][1])))
        st, recA, recB = emu_slot(pair[0][:2], pair[1][:2], G, R, 0, g['scoring'])
        assert st == 0 and record_string(recA) == pair[0][2] and record_string(recB) == pair[1][2]


def test_emu_golden_fullread_two_pass():
    sc, cases = rebuild_fullread_inputs()
    for k in range(0, len(cases) - 1, 2):
        (sa, aa, ea), (sb, ab, eb) = cases[k], cases[k + 1]
        m = max(len(aa), len(ab))
        # score-pass geometry of the engine: R = 8
        G = 4 if m <= 32 else 8 if m <= 64 else 16 if m <= 128 else 32
        st, recA, recB = emu_slot((sa, aa), (sb, ab), G, 8, 1, sc)
        assert st == 0 and record_string(recA) == ea and record_string(recB) == eb


def _mut(rng, s, al):
    o = []
    for c in s:
        x = rng.random()
        if x < 0.04:
            continue
        if x < 0.09:
            o.append(rng.choice(al)); continue
        o.append(c)
        if x < 0.13:
            o.append(rng.choice(al))
    return ''.join(o)


def _gen(rng, mmax, nlo, nhi):
    al = rng.choice(['A', 'AC', 'ACGT', 'ACGTN', 'ACGT'])
    ad = ''.join(rng.choice(al) for _ in range(rng.randint(1, mmax)))
    n = rng.randint(nlo, nhi)
    rd = ''.join(rng.choice(al) for _ in range(n))
    if rng.random() < 0.7:
        for _ in range(rng.randint(1, 3)):
            p = rng.randint(0, len(rd))
            ins = _mut(rng, ad, al)
            if rng.random() < 0.3:
                ins = ins[rng.randint(0, len(ins)):]
            rd = rd[:p] + ins + rd[p:]
    if rng.random() < 0.03:
        rd = ''
    if rng.random() < 0.03:
        ad = ''
    return rd, ad


def test_emu_random_vs_oracle_all_geometries():
    rng = random.Random(4242)
    for it in range(2500):
        G, R = rng.choice([(8, 4), (16, 4), (32, 4), (32, 8), (4, 8), (8, 8), (16, 8), (4, 5), (4, 6), (4, 7), (8, 5), (16, 6), (32, 7)])
        mode = rng.choice([0, 0, 1])
        sc = rng.choice(SCHEMES)
        lo, hi = (50, 700) if mode else (0, 260)
        a = _gen(rng, G * R, lo, hi)
        b = _gen(rng, G * R, lo, hi) if rng.random() < 0.8 else None
        st, ra, rb = emu_slot(a, b, G, R, mode, sc)
        assert st == 0
        assert ra == oracle_record(a[0], a[1], sc), (G, R, mode, sc, a)
        if b:
            assert rb == oracle_record(b[0], b[1], sc), (G, R, mode, sc, b)


def test_emu_tight_window_bound_vs_oracle():
    """The per-alignment window bound (dp_core.cuh window_cols, option tight_window): never larger than the classic
    bound, and the windowed traceback still reproduces the oracle's records -- also for gappy hits whose paths span many
    read-only columns and for schemes with cheap gaps."""
    rng = random.Random(99)
    cheap = [[5, -4, -8, -1], [10, -20, -15, -7], [3, -6, -5, -2], [4, -1, -2, -1], [12, -3, -1, -1], [3, -6, -5, -5]]
    for it in range(1200):
        G, R = rng.choice([(4, 5), (4, 6), (4, 7), (4, 8), (8, 8), (16, 6), (32, 7)])
        sc = rng.choice(cheap)
        if max(abs(x) for x in sc) * (G * R + 3) > 4000:
            continue
        slots = []
        for _ in range(2):
            rd, ad = _gen(rng, G * R, 50, 600)
            if ad and rng.random() < 0.7:           # a gappy copy of the adapter: many read-only gap columns
                al = sorted(set(rd + ad)) or ['A']
                ins = []
                for c in ad:
                    ins.append(c)
                    while rng.random() < 0.35:
                        ins.append(rng.choice(al))
                p = rng.randint(0, len(rd))
                rd = rd[:p] + ''.join(ins) + rd[p:]
            slots.append((rd, ad))
        st, ra, rb = emu_slot(slots[0], slots[1], G, R, 5, sc)
        assert st == 0
        assert ra == oracle_record(slots[0][0], slots[0][1], sc), (G, R, sc, slots[0])
        assert rb == oracle_record(slots[1][0], slots[1][1], sc), (G, R, sc, slots[1])


def test_emu_query_profile_score_pass_vs_oracle():
    """Option "profile": the score pass takes its substitution operands from the query profile (dp_core.cuh profile_word,
    lane_step<.., PROF>) -- one read against two adapters per slot, classic and tight windows, all four group widths at
    R = 8; records equal the oracle's."""
    rng = random.Random(515)
    for it in range(700):
        G = rng.choice([4, 8, 16, 32])
        sc = rng.choice(SCHEMES)
        if max(abs(x) for x in sc) * (G * 8 + 3) > 4000 or not (sc[2] < 0 and sc[3] < 0):
            continue
        rd, adA = _gen(rng, G * 8, 50, 700)
        al = sorted(set(rd + adA)) or ['A']
        adB = ''.join(rng.choice(al) for _ in range(rng.randint(0 if rng.random() < 0.05 else 1, G * 8)))
        if adB and rng.random() < 0.6:
            p = rng.randint(0, len(rd))
            rd = rd[:p] + _mut(rng, adB, al) + rd[p:]
        mode = 1 | 8 | (4 if rng.random() < 0.5 else 0)
        st, ra, rb = emu_slot((rd, adA), (rd, adB), G, 8, mode, sc)
        assert st == 0
        assert ra == oracle_record(rd, adA, sc), (G, mode, sc, rd, adA)
        assert rb == oracle_record(rd, adB, sc), (G, mode, sc, rd, adB)


def test_emu_window_clips_real_adapter_lengths():
    """Two-pass with real adapter lengths (22-111) on multi-kb reads: the window really clips."""
    rng = random.Random(7)
    ads = ['AATGTACTTCGTTCAGTTACGTATTGCT', 'GCAATACGTAACTGAACGAAGT',
           'AATGTACTTCGTTCAGTTACGGCTTGGGTGTTTAACCAAGAAAGTTGTCGGTGTCTTTGTGGTTTTCGCATTTATCGTGAAACGCTTTCGCGTTTTTCGTGCGCCGCTTCA']
    for it in range(40):
        ad = rng.choice(ads)
        n = rng.randint(3000, 9000)
        rd = ''.join(rng.choice('ACGT') for _ in range(n))
        for _ in range(rng.randint(0, 2)):
            p = rng.randint(0, len(rd))
            rd = rd[:p] + _mut(rng, ad, 'ACGT') + rd[p:]
        if rng.random() < 0.3:
            p = rng.randint(0, len(rd) - 40)
            rd = rd[:p] + '-' * 30 + rd[p + 30:]
        m = len(ad)
        G = 4 if m <= 32 else 8 if m <= 64 else 16
        st, ra, _ = emu_slot((rd, ad), None, G, 8, 1, [3, -6, -5, -2])
        assert st == 0 and ra == oracle_record(rd, ad)
