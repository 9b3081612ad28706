"""GPU tier (pytest -m gpu): the CUDA engine, called through the C-ABI, against the committed golden fixtures
(reference outputs) and the oracle on seeded inputs -- bit-exact 9-int records / result strings."""
import random

import numpy as np
import pytest

from helpers import DEFAULT, load_golden, oracle_batch
from test_oracle import SURVEY_EDGE, rebuild_fullread_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def W():
    from porechop_b200 import cpp_function_wrappers as w
    assert w.device_count() > 0
    return w


def run_pairs(W, reads, adapters, pairs, sc):
    """pairs: list of (read index, adapter index) -> records via the pair-list ABI."""
    sbuf, soff = W.pack_sequences(reads)
    abuf, aoff = W.pack_sequences(adapters, offset_dtype=np.int32)
    ps = np.array([p[0] for p in pairs], dtype=np.int32)
    pa = np.array([p[1] for p in pairs], dtype=np.int32)
    return W.adapter_alignment_batch(sbuf, soff, abuf, aoff, sc, ps, pa)


def test_legacy_single_call_strings(W):
    for rd, ad, exp in SURVEY_EDGE:
        assert W.adapter_alignment(rd, ad, DEFAULT) == exp
    g = load_golden('golden_random.json')
    for rd, ad, sc, exp in g[:300]:
        assert W.adapter_alignment(rd, ad, sc) == exp


def test_legacy_call_is_thread_safe(W):
    """The reference calls adapterAlignment from a multiprocessing.dummy.Pool (porechop.py:312,496,579)."""
    from multiprocessing.dummy import Pool
    g = load_golden('golden_random.json')[300:700]
    with Pool(8) as pool:
        got = pool.map(lambda c: W.adapter_alignment(c[0], c[1], c[2]), g)
    assert got == [c[3] for c in g]


def test_golden_random_pair_list(W):
    g = load_golden('golden_random.json')
    by_scheme = {}
    for rd, ad, sc, exp in g:
        by_scheme.setdefault(tuple(sc), []).append((rd, ad, exp))
    for sc, cases in by_scheme.items():
        reads = [c[0] for c in cases]
        ads = [c[1] for c in cases]
        recs = run_pairs(W, reads, ads, [(k, k) for k in range(len(cases))], list(sc))
        for k, c in enumerate(cases):
            assert W.format_record(recs[k]) == c[2], (sc, c[0], c[1])


def test_golden_windows_cross(W):
    g = load_golden('golden_windows.json')
    reads = load_golden('fixture_reads.json')
    wins = []
    for r in reads:
        wins += [r['seq'][:150], r['seq'][-150:]]
    ads = [p[1] for p in g['panel']]
    sbuf, soff = W.pack_sequences(wins)
    abuf, aoff = W.pack_sequences(ads, offset_dtype=np.int32)
    recs = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, g['scoring']).reshape(len(wins), len(ads), 9)
    for ri, kind, ai, exp in g['results']:
        w = 2 * ri + (0 if kind == 'start' else 1)
        assert W.format_record(recs[w, ai]) == exp


def test_golden_fullread_two_pass(W):
    sc, cases = rebuild_fullread_inputs()
    reads = [c[0] for c in cases]
    ads = sorted(set(c[1] for c in cases))
    recs = run_pairs(W, reads, ads, [(k, ads.index(c[1])) for k, c in enumerate(cases)], sc)
    for k, c in enumerate(cases):
        assert W.format_record(recs[k]) == c[2]
    # the same reads as a cross product (same read, several adapters per slot)
    base = [r['seq'] for r in load_golden('fixture_reads.json')]
    sbuf, soff = W.pack_sequences(base)
    abuf, aoff = W.pack_sequences(ads, offset_dtype=np.int32)
    got = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, sc)
    assert np.array_equal(got, oracle_batch(sbuf, soff, abuf, aoff, sc))


def test_synthetic_windows_vs_oracle(W):
    from porechop_b200 import workloads as wl
    yt, yb = wl.nsk007()
    _, sw, ew = wl.synth_end_windows(20000, yt, yb)
    for win, ad in ((sw, yt), (ew, yb)):
        sbuf, soff = wl.windows_to_batch(win)
        abuf, aoff = wl.pack_adapters([ad])
        got = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING)
        assert np.array_equal(got, oracle_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING))


def test_demux_cross_all_adapters_vs_oracle(W):
    from porechop_b200 import workloads as wl
    starts, ends = wl.demux_adapters()
    _, sw, ew = wl.synth_end_windows(300, starts[5], ends[5], seed=5)
    for win, ads in ((sw, starts), (ew, ends)):
        sbuf, soff = wl.windows_to_batch(win)
        abuf, aoff = wl.pack_adapters(ads)
        got = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING)
        assert np.array_equal(got, oracle_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING))


def test_ragged_and_edge_inputs(W):
    rng = random.Random(3)
    yt = 'AATGTACTTCGTTCAGTTACGTATTGCT'
    reads = ['', 'A', 'N' * 20, '-' * 20, 'acgu' * 10, yt, yt[5:], 'GG' + yt + 'GG', 'ACGT' * 300]
    reads += [''.join(rng.choice('ACGTN') for _ in range(rng.randint(1, 400))) for _ in range(200)]
    ads = ['', 'A', yt, 'GCAATACGTAACTGAACGAAGT', 'ACGT' * 10, 'N' * 5, 'ACGT' * 30, 'ACGT' * 60]
    sbuf, soff = W.pack_sequences(reads)
    abuf, aoff = W.pack_sequences(ads, offset_dtype=np.int32)
    for sc in ([3, -6, -5, -2], [3, -6, -5, -5], [2, -3, -2, -5]):
        got = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, sc)
        assert np.array_equal(got, oracle_batch(sbuf, soff, abuf, aoff, sc)), sc


def test_long_reads_two_pass_vs_oracle(W):
    from porechop_b200 import workloads as wl
    yt, yb = wl.nsk007()
    buf, off = wl.synth_reads(120, yt, yb, seed=11, chimera_p=0.3, max_len=20000)
    full = wl.demux_adapters()[0][-1]      # a 111-nt full rapid barcode adapter
    abuf, aoff = wl.pack_adapters([yt, yb, full])
    got = W.adapter_alignment_batch(buf, off, abuf, aoff, wl.DEFAULT_SCORING)
    assert np.array_equal(got, oracle_batch(buf, off, abuf, aoff, wl.DEFAULT_SCORING))
    # masked re-alignment round (Phase C semantics): mask every hit with '-' and align again as a pair list
    recs = got.reshape(120, 3, 9)
    masked = buf.copy()
    for r in range(120):
        rs, re_ = recs[r, 0, 0], recs[r, 0, 1] + 1
        if rs >= 0 and recs[r, 0, 7] * 100 >= 70 * recs[r, 0, 8]:
            masked[off[r] + rs: off[r] + re_] = ord('-')
    ps = np.repeat(np.arange(120, dtype=np.int32), 2)
    pa = np.tile(np.array([0, 1], dtype=np.int32), 120)
    got2 = W.adapter_alignment_batch(masked, off, abuf, aoff, wl.DEFAULT_SCORING, ps, pa)
    assert np.array_equal(got2, oracle_batch(masked, off, abuf, aoff, wl.DEFAULT_SCORING, ps, pa))


def test_read_length_sweep_barcodes(W):
    """BASELINE config 5 shape: read lengths 500 bp - 100 kb x forward barcode sequences (24 nt), full-read scan."""
    from porechop_b200 import workloads as wl
    rng = np.random.default_rng(5)
    sets = [s for s in wl.load_adapter_sets()['sets'] if s['name'].endswith('(forward)')][:6]
    ads = [s['start'][1] for s in sets] + [s['end'][1] for s in sets]
    reads = []
    for L in (500, 1000, 2000, 5000, 10000, 20000, 50000, 100000):
        body = bytes(np.frombuffer(b'ACGT', dtype=np.uint8)[rng.integers(0, 4, L)]).decode()
        p = int(rng.integers(0, L - 30))
        reads.append(body[:p] + ads[int(rng.integers(len(ads)))] + body[p + 24:])
    sbuf, soff = W.pack_sequences(reads)
    abuf, aoff = wl.pack_adapters(ads)
    got = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING)
    assert np.array_equal(got, oracle_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING))


def test_score_pass_with_and_without_query_profile(W):
    """The score pass has two substitution paths (query profile = default for same-read slots, computed operands = option
    profile=0, odd classes, pair lists) and two window bounds (per-alignment = default, per-adapter): same records."""
    from porechop_b200 import workloads as wl
    yt, yb = wl.nsk007()
    buf, off = wl.synth_reads(60, yt, yb, seed=31, chimera_p=0.3, max_len=9000)
    abuf, aoff = wl.pack_adapters([yt, yb])
    exp = oracle_batch(buf, off, abuf, aoff, wl.DEFAULT_SCORING)
    got0 = W.adapter_alignment_batch(buf, off, abuf, aoff, wl.DEFAULT_SCORING)
    try:
        W.set_option('profile', 0)
        got1 = W.adapter_alignment_batch(buf, off, abuf, aoff, wl.DEFAULT_SCORING)
        W.set_option('tight_window', 0)
        got2 = W.adapter_alignment_batch(buf, off, abuf, aoff, wl.DEFAULT_SCORING)
        # a large mismatch penalty (still inside the int16 domain)
        sc = [3, -120, -5, -2]
        got3 = W.adapter_alignment_batch(buf, off, abuf, aoff, sc)
    finally:
        W.set_option('profile', 1)
        W.set_option('tight_window', 1)
    assert np.array_equal(got0, exp) and np.array_equal(got1, exp) and np.array_equal(got2, exp)
    assert np.array_equal(got3, oracle_batch(buf, off, abuf, aoff, sc))


def test_generic_int32_path(W):
    """Schemes / adapters outside the int16 domain take the int32 kernel: positive gap score, huge scores, m > 256."""
    rng = random.Random(8)
    reads = [''.join(rng.choice('ACGT') for _ in range(rng.randint(1, 300))) for _ in range(40)]
    ads = [''.join(rng.choice('ACGT') for _ in range(m)) for m in (5, 28, 300)]
    sbuf, soff = W.pack_sequences(reads)
    abuf, aoff = W.pack_sequences(ads, offset_dtype=np.int32)
    for sc in ([3, -6, 1, -2], [3000, -6000, -5000, -2000], [3, -6, -5, -2], [1, 2, -1, -1]):
        got = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, sc)
        assert np.array_equal(got, oracle_batch(sbuf, soff, abuf, aoff, sc)), sc


def test_single_pass_long_windows_and_global_staging(W):
    """One-pass (no score pass) run over multi-kb sequences with the packed bases staged in global scratch."""
    from porechop_b200 import workloads as wl
    yt, yb = wl.nsk007()
    buf, off = wl.synth_reads(40, yt, yb, seed=21, max_len=3000)
    abuf, aoff = wl.pack_adapters([yt, yb])
    exp = oracle_batch(buf, off, abuf, aoff, wl.DEFAULT_SCORING)
    try:
        W.set_option('hbuf', 'global')
        W.set_option('direct_max', 100000)
        got = W.adapter_alignment_batch(buf, off, abuf, aoff, wl.DEFAULT_SCORING)
    finally:
        W.set_option('hbuf', 'auto')
        W.set_option('direct_max', 160)
    assert np.array_equal(got, exp)


def test_scratch_cap_option_does_not_change_results(W):
    from porechop_b200 import workloads as wl
    yt, yb = wl.nsk007()
    _, sw, _ = wl.synth_end_windows(300000, yt, yb, seed=123)
    sbuf, soff = wl.windows_to_batch(sw)
    abuf, aoff = wl.pack_adapters([yt])
    a = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING)
    try:
        W.set_option('scratch_mb', 72)
        W.set_option('chunk_tasks', 50000)
        b = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING)
    finally:
        W.set_option('scratch_mb', 128)
        W.set_option('chunk_tasks', 131072)
    assert np.array_equal(a, b)


def test_device_resident_api_equals_host_api(W):
    import torch
    from porechop_b200 import workloads as wl
    yt, yb = wl.nsk007()
    _, sw, _ = wl.synth_end_windows(50000, yt, yb, seed=77)
    sbuf, soff = wl.windows_to_batch(sw)
    abuf, aoff = wl.pack_adapters([yt, yb])
    host = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING)
    d_seq = torch.from_numpy(sbuf).cuda()
    d_off = torch.from_numpy(soff).cuda()
    d_out = torch.empty((len(soff) - 1) * 2 * 9, dtype=torch.int32, device='cuda')
    torch.cuda.synchronize()
    W.adapter_alignment_batch_device(d_seq.data_ptr(), d_off.data_ptr(), len(soff) - 1, d_seq.numel(), 150, abuf, aoff,
                                     wl.DEFAULT_SCORING, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    W.synchronize()
    assert np.array_equal(d_out.cpu().numpy().reshape(-1, 9), host)


def test_full_size_properties(W):
    """BASELINE config-1 size (1M reads): properties that need no oracle -- the run is deterministic, implanted
    untruncated adapters are found at read start with the reference's coordinates, and a 1 % sample equals the oracle."""
    from porechop_b200 import workloads as wl
    yt, yb = wl.nsk007()
    n = 1000000
    _, sw, _ = wl.synth_end_windows(n, yt, yb)
    sbuf, soff = wl.windows_to_batch(sw)
    abuf, aoff = wl.pack_adapters([yt])
    a = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING)
    b = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING)
    assert np.array_equal(a, b)
    assert (a[:, 4] >= 0).all() and (a[:, 6] >= 0).all() and (a[:, 5] <= np.maximum(a[:, 6], 0)).all()
    assert (a[:, 3] <= 27).all() and (a[:, 1] <= 149).all()
    strong = a[:, 7] * 100 >= 60 * a[:, 8]
    assert 0.7 < strong.mean() < 0.9          # ~80 % of reads carry a (mutated, truncated) start adapter
    idx = np.arange(0, n, 100)
    sub = sw[idx]
    s2, o2 = wl.windows_to_batch(sub)
    assert np.array_equal(a[idx], oracle_batch(s2, o2, abuf, aoff, wl.DEFAULT_SCORING))
