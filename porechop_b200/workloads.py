"""
Synthetic workloads of BASELINE.json / SURVEY.md 8(d) (seed 20260923) and adapter tables.

Reads: length L = clamp(round(LogNormal(mu = ln 8000 - sigma^2/2, sigma = 0.5)), 200, 60000), bases iid
uniform ACGT, start adapter prepended with p = 0.8, end adapter appended with p = 0.5, every implanted copy
passed through an ONT-like error channel (3 % deletion, 4 % substitution, 3 % insertion per base) and a uniform
0-10 nt truncation of its outer end.  Implanted copies are drawn from a pool of pre-mutated variants per
adapter (POOL distinct copies) so that a million reads are generated with numpy only.

The adapter sequences come from tests/golden/adapters.json, a data fixture written by
tests/golden/make_golden.py from the reference's table (porechop/adapters.py:77-498).
"""
import json
import os
import struct

import numpy as np

SEED = 20260923
POOL = 4096
_HERE = os.path.dirname(os.path.abspath(__file__))
_ADAPTERS_JSON = os.path.join(os.path.dirname(_HERE), 'tests', 'golden', 'adapters.json')
DEFAULT_SCORING = (3, -6, -5, -2)    # porechop/porechop.py:145
END_SIZE = 150                       # porechop/porechop.py:130


def load_adapter_sets():
    with open(_ADAPTERS_JSON) as f:
        return json.load(f)


def nsk007():
    """(Y_Top, Y_Bottom) of the set named 'SQK-NSK007' (= the LSK109 Y-adapter)."""
    for s in load_adapter_sets()['sets']:
        if s['name'] == 'SQK-NSK007':
            return s['start'][1], s['end'][1]
    raise KeyError('SQK-NSK007')


def demux_adapters():
    """Config 3: all 119 table sets + 12 native-full + 96 rapid-full(new): (227 start seqs, 129 end seqs)."""
    d = load_adapter_sets()
    sets = list(d['sets'])
    full = d['full_barcode_sets']
    sets += [s for s in full if s['name'].startswith('Native barcoding')]
    sets += [s for s in full if s['name'].endswith('(full sequence, new)')]
    starts = [s['start'][1] for s in sets if s['start']]
    ends = [s['end'][1] for s in sets if s['end']]
    return starts, ends


def all_table_sequences():
    """Phase A: the 236 sequences of the 119 table sets, as (start list, end list)."""
    d = load_adapter_sets()
    return ([s['start'][1] for s in d['sets'] if s['start']], [s['end'][1] for s in d['sets'] if s['end']])


_ACGT = np.frombuffer(b'ACGT', dtype=np.uint8)


def _mutate(rng, seq):
    out = []
    for c in seq:
        x = rng.random()
        if x < 0.03:
            continue
        if x < 0.07:
            out.append(int(_ACGT[rng.integers(4)]))
            continue
        out.append(c)
        if x < 0.10:
            out.append(int(_ACGT[rng.integers(4)]))
    return bytes(out)


def _variant_pool(rng, adapter, truncate_front):
    """POOL mutated + outer-end-truncated copies as a padded matrix and their lengths."""
    a = adapter.encode()
    var = []
    for _ in range(POOL):
        v = _mutate(rng, a)
        t = int(rng.integers(0, 11))
        v = v[t:] if truncate_front else (v[:len(v) - t] if t else v)
        var.append(v)
    w = max(1, max(len(v) for v in var))
    mat = np.zeros((POOL, w), dtype=np.uint8)
    lens = np.zeros(POOL, dtype=np.int64)
    for k, v in enumerate(var):
        mat[k, :len(v)] = np.frombuffer(v, dtype=np.uint8) if v else []
        lens[k] = len(v)
    return mat, lens


def read_lengths(rng, n):
    sigma = 0.5
    L = np.rint(rng.lognormal(np.log(8000.0) - sigma * sigma / 2, sigma, n))
    return np.clip(L, 200, 60000).astype(np.int64)


def synth_end_windows(n_reads, start_adapter, end_adapter, seed=SEED, end_size=END_SIZE):
    """
    The two end windows of n_reads synthetic reads (the only bases an end-trim pass sends through the ABI):
    returns (lengths[n], start_windows uint8[n, end_size], end_windows uint8[n, end_size]).
    Every read is >= 2*end_size long by construction of the length law (min 200 < 300 is clamped up here and
    reported), so the two windows are independent.
    """
    rng = np.random.default_rng(seed)
    L = np.maximum(read_lengths(rng, n_reads), 2 * end_size)
    sw = _ACGT[rng.integers(0, 4, size=(n_reads, end_size), dtype=np.uint8)]
    ew = _ACGT[rng.integers(0, 4, size=(n_reads, end_size), dtype=np.uint8)]
    col = np.arange(end_size)[None, :]
    if start_adapter:
        mat, lens = _variant_pool(rng, start_adapter, truncate_front=True)
        has = rng.random(n_reads) < 0.8
        pick = rng.integers(0, POOL, n_reads)
        w = min(mat.shape[1], end_size)
        vl = np.where(has, np.minimum(lens[pick], w), 0)[:, None]
        src = mat[pick][:, :w]
        # variant occupies columns [0, vl); random read bases follow
        shifted = np.empty_like(sw)
        idx = np.clip(col - vl, 0, end_size - 1)
        shifted[:] = np.take_along_axis(sw, idx, axis=1)
        m = col < vl
        shifted[:, :w] = np.where(m[:, :w], src, shifted[:, :w])
        sw = shifted
    if end_adapter:
        mat, lens = _variant_pool(rng, end_adapter, truncate_front=False)
        has = rng.random(n_reads) < 0.5
        pick = rng.integers(0, POOL, n_reads)
        w = min(mat.shape[1], end_size)
        vl = np.where(has, np.minimum(lens[pick], w), 0)[:, None]
        # variant occupies the last vl columns: column c holds variant[c - (end_size - vl)]
        k = col - (end_size - vl)
        m = k >= 0
        src = np.take_along_axis(np.pad(mat[pick][:, :w], ((0, 0), (0, end_size - w))), np.clip(k, 0, end_size - 1), axis=1)
        ew = np.where(m, src, ew)
    return L, np.ascontiguousarray(sw), np.ascontiguousarray(ew)


def synth_reads(n_reads, start_adapter, end_adapter, seed=SEED, chimera_p=0.0, max_len=60000):
    """Full synthetic reads (for middle-adapter scans / the oracle harness): (uint8 buffer, int64 offsets)."""
    rng = np.random.default_rng(seed)
    L = np.minimum(read_lengths(rng, n_reads), max_len)
    smat, slen = _variant_pool(rng, start_adapter, True) if start_adapter else (None, None)
    emat, elen = _variant_pool(rng, end_adapter, False) if end_adapter else (None, None)
    parts = []
    for r in range(n_reads):
        body = _ACGT[rng.integers(0, 4, size=int(L[r]), dtype=np.uint8)]
        pieces = []
        if smat is not None and rng.random() < 0.8:
            k = int(rng.integers(POOL))
            pieces.append(smat[k, :slen[k]])
        if chimera_p and rng.random() < chimera_p and L[r] > 2200:
            pos = int(rng.integers(1000, int(L[r]) - 1000))
            ke, ks = int(rng.integers(POOL)), int(rng.integers(POOL))
            pieces += [body[:pos], emat[ke, :elen[ke]], smat[ks, :slen[ks]], body[pos:]]
        else:
            pieces.append(body)
        if emat is not None and rng.random() < 0.5:
            k = int(rng.integers(POOL))
            pieces.append(emat[k, :elen[k]])
        parts.append(np.concatenate(pieces))
    off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum([len(p) for p in parts], out=off[1:])
    return np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0, np.uint8), off


def _scatter(buf, starts, mat, pick, lens):
    """buf[starts[r] : starts[r] + lens[r]] = mat[pick[r], :lens[r]] for every r with lens[r] > 0 (vectorised)."""
    rows = np.nonzero(lens > 0)[0]
    if not len(rows):
        return
    l = lens[rows].astype(np.int64)
    tot = int(l.sum())
    first = np.cumsum(l) - l
    within = np.arange(tot, dtype=np.int64) - np.repeat(first, l)
    buf[np.repeat(starts[rows].astype(np.int64), l) + within] = mat[np.repeat(pick[rows], l), within]


def synth_reads_fast(n_reads, start_adapter, end_adapter, seed=SEED, chimera_p=0.0, max_len=60000, out=None):
    """Full synthetic reads of SURVEY 8(d) without a per-read Python loop (10^6+ reads): same length law, implant
    probabilities, error channel (variant pools) and chimera rule as synth_reads.  Every per-read attribute has its own
    generator keyed by (seed, attribute) and the bodies are pbioRandomBases(seed), so the first k reads (and their bytes)
    are the same for every n_reads >= k -- the CPU reference arm times a prefix of exactly the batch the GPU arm aligns.
    `out`: optional uint8 buffer (e.g. pinned memory) the bases are written into.  Returns (buffer, int64 offsets)."""
    from . import hostio
    g = lambda k: np.random.default_rng([int(seed), k])
    n = int(n_reads)
    Lb = np.minimum(read_lengths(g(0), n), max_len)
    zero = np.zeros(n, dtype=np.int64)
    if start_adapter:
        smat, slen = _variant_pool(g(1), start_adapter, True)
        pick_s = g(4).integers(0, POOL, n)
        ls = np.where(g(3).random(n) < 0.8, slen[pick_s], 0)
    else:
        smat, slen, pick_s, ls = None, None, None, zero
    if end_adapter:
        emat, elen = _variant_pool(g(2), end_adapter, False)
        pick_e = g(6).integers(0, POOL, n)
        le = np.where(g(5).random(n) < 0.5, elen[pick_e], 0)
    else:
        emat, elen, pick_e, le = None, None, None, zero
    lce = lcs = zero
    if chimera_p and smat is not None and emat is not None:
        chim = (g(7).random(n) < chimera_p) & (Lb > 2200)
        pos = 1000 + np.floor(g(8).random(n) * np.maximum(Lb - 2000, 1)).astype(np.int64)
        ke, ks = g(9).integers(0, POOL, n), g(10).integers(0, POOL, n)
        lce, lcs = np.where(chim, elen[ke], 0), np.where(chim, slen[ks], 0)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(Lb + ls + le + lce + lcs, out=off[1:])
    buf = hostio.random_bases(int(off[-1]), seed, out)
    if smat is not None:
        _scatter(buf, off[:-1], smat, pick_s, ls)
    if lce is not zero:
        _scatter(buf, off[:-1] + ls + pos, emat, ke, lce)
        _scatter(buf, off[:-1] + ls + pos + lce, smat, ks, lcs)
    if emat is not None:
        _scatter(buf, off[1:] - le, emat, pick_e, le)
    return buf, off


def forward_barcode_sequences():
    """Config 5: the 192 sequences (24 nt each) of the 96 forward barcode sets, start sequences then end sequences."""
    sets = [s for s in load_adapter_sets()['sets'] if s['name'].endswith('(forward)')]
    return [s['start'][1] for s in sets] + [s['end'][1] for s in sets]


def synth_fixed_length_reads(n_reads, length, barcodes, seed=SEED, out=None):
    """BASELINE config 5 (read-length sweep): n_reads reads of exactly `length` bases, iid ACGT, each with one randomly
    chosen barcode sequence written over a random position (so every read has a real hit somewhere).  Vectorised;
    prefix-consistent in n_reads like synth_reads_fast."""
    from . import hostio
    n, length = int(n_reads), int(length)
    buf = hostio.random_bases(n * length, int(seed), out)
    w = max(len(b) for b in barcodes)
    mat = np.zeros((len(barcodes), w), dtype=np.uint8)
    blen = np.zeros(len(barcodes), dtype=np.int64)
    for k, b in enumerate(barcodes):
        mat[k, :len(b)] = np.frombuffer(b.encode(), dtype=np.uint8)
        blen[k] = len(b)
    which = np.random.default_rng([int(seed), 1]).integers(0, len(barcodes), n)
    l = np.where(blen[which] <= length, blen[which], 0)
    p = np.floor(np.random.default_rng([int(seed), 2]).random(n) * (length - l + 1)).astype(np.int64)
    _scatter(buf, np.arange(n, dtype=np.int64) * length + p, mat, which, l)
    return buf, np.arange(n + 1, dtype=np.int64) * length


def windows_to_batch(windows):
    """uint8[n, w] -> (flat buffer, int64 offsets) for the C-ABI."""
    n, w = windows.shape
    return windows.reshape(-1), (np.arange(n + 1, dtype=np.int64) * w)


def pack_adapters(seqs):
    bs = [s.encode() for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int32)
    np.cumsum([len(b) for b in bs], out=off[1:])
    return np.frombuffer(b''.join(bs), dtype=np.uint8).copy(), off


def write_harness_file(path, seq_buf, seq_off, ad_buf, ad_off, scoring, pair_seq=None, pair_adapter=None):
    """Workload file for oracle/ref_harness.cpp (format documented there)."""
    n_seqs, n_ad = len(seq_off) - 1, len(ad_off) - 1
    cross = pair_seq is None
    n_pairs = n_seqs * n_ad if cross else len(pair_seq)
    with open(path, 'wb') as f:
        f.write(struct.pack('<qqq', n_seqs, n_ad, n_pairs))
        f.write(struct.pack('<iiiiii', int(scoring[0]), int(scoring[1]), int(scoring[2]), int(scoring[3]), 1 if cross else 0, 0))
        f.write(np.ascontiguousarray(seq_off, dtype=np.int64).tobytes())
        f.write(np.ascontiguousarray(seq_buf, dtype=np.uint8).tobytes())
        f.write(np.ascontiguousarray(ad_off, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(ad_buf, dtype=np.uint8).tobytes())
        if not cross:
            f.write(np.ascontiguousarray(pair_seq, dtype=np.int32).tobytes())
            f.write(np.ascontiguousarray(pair_adapter, dtype=np.int32).tobytes())
