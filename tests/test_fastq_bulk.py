"""Bulk FASTQ ingest + vectorised end-trim rule (porechop_b200/fastq.py).  CPU tier runs it with the oracle standing in
for the engine; the GPU tier (marked) runs the same through the real engine.  Expected values are the reference's own
Phase B results (tests/golden/golden_phases.json)."""
import numpy as np
import pytest

from helpers import load_golden, oracle_batch

SC = [3, -6, -5, -2]


def fastq_text(reads, crlf=False):
    nl = '\r\n' if crlf else '\n'
    return ''.join('@%s%s%s%s+%s%s%s' % (r['name'], nl, r['seq'], nl, nl, 'I' * len(r['seq']), nl) for r in reads).encode()


def fixture(file_name):
    return [r for r in load_golden('fixture_reads.json') if r['file'] == file_name]


@pytest.mark.parametrize('crlf', [False, True])
def test_parse_fastq_roundtrip(crlf):
    from porechop_b200.fastq import parse_fastq
    reads = fixture('test_one_adapter_set.fastq') + [{'name': 'tiny', 'seq': 'ACGTN'}, {'name': 'empty', 'seq': ''}]
    b = parse_fastq(fastq_text(reads, crlf))
    assert len(b) == len(reads) and b.names == [r['name'] for r in reads]
    for i, r in enumerate(reads):
        assert bytes(b.seq[b.seq_off[i]:b.seq_off[i + 1]]).decode() == r['seq']
        assert b.qual_off[i + 1] - b.qual_off[i] == len(r['seq'])
    with pytest.raises(ValueError):
        parse_fastq(b'@x\nACGT\n+\n')


def test_end_windows_match_python_slices():
    from porechop_b200.fastq import end_windows, parse_fastq
    reads = fixture('test_barcodes.fastq') + [{'name': 's1', 'seq': 'ACGTACGTAC'}, {'name': 's2', 'seq': 'A' * 150}, {'name': 's3', 'seq': ''}]
    b = parse_fastq(fastq_text(reads))
    (sb, so), (eb, eo) = end_windows(b.seq, b.seq_off, 150)
    for i, r in enumerate(reads):
        assert bytes(sb[so[i]:so[i + 1]]).decode() == r['seq'][:150]
        assert bytes(eb[eo[i]:eo[i + 1]]).decode() == (r['seq'][-150:] if r['seq'] else '')


def _check_case(case_index):
    from porechop_b200.fastq import parse_fastq, trim_end_adapters
    case = load_golden('golden_phases.json')[case_index]
    ad = load_golden('adapters.json')
    sets = {d['name']: d for d in ad['sets'] + ad['full_barcode_sets']}
    matching = [sets[n] for n in case['matching_sets']]
    starts = [m['start'][1] for m in matching if m['start']]
    ends = [m['end'][1] for m in matching if m['end']]
    b = parse_fastq(fastq_text(fixture(case['file'])))
    st, et, _, _ = trim_end_adapters(b, starts, ends, SC)
    assert list(st) == [r['start_trim_amount'] for r in case['reads']]
    assert list(et) == [r['end_trim_amount'] for r in case['reads']]


def _check_search(case_index):
    from porechop_b200.fastq import parse_fastq, search_adapter_sets
    case = load_golden('golden_phases.json')[case_index]
    ad = load_golden('adapters.json')
    sets = [(d['name'], d['start'] or None, d['end'] or None) for d in ad['sets']]
    b = parse_fastq(fastq_text(fixture(case['file'])))
    bs, be = search_adapter_sets(b, sets, SC)
    want = {nm: (s, e) for nm, s, e in case['set_scores']}
    assert len(want) == len(sets)
    for (nm, _, _), s, e in zip(sets, bs, be):
        assert (s, e) == want[nm], nm


@pytest.mark.parametrize('case_index', [0, 3])
def test_search_adapter_sets_with_oracle_engine(monkeypatch, case_index):
    """Phase A on flat buffers: best_start_score / best_end_score of all 119 table sets = the reference's."""
    from porechop_b200 import fastq

    def fake(seq_buf, seq_off, ad_buf, ad_off, scoring, pair_seq=None, pair_adapter=None, out=None):
        return oracle_batch(seq_buf, seq_off, ad_buf, ad_off, scoring, pair_seq, pair_adapter)
    monkeypatch.setattr(fastq.W, 'adapter_alignment_batch', fake)
    _check_search(case_index)


@pytest.mark.gpu
@pytest.mark.parametrize('case_index', [0, 3])
def test_search_adapter_sets_gpu(case_index):
    _check_search(case_index)


@pytest.mark.parametrize('case_index', [0, 2, 3])
def test_bulk_end_trim_with_oracle_engine(monkeypatch, case_index):
    from porechop_b200 import fastq

    def fake(seq_buf, seq_off, ad_buf, ad_off, scoring, pair_seq=None, pair_adapter=None, out=None):
        return oracle_batch(seq_buf, seq_off, ad_buf, ad_off, scoring, pair_seq, pair_adapter)
    monkeypatch.setattr(fastq.W, 'adapter_alignment_batch', fake)
    _check_case(case_index)


@pytest.mark.gpu
@pytest.mark.parametrize('case_index', [0, 2, 3])
def test_bulk_end_trim_gpu(case_index):
    _check_case(case_index)


def _messy_fastq(seed, n):
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n):
        L = int(rng.integers(0, 400))
        alphabet = 'ACGTacgtUuNn' if i % 5 == 0 else 'ACGT'
        seq = ''.join(rng.choice(list(alphabet), L))
        ql = L if i % 7 else max(0, L - int(rng.integers(0, 20)))
        qual = ''.join(chr(c) for c in rng.integers(33, 74, ql))
        nl = '\r\n' if i % 11 == 0 else '\n'
        pad = ' \t' if i % 13 == 0 else ''
        recs.append('@r%d%s%s%s%s%s%s+%s%s%s' % (i, ' desc x' if i % 3 == 0 else '', pad, nl, seq, pad, nl, nl, qual, nl))
    text = ''.join(recs)
    return text[:-1] if seed % 2 else text            # odd seeds: no newline at the end of the file


@pytest.mark.parametrize('seed', [1, 2, 3])
def test_native_hostio_equals_numpy_implementation(monkeypatch, seed):
    """libhostio.so (C) and the numpy implementations of parse / window gather / emit give identical buffers on messy
    input: CRLF, blanks, lower case, RNA reads, short qualities, empty reads, missing final newline."""
    from porechop_b200 import fastq, hostio
    assert hostio.LIB is not None, 'libhostio.so not built (python -m porechop_b200.build)'
    data = _messy_fastq(seed, 300).encode()
    native = fastq.parse_fastq(data)
    rng = np.random.default_rng(seed)
    st, et = rng.integers(0, 60, len(native)) * (rng.random(len(native)) < 0.6), rng.integers(0, 160, len(native)) * (rng.random(len(native)) < 0.6)
    middle = {int(i): [(int(a), int(a) + 30)] for i, a in zip(rng.choice(len(native), 40, replace=False), rng.integers(-10, 200, 40))}
    outs = [fastq.emit(native, st, et, middle, fmt, 20) for fmt in ('fastq', 'fasta')]
    wins = fastq.end_windows(native.seq, native.seq_off, 150)
    arr = fastq.emit(native, st, et, middle, 'fastq', 20, as_array=True)
    assert isinstance(arr, np.ndarray) and arr.tobytes() == outs[0]
    monkeypatch.setattr(hostio, 'LIB', None)
    slow = fastq.parse_fastq(data)
    assert slow.names == native.names and (slow.rna == native.rna).all() and native.rna.any()
    for k in ('seq', 'seq_off', 'qual', 'qual_off', 'name_buf', 'name_off'):
        assert np.array_equal(getattr(slow, k), getattr(native, k)), k
    assert [fastq.emit(slow, st, et, middle, fmt, 20) for fmt in ('fastq', 'fasta')] == outs
    for (a, ao), (b, bo) in zip(wins, fastq.end_windows(slow.seq, slow.seq_off, 150)):
        assert np.array_equal(a, b) and np.array_equal(ao, bo)
    for bad in (b'@x\nACGT\n+\n', b'xx\nAC\n+\nII\n'):
        for lib in (None, hostio._load()):
            monkeypatch.setattr(hostio, 'LIB', lib)
            with pytest.raises(ValueError):
                fastq.parse_fastq(bad)


def test_parse_fasta_multiline_native_and_numpy(monkeypatch):
    """multi-line FASTA with blank lines, CRLF, padding, lower case and an RNA record; both implementations agree and
    match a plain Python restatement of the loader."""
    from porechop_b200 import fastq, hostio
    text = ('>r1 first read\nACGTacgt\nGGGG\n\n>r2\r\n  TTTTAAAA  \r\nCC\r\n>rna one\nUUUUACGU\n>empty\n>last\nACGT')
    want = [('r1 first read', 'ACGTACGTGGGG'), ('r2', 'TTTTAAAACC'), ('rna one', 'TTTTACGT'), ('empty', ''), ('last', 'ACGT')]
    for lib in (hostio._load(), None):
        monkeypatch.setattr(hostio, 'LIB', lib)
        b = fastq.parse_fasta(text.encode())
        got = [(nm, bytes(b.seq[b.seq_off[i]:b.seq_off[i + 1]]).decode()) for i, nm in enumerate(b.names)]
        assert got == want and list(b.rna) == [False, False, True, False, False]
        assert bytes(b.qual) == b'+' * len(b.seq) and np.array_equal(b.qual_off, b.seq_off)
        assert fastq.emit(b, fmt='fasta') == b'>r1 first read\nACGTACGTGGGG\n>r2\nTTTTAAAACC\n>rna one\nUUUUACGU\n>last\nACGT\n'
        assert fastq.parse_reads(text.encode())[1] == 'fasta' and fastq.parse_reads(b'@a\nA\n+\nI\n')[1] == 'fastq'
        assert len(fastq.parse_fasta(b'')) == 0
        for bad in (b'ACGT\n>x\nAC\n', b'>\nACGT\n'):
            with pytest.raises(ValueError):
                fastq.parse_fasta(bad)
        with pytest.raises(ValueError):
            fastq.parse_reads(b'xyz')


def test_end_trim_rule_native_equals_numpy(monkeypatch):
    """pbioEndTrim / pbioFullScores against the numpy rule on random records that hit every branch: threshold ties,
    read_end == end_size, read_start == 0, short alignments, 0/0 (NaN) and failed alignments."""
    from porechop_b200 import fastq, hostio
    assert hostio.LIB is not None
    rng = np.random.default_rng(9)
    n, a = 3000, 7
    r = np.zeros((n, a, 9), dtype=np.int32)
    r[..., 0] = rng.integers(0, 150, (n, a)) * (rng.random((n, a)) < 0.8)
    r[..., 1] = np.minimum(r[..., 0] + rng.integers(0, 40, (n, a)), 149)
    r[..., 6] = rng.integers(0, 40, (n, a))
    r[..., 5] = np.minimum((r[..., 6] * rng.choice([0.5, 0.75, 0.8, 1.0], (n, a))).astype(np.int32), r[..., 6])
    r[..., 8] = rng.integers(1, 30, (n, a))
    r[..., 7] = (r[..., 8] * rng.random((n, a))).astype(np.int32)
    fail = rng.random((n, a)) < 0.05
    r[fail] = [-1, 0, -1, 0, -2147483648, 0, 0, 0, 0]
    for thr, mts, extra in [(75.0, 4, 2), (80.0, 1, 0), (50.0, 10, 5)]:
        native = fastq.end_trim_amounts(r, r[:, ::-1], 150, extra, thr, mts)
        cols = [5, 0, 3]
        nf = hostio.full_scores(r, cols)
        monkeypatch.setattr(hostio, 'LIB', None)
        slow = fastq.end_trim_amounts(r, r[:, ::-1], 150, extra, thr, mts)
        sf = fastq.scores_from_records(r[:, cols, :].reshape(-1, 9))[0].reshape(n, len(cols))
        monkeypatch.undo()
        assert np.array_equal(native[0], slow[0]) and np.array_equal(native[1], slow[1]) and native[0].max() > 0
        assert np.array_equal(nf, sf, equal_nan=True)


def test_parallel_gzip_members_roundtrip():
    """pbioGzip: block-parallel multi-member .gz; gunzip gives the input back for sizes around the block boundaries."""
    import gzip
    from porechop_b200 import hostio
    assert hostio.LIB is not None
    rng = np.random.default_rng(2)
    for n, block in [(0, 4096), (1, 4096), (4095, 4096), (4096, 4096), (4097, 4096), (100000, 4096), (300000, 1 << 20)]:
        data = rng.choice(np.frombuffer(b'ACGT5+@\n', dtype=np.uint8), n).tobytes()
        z = hostio.gzip_members(data, 6, block)
        assert z is not None and gzip.decompress(bytes(z)) == data
        assert bytes(z).count(b'\x1f\x8b\x08') >= (n + block - 1) // block
