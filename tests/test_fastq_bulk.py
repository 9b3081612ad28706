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
