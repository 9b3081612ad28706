"""Flat FASTQ pipeline (porechop_b200/fastq.py::trim_fastq) against whole-CLI outputs of the unmodified reference
(tests/golden/golden_emit.json, made by tests/golden/make_golden_emit.py): FASTQ bytes in -> the exact bytes
`porechop -i in.fastq -o out.fastq|fasta` wrote, for five option sets.  CPU tier: the oracle stands in for the engine
(tests only); GPU tier (marked): the real engine."""
import numpy as np
import pytest

from helpers import load_golden, oracle_batch

CASES = ['default', 'small_parts', 'discard_middle', 'no_split_fasta', 'fasta_split']


def _run(case_name):
    from porechop_b200.fastq import trim_fastq
    g = load_golden('golden_emit.json')
    c = g['cases'][case_name]
    sets = [(tuple(s) if s else None, tuple(e) if e else None) for s, e in c['matching_sets']]
    out, info = trim_fastq(g['input_fastq'].encode(), sets, c['scoring'], **c['options'])
    assert out.decode() == c['output']
    return info


def _oracle_engine(monkeypatch):
    from porechop_b200 import fastq

    def fake(seq_buf, seq_off, ad_buf, ad_off, scoring, pair_seq=None, pair_adapter=None, out=None):
        return oracle_batch(np.asarray(seq_buf), np.asarray(seq_off), np.asarray(ad_buf), np.asarray(ad_off), list(scoring),
                            pair_seq, pair_adapter)
    monkeypatch.setattr(fastq.W, 'adapter_alignment_batch', fake)


@pytest.mark.parametrize('case_name', CASES)
def test_trim_fastq_matches_reference_cli_oracle_engine(monkeypatch, case_name):
    _oracle_engine(monkeypatch)
    info = _run(case_name)
    if case_name == 'default':
        assert info['n_reads'] == 26 and len(info['middle']) >= 4       # the chimeras were found


@pytest.mark.gpu
@pytest.mark.parametrize('case_name', CASES)
def test_trim_fastq_matches_reference_cli_gpu(case_name):
    _run(case_name)


def test_trimmed_ranges_are_python_slices():
    from porechop_b200.fastq import trimmed_ranges
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 200, 500)
    st = rng.integers(0, 160, 500) * (rng.random(500) < 0.7)
    et = rng.integers(0, 160, 500) * (rng.random(500) < 0.7)
    a, b = trimmed_ranges(lens, st, et)
    for L, s, e, x, y in zip(lens, st, et, a, b):
        ref = list(range(L)) if (not s and not e) else list(range(L))[s:L - e]
        assert list(range(L))[x:y] == ref


def test_emit_chunking_and_select(monkeypatch):
    """tiny chunk size (several assemble calls) and a bin selection give the same bytes as one pass / a filtered pass."""
    from porechop_b200 import fastq
    _oracle_engine(monkeypatch)
    g = load_golden('golden_emit.json')
    b = fastq.parse_fastq(g['input_fastq'].encode())
    st = np.arange(len(b)) % 7
    et = np.arange(len(b)) % 5
    whole = fastq.emit(b, st, et)
    assert fastq.emit(b, st, et, chunk_bytes=300) == whole
    sel = np.arange(len(b)) % 2 == 0
    parts = [fastq.emit(b, st, et, select=(np.arange(len(b)) == i)) for i in range(len(b))]
    assert b''.join(parts) == whole
    assert fastq.emit(b, st, et, select=sel) == b''.join(p for i, p in enumerate(parts) if sel[i])
    assert fastq.emit(b, untrimmed=True, fmt='fasta').count(b'>') == len(b) - 1      # the empty read is not written
