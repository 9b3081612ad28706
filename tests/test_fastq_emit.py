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


def _oracle_decisions(monkeypatch):
    """adapterEndDecisions stand-in for the CPU tier: oracle records reduced by the decision kernel's own core functions
    (dp_core.cuh end_trim_candidate / score_pair through tests/emu) with the product's threshold table."""
    import ctypes
    from helpers import emu_lib
    from porechop_b200 import fastq
    emu = emu_lib()
    emu.emu_decide.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int, ctypes.c_int32, ctypes.c_int32,
                               ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                               ctypes.c_void_p]
    emu.emu_decide.restype = ctypes.c_int
    calls = []

    def fake(batches, scoring, end_size, extra_trim_size, end_threshold, min_trim_size, want_records=False, want_top2=False):
        outs = []
        for seq_buf, seq_off, ad_buf, ad_off, is_start, cols in batches:
            seq_off, ad_off = np.asarray(seq_off, dtype=np.int64), np.asarray(ad_off, dtype=np.int32)
            n, na = len(seq_off) - 1, len(ad_off) - 1
            cols = np.ascontiguousarray(cols, dtype=np.int32)
            trim = np.zeros(n, dtype=np.int32)
            pairs = np.zeros((n, len(cols)), dtype=np.uint32)
            if n and na:
                rec = np.ascontiguousarray(oracle_batch(np.asarray(seq_buf), seq_off, np.asarray(ad_buf), ad_off, list(scoring)))
                L = int(np.diff(seq_off).max()) + int(np.diff(ad_off).max()) + 2
                cmin = fastq.W.trim_threshold_table(end_threshold, L)
                ovf = emu.emu_decide(rec.ctypes.data, n, na, 1 if is_start else 0, end_size, extra_trim_size, min_trim_size,
                                     cmin.ctypes.data, L, cols.ctypes.data, len(cols), trim.ctypes.data, pairs.ctypes.data)
                assert ovf == 0
            p16 = np.stack([(pairs & 0xFFFF).astype(np.uint16), (pairs >> 16).astype(np.uint16)], axis=-1)
            if want_top2:
                # the device ranking restated with exact fractions: identity descending, ties in column order
                from fractions import Fraction
                top2 = np.tile(np.array([-1, 0, 1], dtype=np.int32), (n, 2))
                for i in range(n):
                    order = sorted(range(len(cols)), key=lambda k: (-Fraction(int(p16[i, k, 0]), max(int(p16[i, k, 1]), 1)), k))
                    for t, k in enumerate(order[:2]):
                        top2[i, 3 * t:3 * t + 3] = (k, p16[i, k, 0], p16[i, k, 1])
                outs.append((trim, top2, None))
                continue
            outs.append((trim, p16, None))
        calls.append(len(batches))
        return outs
    monkeypatch.setattr(fastq.W, 'adapter_end_decisions', fake)
    monkeypatch.setattr(fastq, 'DEVICE_DECISIONS', True)
    return calls


@pytest.mark.parametrize('case_name', CASES)
def test_trim_fastq_device_decisions_match_reference_cli(monkeypatch, case_name):
    """The flat pipeline with the end-trim decisions taken by the device path (adapterEndDecisions; here its core
    functions on oracle records): output files still byte-identical to the reference CLI."""
    _oracle_engine(monkeypatch)
    calls = _oracle_decisions(monkeypatch)
    _run(case_name)
    assert calls == [2]                                   # one submit: start + end windows


@pytest.mark.parametrize('case_name', BARCODE_CASES if 'BARCODE_CASES' in globals() else
                         ['bins_default', 'bins_two_barcodes', 'bins_loose_discard', 'bins_fasta_untrimmed'])
def test_demux_fastq_device_decisions_match_reference_cli(monkeypatch, case_name):
    """Barcode bins from the device's score pairs instead of the full records: every bin file byte-identical."""
    _oracle_engine(monkeypatch)
    calls = _oracle_decisions(monkeypatch)
    info = _run_demux(case_name)
    assert calls == [2] and len(info['calls']) == info['n_reads'] == 18


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


BARCODE_CASES = ['bins_default', 'bins_two_barcodes', 'bins_loose_discard', 'bins_fasta_untrimmed']


def _run_demux(case_name):
    from porechop_b200.fastq import demux_fastq
    g = load_golden('golden_emit.json')
    c = g['barcode_cases'][case_name]
    bins, info = demux_fastq(g['barcoded_fastq'].encode(), c['matching_sets'], c['scoring'], **c['options'])
    fmt = c['options']['fmt']
    assert sorted(k + '.' + fmt for k in bins) == sorted(c['bins'])
    for k, v in bins.items():
        assert v.decode() == c['bins'][k + '.' + fmt], k
    return info


@pytest.mark.parametrize('case_name', BARCODE_CASES)
def test_demux_fastq_matches_reference_cli_oracle_engine(monkeypatch, case_name):
    """`porechop -b dir`: every bin file of the reference CLI, byte for byte (barcode calls incl. --require_two_barcodes,
    thresholds, --discard_unassigned, --untrimmed)."""
    _oracle_engine(monkeypatch)
    info = _run_demux(case_name)
    assert len(info['calls']) == info['n_reads'] == 18


@pytest.mark.gpu
@pytest.mark.parametrize('case_name', BARCODE_CASES)
def test_demux_fastq_matches_reference_cli_gpu(case_name):
    _run_demux(case_name)


def _call_one(start_items, end_items, thr, diff, two):
    """scalar restatement of the reference's rule (nanopore_read.py:399-470) on (name, score) lists: the checker."""
    sd, ed = {}, {}
    for k, v in start_items:
        sd[k] = v
    for k, v in end_items:
        ed[k] = v
    ss = sorted(sd.items(), reverse=True, key=lambda x: x[1])
    es = sorted(ed.items(), reverse=True, key=lambda x: x[1])
    none = ('none', 0.0)
    bs, b2s = (ss + [none, none])[0], (ss + [none, none])[1]
    be, b2e = (es + [none, none])[0], (es + [none, none])[1]
    if two:
        ok = bs[1] >= thr and be[1] >= thr and bs[1] >= b2s[1] + diff and be[1] >= b2e[1] + diff and bs[0] == be[0]
        return bs[0] if ok else 'none'
    seen, merged = set(), []
    for k, v in sorted(ss + es, reverse=True, key=lambda x: x[1]):
        if k not in seen:
            merged.append((k, v))
            seen.add(k)
    b, b2 = (merged + [none, none])[0], (merged + [none, none])[1]
    return b[0] if (b[1] >= thr and b[1] >= b2[1] + diff) else 'none'


def test_call_barcodes_against_scalar_rule():
    """random score matrices from a small value set (many ties), repeated barcode names (dict semantics: first
    position, last value), empty sides, all option combinations."""
    from porechop_b200.fastq import call_barcodes
    rng = np.random.default_rng(11)
    vals = np.array([0.0, 55.0, 70.0, 75.0, 80.0, 85.0, 90.0, 100.0])
    pool = ['BC01', 'BC02', 'BC03', 'BC04', 'BC05']
    for trial in range(60):
        ns, ne = int(rng.integers(0, 6)), int(rng.integers(0, 6))
        s_names = [pool[i] for i in rng.integers(0, 5, ns)]
        e_names = [pool[i] for i in rng.integers(0, 5, ne)]
        S, E = vals[rng.integers(0, len(vals), (40, ns))], vals[rng.integers(0, len(vals), (40, ne))]
        for thr, diff, two in [(75.0, 5.0, False), (75.0, 5.0, True), (60.0, 0.0, False), (85.0, 10.0, True)]:
            got = call_barcodes(S, s_names, E, e_names, thr, diff, two)
            want = [_call_one(list(zip(s_names, S[i])), list(zip(e_names, E[i])), thr, diff, two) for i in range(40)]
            assert got == want, (trial, thr, diff, two)
    assert call_barcodes(np.zeros((2, 0)), [], np.zeros((2, 0)), []) == ['none', 'none']
    got = call_barcodes(np.array([[95.0], [95.0]]), ['BC01'], np.zeros((2, 0)), [], albacore_calls=['BC02', None])
    assert got == ['none', 'BC01']


def test_flat_pipeline_bench_tool_dry_run(monkeypatch):
    """tools/flat_pipeline_bench.py (the GPU-side stage timer) runs end to end; here with the oracle as the engine."""
    import importlib.util
    import os
    from helpers import ROOT
    _oracle_engine(monkeypatch)
    spec = importlib.util.spec_from_file_location('flat_pipeline_bench', os.path.join(ROOT, 'tools', 'flat_pipeline_bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.run(24, 1)
    assert r['reads'] == 24 and r['out_bytes'] > 0 and set(r['seconds']) == {'parse', 'end_trim', 'middle', 'emit'}
