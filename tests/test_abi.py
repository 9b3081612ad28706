"""CPU tier: the C-ABI library loads, exports every symbol include/porechop_b200.h declares, formats records like the
reference, and fails LOUDLY (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import ROOT, load_golden, oracle_record


def header_functions(header='porechop_b200.h'):
    src = open(os.path.join(ROOT, 'include', header)).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b([A-Za-z_][A-Za-z0-9_]*)\s*\(', src)) - {'defined'})


def test_library_exports_every_declared_symbol():
    from porechop_b200 import cpp_function_wrappers as W
    names = header_functions()
    assert 'adapterAlignment' in names and 'freeCString' in names and 'adapterAlignmentBatch' in names
    lib = ctypes.CDLL(W.SO_FILE_FULL)
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
    assert sorted(W.EXPORTED_SYMBOLS) == names


def test_hostio_library_exports_every_declared_symbol():
    from porechop_b200 import hostio
    names = header_functions('porechop_b200_io.h')
    lib = ctypes.CDLL(hostio._PATH)
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
    assert sorted(hostio.EXPORTED_SYMBOLS) == names


def test_format_record_matches_reference_strings():
    from porechop_b200 import cpp_function_wrappers as W
    from porechop_b200.align import record_string
    for rd, ad, sc, exp in load_golden('golden_random.json')[:600]:
        rec = oracle_record(rd, ad, sc)
        assert W.format_record(rec) == exp
        assert record_string(rec) == exp


def test_scores_from_records_equal_reference_parse():
    from porechop_b200.align import scores_from_records
    cases = load_golden('golden_random.json')[:800]
    recs = np.array([oracle_record(rd, ad, sc) for rd, ad, sc, _ in cases], dtype=np.int32)
    full, part, rs, re_ = scores_from_records(recs)
    for k, (_, _, _, exp) in enumerate(cases):
        p = exp.split(',')
        if int(p[0]) == -1:
            assert (full[k], part[k], rs[k], re_[k]) == (0.0, 0.0, -1, 0)
            continue
        ef, ep = float(p[6]), float(p[5])
        assert rs[k] == int(p[0]) and re_[k] == int(p[1]) + 1
        assert (full[k] == ef) or (np.isnan(full[k]) and np.isnan(ef))
        assert (part[k] == ep) or (np.isnan(part[k]) and np.isnan(ep))


def _no_gpu():
    from porechop_b200 import cpp_function_wrappers as W
    return W.device_count() == 0


@pytest.mark.skipif(not _no_gpu(), reason='a CUDA device is present')
def test_no_device_fails_loudly_no_cpu_fallback():
    from porechop_b200 import cpp_function_wrappers as W
    with pytest.raises(W.EngineError):
        W.adapter_alignment('ACGT', 'ACGT', [3, -6, -5, -2])
    buf, off = W.pack_sequences(['ACGTACGT'])
    abuf, aoff = W.pack_sequences(['ACGT'], offset_dtype=np.int32)
    with pytest.raises(W.EngineError) as e:
        W.adapter_alignment_batch(buf, off, abuf, aoff, [3, -6, -5, -2])
    assert '100' in str(e.value)


def test_invalid_arguments_are_rejected_before_the_device_is_touched():
    """Argument validation is pure host code (engine.cu validate_and_plan): a bad call fails with PB200_ERR_ARG (102)
    whether or not a device is present."""
    from porechop_b200 import cpp_function_wrappers as W
    sc = [3, -6, -5, -2]
    buf, off = W.pack_sequences(['ACGTACGT', 'ACGT', 'TTTTTT'])
    abuf, aoff = W.pack_sequences(['ACGT', 'GG'], offset_dtype=np.int32)

    def err(*a, **k):
        with pytest.raises(W.EngineError) as e:
            W.adapter_alignment_batch(*a, **k)
        return str(e.value)

    # (cross mode checks the sequence offsets chunk by chunk while the pipeline runs -- tests/test_gpu_zz_options.py)
    bad_off = off.copy(); bad_off[1], bad_off[2] = off[2], off[1]
    bad_aoff = aoff.copy(); bad_aoff[1] = 7
    assert 'error 102' in err(buf, off, abuf, bad_aoff, sc)
    assert 'error 102' in err(buf, off, abuf, aoff, sc, pair_seq=[0, 3], pair_adapter=[0, 1])      # sequence index out of range
    assert 'error 102' in err(buf, off, abuf, aoff, sc, pair_seq=[0, 1], pair_adapter=[0, -1])
    assert 'error 102' in err(buf, bad_off, abuf, aoff, sc, pair_seq=[0, 1], pair_adapter=[0, 1])
    rc = W.C_LIB.adapterAlignmentBatch(None, None, 1, None, None, 1, None, None, 1, 3, -6, -5, -2, None)
    assert rc == 102                                                                              # NULL pointers
    rc = W.C_LIB.adapterAlignmentBatch(None, None, 2, None, None, 2, None, None, 3, 3, -6, -5, -2, None)
    assert rc == 102                                                                              # cross mode: n_pairs != n_seqs*n_adapters
    if _no_gpu():
        assert 'error 100' in err(buf, off, abuf, aoff, sc)                                       # valid call: only the device is missing
    # the multi-batch submit validates every batch the same way
    with pytest.raises(W.EngineError) as e:
        W.adapter_alignment_batch_multi([(buf, off, abuf, aoff), (buf, off, abuf, bad_aoff)], sc)
    assert 'error 102' in str(e.value)
    assert W.adapter_alignment_batch_multi([], sc) == []
    empty = W.adapter_alignment_batch_multi([(buf[:0], off[:1], abuf, aoff)], sc)                 # no sequences: nothing to do
    assert len(empty) == 1 and empty[0].shape == (0, 9)
    if _no_gpu():
        with pytest.raises(W.EngineError) as e:
            W.adapter_alignment_batch_multi([(buf, off, abuf, aoff), (buf, off, abuf, aoff)], sc)
        assert 'error 100' in str(e.value)


def test_end_decisions_arguments_validated_on_the_host():
    from porechop_b200 import cpp_function_wrappers as W
    sc = [3, -6, -5, -2]
    buf, off = W.pack_sequences(['ACGTACGT', 'ACGT', 'TTTTTT'])
    abuf, aoff = W.pack_sequences(['ACGT', 'GG'], offset_dtype=np.int32)

    def err(batches, thr=75.0):
        with pytest.raises(W.EngineError) as e:
            W.adapter_end_decisions(batches, sc, 150, 2, thr, 4)
        return str(e.value)
    assert 'error 102' in err([(buf, off, abuf, aoff, True, [0, 2])])            # score column out of range
    assert 'error 102' in err([(buf, off, abuf, aoff, True, [])], thr=-1.0)      # negative threshold: host rule only
    bad_aoff = aoff.copy(); bad_aoff[1] = 7
    assert 'error 102' in err([(buf, off, abuf, aoff, True, []), (buf, off, abuf, bad_aoff, False, [1])])
    # no adapters: nothing aligns, nothing is trimmed -- decided without a device
    (trim, pairs, rec), = W.adapter_end_decisions([(buf, off, abuf[:0], aoff[:1], True, [])], sc, 150, 2, 75.0, 4)
    assert trim.tolist() == [0, 0, 0] and pairs.shape == (3, 0, 2) and rec is None
    if _no_gpu():
        assert 'error 100' in err([(buf, off, abuf, aoff, True, [1])])


def test_empty_inputs_need_no_device():
    # the -1 record of an empty read/adapter is produced without touching the device (reference: no DP either)
    from porechop_b200 import cpp_function_wrappers as W
    assert W.adapter_alignment('', 'ACGT', [3, -6, -5, -2]) == '-1,0,-1,0,-2147483648,0.000000,0.000000'
    assert W.adapter_alignment('ACGT', '', [3, -6, -5, -2]) == '-1,0,-1,0,-2147483648,0.000000,0.000000'


def test_scores_native_equals_numpy(monkeypatch):
    """libhostio's pbioScores and the numpy implementation of scores_from_records agree (incl. NaN for 0/0, failed
    alignments, and lengths beyond the table bound, which take the per-record path)."""
    from porechop_b200 import align, hostio
    assert hostio.LIB is not None
    rng = np.random.default_rng(3)
    n = 20000
    r = np.zeros((n, 9), dtype=np.int32)
    r[:, 6] = rng.integers(0, 300, n)
    r[:, 5] = (r[:, 6] * rng.random(n)).astype(np.int32)
    r[:, 8] = rng.integers(0, 120, n)
    r[:, 7] = (r[:, 8] * rng.random(n)).astype(np.int32)
    r[:, 0] = rng.integers(0, 100, n)
    r[:, 1] = r[:, 0] + rng.integers(0, 50, n)
    r[::97, 0] = -1
    r[::97, 4] = -2147483648
    big = r[:500].copy()
    big[:, 6] += 5000                                 # beyond the dense table: per-record formatting
    for rec in (r, big, r[:0]):
        native = align.scores_from_records(rec)
        monkeypatch.setattr(hostio, 'LIB', None)
        slow = align.scores_from_records(rec)
        monkeypatch.undo()
        for a, b in zip(native, slow):
            assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)
