"""
ctypes wrapper around cpp_functions.so -- the B200 adapter-alignment engine.

Mirror of the reference's porechop/cpp_function_wrappers.py (same module name, same library file name next
to the module, same `adapter_alignment(read_sequence, adapter_sequence, scoring_scheme_vals) -> str`
signature and result string, reference lines 21-63), plus the batched call the Python host uses to submit
every (read window, adapter) / (full read, adapter) pair at once.

There is no CPU fallback: if cpp_functions.so is missing the import exits like the reference does
(cpp_function_wrappers.py:23-24), and every call raises if no sm_100 CUDA device is usable.
"""

import os
import sys
from ctypes import CDLL, POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_longlong, c_void_p, cast, \
    create_string_buffer

import numpy as np

SO_FILE = 'cpp_functions.so'
SO_FILE_FULL = os.path.join(os.path.dirname(os.path.realpath(__file__)), SO_FILE)
if not os.path.isfile(SO_FILE_FULL):
    sys.exit('could not find ' + SO_FILE + ' - please reinstall (python -m porechop_b200.build)')
C_LIB = CDLL(SO_FILE_FULL)

# ---- reference ABI (porechop/include/adapter_align.h:13-15) ----
C_LIB.adapterAlignment.argtypes = [c_char_p,  # Read sequence
                                   c_char_p,  # Adapter sequence
                                   c_int,     # Match score
                                   c_int,     # Mismatch score
                                   c_int,     # Gap open score
                                   c_int]     # Gap extension score
C_LIB.adapterAlignment.restype = c_void_p     # String describing alignment
C_LIB.freeCString.argtypes = [c_void_p]
C_LIB.freeCString.restype = None

# ---- batched ABI (include/porechop_b200.h) ----
C_LIB.adapterAlignmentBatch.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                        c_int64, c_int, c_int, c_int, c_int, c_void_p]
C_LIB.adapterAlignmentBatch.restype = c_int


class BatchDesc(Structure):
    """pb200_batch_t (include/porechop_b200.h)"""
    _fields_ = [('seqs', c_void_p), ('seq_off', c_void_p), ('n_seqs', c_int64),
                ('adapters', c_void_p), ('ad_off', c_void_p), ('n_adapters', c_int32), ('out', c_void_p)]


class EndBatchDesc(Structure):
    """pb200_end_batch_t (include/porechop_b200.h)"""
    _fields_ = [('batch', BatchDesc), ('is_start', c_int32), ('end_size', c_int32), ('extra_trim_size', c_int32),
                ('min_trim_size', c_int32), ('end_threshold', c_double), ('score_cols', c_void_p), ('n_score_cols', c_int32),
                ('trim', c_void_p), ('score_pairs', c_void_p), ('top2', c_void_p)]


C_LIB.adapterEndDecisions.argtypes = [POINTER(EndBatchDesc), c_int, c_int, c_int, c_int, c_int]
C_LIB.adapterEndDecisions.restype = c_int
C_LIB.pb200TrimThresholdTable.argtypes = [c_double, c_int32, c_void_p]
C_LIB.pb200TrimThresholdTable.restype = c_int
C_LIB.adapterAlignmentBatchMulti.argtypes = [POINTER(BatchDesc), c_int, c_int, c_int, c_int, c_int]
C_LIB.adapterAlignmentBatchMulti.restype = c_int
C_LIB.adapterAlignmentBatchDevice.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                              c_int32, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
C_LIB.adapterAlignmentBatchDevice.restype = c_int
C_LIB.pb200FormatRecord.argtypes = [c_void_p, c_char_p, c_int]
C_LIB.pb200FormatRecord.restype = c_int
C_LIB.pb200DeviceCount.restype = c_int
C_LIB.pb200SetDevice.argtypes = [c_int]
C_LIB.pb200SetDevice.restype = c_int
C_LIB.pb200Synchronize.restype = c_int
C_LIB.pb200LastError.restype = c_char_p
C_LIB.pb200KernelLaunches.restype = c_longlong
C_LIB.pb200TimingEnable.argtypes = [c_int]
C_LIB.pb200TimingEnable.restype = None
C_LIB.pb200TimingRead.argtypes = [POINTER(c_double), POINTER(c_longlong), POINTER(c_double), c_int]
C_LIB.pb200TimingRead.restype = c_int
C_LIB.pb200TimingReadKinds.argtypes = [POINTER(c_double), POINTER(c_longlong), POINTER(c_double), c_int]
C_LIB.pb200TimingReadKinds.restype = c_int
C_LIB.pb200SetOption.argtypes = [c_char_p, c_char_p]
C_LIB.pb200SetOption.restype = c_int
C_LIB.pb200PackNibbles.argtypes = [c_void_p, c_int64, c_void_p, c_int]
C_LIB.pb200PackNibbles.restype = c_int

EXPORTED_SYMBOLS = ['adapterAlignment', 'freeCString', 'adapterAlignmentBatch', 'adapterAlignmentBatchMulti',
                    'adapterAlignmentBatchDevice', 'adapterEndDecisions', 'pb200TrimThresholdTable',
                    'pb200FormatRecord', 'pb200DeviceCount', 'pb200SetDevice', 'pb200Synchronize', 'pb200LastError',
                    'pb200KernelLaunches', 'pb200TimingEnable', 'pb200TimingRead', 'pb200TimingReadKinds', 'pb200SetOption',
                    'pb200GetOption', 'pb200HostBuffer', 'pb200PackNibbles']

RECORD_INTS = 9
SCORE_EMPTY = -2147483648


class EngineError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise EngineError('porechop_b200 engine error %d: %s' % (rc, C_LIB.pb200LastError().decode()))


def adapter_alignment(read_sequence, adapter_sequence, scoring_scheme_vals):
    """
    Python wrapper for the adapterAlignment C function (same contract as the reference,
    cpp_function_wrappers.py:42-53): returns 'rs,re,as,ae,score,aligned%ID,full%ID'.
    """
    match_score = scoring_scheme_vals[0]
    mismatch_score = scoring_scheme_vals[1]
    gap_open_score = scoring_scheme_vals[2]
    gap_extend_score = scoring_scheme_vals[3]
    ptr = C_LIB.adapterAlignment(read_sequence.encode('utf-8'), adapter_sequence.encode('utf-8'),
                                 match_score, mismatch_score, gap_open_score, gap_extend_score)
    if not ptr:
        raise EngineError('porechop_b200: adapterAlignment failed: ' + C_LIB.pb200LastError().decode())
    return c_string_to_python_string(ptr)


def c_string_to_python_string(c_string):
    """
    Casts a C string to a Python string and then frees the C string (reference lines 56-63).
    """
    python_string = cast(c_string, c_char_p).value.decode()
    C_LIB.freeCString(c_string)
    return python_string


# ---------------------------------------------------------------------------------------------------------
def pack_sequences(seqs, offset_dtype=np.int64):
    """list of str/bytes -> (uint8 buffer, offsets[n+1]).  No per-base Python work (one join + frombuffer)."""
    bs = [s.encode('ascii', 'replace') if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=offset_dtype)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:])
    buf = np.frombuffer(b''.join(bs), dtype=np.uint8) if bs and off[-1] else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(buf), off


def _ptr(a):
    return a.ctypes.data_as(c_void_p) if a is not None else None


def adapter_alignment_batch(seq_buf, seq_off, ad_buf, ad_off, scoring_scheme_vals, pair_seq=None, pair_adapter=None,
                            out=None):
    """
    Batched alignment through the C-ABI with HOST buffers (numpy arrays; pinned memory works too if the
    arrays wrap it).  seq_buf/ad_buf: uint8, seq_off: int64[n_seqs+1], ad_off: int32[n_adapters+1].
    pair_seq/pair_adapter: int32[n_pairs] or both None for the full cross product (sequence-major).
    Returns int32[n_pairs, 9] records (see include/porechop_b200.h).
    """
    seq_buf = np.ascontiguousarray(seq_buf, dtype=np.uint8)
    ad_buf = np.ascontiguousarray(ad_buf, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    ad_off = np.ascontiguousarray(ad_off, dtype=np.int32)
    n_seqs, n_ad = len(seq_off) - 1, len(ad_off) - 1
    if pair_seq is None:
        n_pairs = n_seqs * n_ad
    else:
        pair_seq = np.ascontiguousarray(pair_seq, dtype=np.int32)
        pair_adapter = np.ascontiguousarray(pair_adapter, dtype=np.int32)
        n_pairs = len(pair_seq)
    if out is None:
        out = np.empty((n_pairs, RECORD_INTS), dtype=np.int32)
    ma, mi, go, ge = [int(x) for x in scoring_scheme_vals]
    _check(C_LIB.adapterAlignmentBatch(_ptr(seq_buf), _ptr(seq_off), n_seqs, _ptr(ad_buf), _ptr(ad_off), n_ad,
                                       _ptr(pair_seq), _ptr(pair_adapter), n_pairs, ma, mi, go, ge, _ptr(out)))
    return out


def adapter_alignment_batch_multi(batches, scoring_scheme_vals):
    """
    Several cross-product batches in one submit (adapterAlignmentBatchMulti): `batches` is a list of
    (seq_buf, seq_off, ad_buf, ad_off) or (seq_buf, seq_off, ad_buf, ad_off, out) tuples with the array conventions of
    adapter_alignment_batch.  Returns the list of int32[n_seqs * n_adapters, 9] record arrays, one per batch.
    """
    descs = (BatchDesc * max(len(batches), 1))()
    keep, outs = [], []
    for k, b in enumerate(batches):
        seq_buf = np.ascontiguousarray(b[0], dtype=np.uint8)
        seq_off = np.ascontiguousarray(b[1], dtype=np.int64)
        ad_buf = np.ascontiguousarray(b[2], dtype=np.uint8)
        ad_off = np.ascontiguousarray(b[3], dtype=np.int32)
        n_seqs, n_ad = len(seq_off) - 1, len(ad_off) - 1
        out = b[4] if len(b) > 4 and b[4] is not None else np.empty((n_seqs * n_ad, RECORD_INTS), dtype=np.int32)
        keep.append((seq_buf, seq_off, ad_buf, ad_off, out))
        outs.append(out)
        descs[k] = BatchDesc(seq_buf.ctypes.data, seq_off.ctypes.data, n_seqs, ad_buf.ctypes.data, ad_off.ctypes.data, n_ad,
                             out.ctypes.data)
    ma, mi, go, ge = [int(x) for x in scoring_scheme_vals]
    _check(C_LIB.adapterAlignmentBatchMulti(descs, len(batches), ma, mi, go, ge))
    return outs


def adapter_end_decisions(batches, scoring_scheme_vals, end_size, extra_trim_size, end_threshold, min_trim_size,
                          want_records=False, out_arrays=None, want_top2=False):
    """
    End-trim decisions on the device (adapterEndDecisions): `batches` is a list of
    (seq_buf, seq_off, ad_buf, ad_off, is_start, score_cols) -- windows x adapters, which trim rule, and the adapter
    indices whose full-adapter identity the host still needs (barcode columns; may be empty).
    Returns one (trim int32[n], pairs uint16[n, n_cols, 2], records or None) per batch.
    out_arrays: optional preallocated [(trim, pairs, records-or-None), ...] (e.g. views of pinned memory) to write into.
    want_top2: the barcode ranking stays on the device too -- the second element of every result is then int32[n, 6] =
    (position in score_cols, match_ad, len_ad) of the best and the second-best score column (determine_barcode's sorted order:
    identity descending, ties in score_cols order; position -1 = no such column) instead of the pairs of all columns.
    """
    descs = (EndBatchDesc * max(len(batches), 1))()
    keep, outs = [], []
    for k, b in enumerate(batches):
        seq_buf = np.ascontiguousarray(b[0], dtype=np.uint8)
        seq_off = np.ascontiguousarray(b[1], dtype=np.int64)
        ad_buf = np.ascontiguousarray(b[2], dtype=np.uint8)
        ad_off = np.ascontiguousarray(b[3], dtype=np.int32)
        cols = np.ascontiguousarray(b[5] if b[5] is not None else [], dtype=np.int32)
        n_seqs, n_ad = len(seq_off) - 1, len(ad_off) - 1
        if out_arrays is not None:
            trim, pairs, rec = out_arrays[k]
            assert trim.dtype == np.int32 and len(trim) == n_seqs and trim.flags.c_contiguous and pairs.flags.c_contiguous
            if want_top2:
                assert pairs.dtype == np.int32 and pairs.shape == (n_seqs, 6)
            else:
                assert pairs.dtype == np.uint16 and pairs.shape == (n_seqs, len(cols), 2)
        else:
            trim = np.zeros(n_seqs, dtype=np.int32)
            if want_top2:
                pairs = np.tile(np.array([-1, 0, 1], dtype=np.int32), (n_seqs, 2))
            else:
                pairs = np.zeros((n_seqs, len(cols), 2), dtype=np.uint16)
            rec = np.empty((n_seqs * n_ad, RECORD_INTS), dtype=np.int32) if want_records else None
        keep.append((seq_buf, seq_off, ad_buf, ad_off, cols))
        outs.append((trim, pairs, rec))
        descs[k] = EndBatchDesc(BatchDesc(seq_buf.ctypes.data, seq_off.ctypes.data, n_seqs, ad_buf.ctypes.data,
                                          ad_off.ctypes.data, n_ad, rec.ctypes.data if rec is not None else None),
                                1 if b[4] else 0, int(end_size), int(extra_trim_size), int(min_trim_size), float(end_threshold),
                                cols.ctypes.data if len(cols) else None, len(cols), trim.ctypes.data,
                                pairs.ctypes.data if (len(cols) and not want_top2) else None,
                                pairs.ctypes.data if (len(cols) and want_top2) else None)
    ma, mi, go, ge = [int(x) for x in scoring_scheme_vals]
    _check(C_LIB.adapterEndDecisions(descs, len(batches), ma, mi, go, ge))
    return outs


def trim_threshold_table(end_threshold, length):
    """cmin[l] of pb200TrimThresholdTable: smallest match count whose float("%f") identity exceeds end_threshold."""
    t = np.zeros(length, dtype=np.int32)
    rc = C_LIB.pb200TrimThresholdTable(float(end_threshold), int(length), _ptr(t))
    if rc != 0:
        raise EngineError('pb200TrimThresholdTable: error %d' % rc)
    return t


def adapter_alignment_batch_device(d_seqs_ptr, d_seq_off_ptr, n_seqs, total_seq_bytes, max_seq_len, ad_buf, ad_off,
                                   scoring_scheme_vals, d_out_ptr, stream_ptr=0):
    """Cross-product batch with the bulk data already in device memory (raw device pointers as ints)."""
    ad_buf = np.ascontiguousarray(ad_buf, dtype=np.uint8)
    ad_off = np.ascontiguousarray(ad_off, dtype=np.int32)
    ma, mi, go, ge = [int(x) for x in scoring_scheme_vals]
    _check(C_LIB.adapterAlignmentBatchDevice(c_void_p(d_seqs_ptr), c_void_p(d_seq_off_ptr), n_seqs, total_seq_bytes,
                                             max_seq_len, _ptr(ad_buf), _ptr(ad_off), len(ad_off) - 1, ma, mi, go, ge,
                                             c_void_p(d_out_ptr), c_void_p(stream_ptr)))


def synchronize():
    _check(C_LIB.pb200Synchronize())


def format_record(rec):
    """One 9-int record -> the reference result string."""
    rec = np.ascontiguousarray(rec, dtype=np.int32)
    buf = create_string_buffer(96)
    n = C_LIB.pb200FormatRecord(_ptr(rec), buf, 96)
    if n < 0:
        raise EngineError('format failed')
    return buf.value.decode()


def device_count():
    return int(C_LIB.pb200DeviceCount())


def kernel_launches():
    return int(C_LIB.pb200KernelLaunches())


def timing_enable(on=True):
    C_LIB.pb200TimingEnable(1 if on else 0)


def timing_read(reset=True):
    ms, n, cells = c_double(0), c_longlong(0), c_double(0)
    _check(C_LIB.pb200TimingRead(ms, n, cells, 1 if reset else 0))
    return ms.value, n.value


TIMING_KINDS = ('trace_kernel', 'trace_kernel<score-only>', 'trace_kernel<window pass>', 'score_kernel')


def timing_read_kinds(reset=True):
    """{kind: {'ms': total CUDA-event ms, 'n': launches[, 'cells': DP cells of the window pass]}} of the timed DP launches"""
    ms, n, wc = (c_double * 4)(), (c_longlong * 4)(), c_double(0)
    _check(C_LIB.pb200TimingReadKinds(ms, n, wc, 1 if reset else 0))
    out = {}
    for k, name in enumerate(TIMING_KINDS):
        if n[k]:
            out[name] = {'ms': ms[k], 'n': int(n[k])}
            if k == 2:
                out[name]['cells'] = wc.value
    return out


def pack_nibbles(ascii_buf, threads=0):
    """uint8 ASCII bases -> uint8[(n+1)//2], two 4-bit Dna5 codes per byte (the host half of option h2d_pack)."""
    a = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
    out = np.zeros((len(a) + 1) // 2, dtype=np.uint8)
    _check(C_LIB.pb200PackNibbles(_ptr(a), len(a), _ptr(out), int(threads)))
    return out


def set_option(name, value):
    _check(C_LIB.pb200SetOption(name.encode(), str(value).encode()))


def pinned_buffer(slot, nbytes):
    """uint8[nbytes] view of the library's pinned staging buffer `slot` (pb200HostBuffer), or None without a device.  The
    view is valid until the next call for the same slot."""
    import ctypes
    C_LIB.pb200HostBuffer.argtypes = [c_int, ctypes.c_size_t]
    C_LIB.pb200HostBuffer.restype = c_void_p
    p = C_LIB.pb200HostBuffer(int(slot), int(max(nbytes, 1)))
    if not p:
        return None
    return np.ctypeslib.as_array(cast(p, POINTER(ctypes.c_uint8)), shape=(int(max(nbytes, 1)),))[:int(nbytes)]


def get_option(name):
    """current integer value of a tunable (pb200GetOption); 'h2d_pack_large_submit' = what h2d_pack=auto resolves to here"""
    C_LIB.pb200GetOption.argtypes = [c_char_p]
    C_LIB.pb200GetOption.restype = c_int
    return int(C_LIB.pb200GetOption(name.encode()))
