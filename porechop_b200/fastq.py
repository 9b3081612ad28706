"""
Bulk FASTQ ingest and end-trim decisions without per-read Python objects (host side).

The reference loads every read into a Python `NanoporeRead` (misc.py:109-168) and decides trims read by read
(nanopore_read.py:166-208).  For 10^6-10^7 reads that host work would hide the GPU gain, so this
module keeps reads as flat numpy buffers end to end:

  parse_fastq(data)          4-line FASTQ bytes -> flat sequence / quality buffers + offsets (vectorised newline scan)
  end_windows(...)           the `seq[:end_size]` / `seq[-end_size:]` windows of every read as one ragged batch
  end_trim_amounts(...)      the reference's start/end trim rule applied to whole record arrays at once
  trim_end_adapters(...)     parse -> windows -> two batched engine calls -> trim amounts (a whole Phase B for one set list)

The alignment engine is case-insensitive and maps U to T itself (Dna5 table), so sequences are not rewritten; the
`rna` flag / upper-casing of nanopore_read.py:26-31 only matter when reads are written back out.
"""
import numpy as np

from . import cpp_function_wrappers as W
from .align import scores_from_records


class FastqBatch:
    """Flat view of a FASTQ chunk: read i has name names[i], bases seq[seq_off[i]:seq_off[i+1]], same for qual."""

    def __init__(self, names, seq, seq_off, qual, qual_off):
        self.names, self.seq, self.seq_off, self.qual, self.qual_off = names, seq, seq_off, qual, qual_off

    def __len__(self):
        return len(self.seq_off) - 1

    def lengths(self):
        return np.diff(self.seq_off)


def parse_fastq(data):
    """data: bytes / uint8 array of 4-line FASTQ records ('@name', bases, '+', qualities).  Returns a FastqBatch.
    Multi-line records are not supported (the reference's loader assumes 4 lines as well, misc.py:143-168)."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    if buf.size and buf[-1] != 10:
        buf = np.concatenate([buf, np.array([10], dtype=np.uint8)])
    nl = np.flatnonzero(buf == 10)
    if len(nl) % 4 != 0:
        raise ValueError('FASTQ chunk is not a whole number of 4-line records')
    n = len(nl) // 4
    starts = np.concatenate([[0], nl[:-1] + 1])            # start of every line
    ends = nl.copy()                                       # exclusive end of every line (at '\n'; strip '\r')
    cr = (ends > starts) & (buf[np.maximum(ends - 1, 0)] == 13)
    ends = ends - cr
    if n and not (buf[starts[0::4]] == ord('@')).all():
        raise ValueError('FASTQ record does not start with @')
    s0, s1 = starts[1::4], ends[1::4]
    q0, q1 = starts[3::4], ends[3::4]
    seq, seq_off = _gather_ranges(buf, s0, s1)
    qual, qual_off = _gather_ranges(buf, q0, q1)
    names = [bytes(buf[a + 1:b]).decode('ascii', 'replace') for a, b in zip(starts[0::4], ends[0::4])]
    return FastqBatch(names, seq, seq_off, qual, qual_off)


def _gather_ranges(buf, a, b):
    """concatenate buf[a[i]:b[i]] for all i -> (flat uint8, int64 offsets) without a Python loop."""
    lens = (b - a).astype(np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    if total == 0:
        return np.zeros(0, dtype=np.uint8), off
    idx = np.arange(total, dtype=np.int64) - np.repeat(off[:-1] - a, lens)
    return np.ascontiguousarray(buf[idx]), off


def end_windows(seq, seq_off, end_size):
    """(start windows, end windows) as two ragged batches (buf, off): window length = min(end_size, read length),
    exactly `seq[:end_size]` and `seq[-end_size:]` of nanopore_read.py:172,194."""
    lens = np.diff(seq_off)
    wl = np.minimum(lens, end_size).astype(np.int64)
    start = _gather_ranges(seq, seq_off[:-1], seq_off[:-1] + wl)
    end = _gather_ranges(seq, seq_off[1:] - wl, seq_off[1:])
    return start, end


def end_trim_amounts(start_records, end_records, end_size, extra_trim_size, end_threshold, min_trim_size):
    """The trim rule of find_start_trim / find_end_trim (nanopore_read.py:178-180, 200-202) on record arrays of
    shape [n_reads, n_adapters, 9].  Returns (start_trim[n], end_trim[n]) -- the max over the adapters that pass."""
    def one(rec, is_start):
        n, a = rec.shape[0], rec.shape[1]
        if a == 0:
            return np.zeros(n, dtype=np.int64)
        _, part, rs, re_ = scores_from_records(rec.reshape(-1, 9))
        part, rs, re_ = part.reshape(n, a), rs.reshape(n, a), re_.reshape(n, a)
        with np.errstate(invalid='ignore'):
            ok = (part > end_threshold) & ((re_ - rs) >= min_trim_size)
        if is_start:
            ok &= re_ != end_size
            amount = re_ + extra_trim_size
        else:
            ok &= rs != 0
            amount = (end_size - rs) + extra_trim_size
        return np.where(ok, amount, 0).max(axis=1)
    return one(np.asarray(start_records), True), one(np.asarray(end_records), False)


def trim_end_adapters(batch, start_adapters, end_adapters, scoring_scheme_vals, end_size=150, extra_trim_size=2,
                      end_threshold=75.0, min_trim_size=4):
    """Phase B for a FastqBatch and fixed adapter lists (sequences): returns (start_trim, end_trim, start_records,
    end_records).  Two batched engine calls, no per-read Python."""
    (sbuf, soff), (ebuf, eoff) = end_windows(batch.seq, batch.seq_off, end_size)
    n = len(batch)

    def run(buf, off, adapters):
        if not adapters or n == 0:
            return np.zeros((n, 0, 9), dtype=np.int32)
        abuf, aoff = W.pack_sequences(adapters, offset_dtype=np.int32)
        return W.adapter_alignment_batch(buf, off, abuf, aoff, scoring_scheme_vals).reshape(n, len(adapters), 9)
    srec, erec = run(sbuf, soff, start_adapters), run(ebuf, eoff, end_adapters)
    st, et = end_trim_amounts(srec, erec, end_size, extra_trim_size, end_threshold, min_trim_size)
    return st, et, srec, erec
