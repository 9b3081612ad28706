"""
Host-side consumption of alignment records -- the batched counterpart of the reference's
`align_adapter()` (porechop/nanopore_read.py:476-491), which parses one result string into
(full_adapter_percent_identity, aligned_region_percent_identity, read_start, read_end).

`scores_from_records` does the same for a whole int32[n, 9] record array without per-alignment Python
objects, and reproduces the reference's float values bit for bit: Python sees float("%f" % (100.0*c/l)), i.e.
the percentage rounded to 6 decimals by printf and parsed back (SURVEY.md 7.4 "Float identity").
"""
import numpy as np

from .cpp_function_wrappers import SCORE_EMPTY


def _percent_exact(count, length):
    """float('%f' % (100.0*count/length)) for arrays, via the (few) unique (count, length) pairs."""
    count = np.asarray(count, dtype=np.int64)
    length = np.asarray(length, dtype=np.int64)
    key = count * (1 << 32) + length
    uniq, inv = np.unique(key, return_inverse=True)
    vals = np.empty(len(uniq), dtype=np.float64)
    for k, u in enumerate(uniq):
        c, l = int(u >> 32), int(u & 0xFFFFFFFF)
        vals[k] = float('nan') if l == 0 else float('%f' % (100.0 * c / l))
    return vals[inv].reshape(count.shape)


def scores_from_records(records):
    """
    records: int32[n, 9] from adapter_alignment_batch.
    Returns (full_score, partial_score, read_start, read_end) arrays with the exact values the reference's
    align_adapter() returns for each alignment (failed alignments: 0.0, 0.0, -1, 0).
    """
    from . import hostio
    if hostio.LIB is not None:
        return hostio.scores(records)                # the same values, one parallel pass in C (libhostio.so)
    r = np.asarray(records, dtype=np.int32).reshape(-1, 9)
    failed = (r[:, 0] == -1) & (r[:, 4] == SCORE_EMPTY)
    full = _percent_exact(r[:, 7], r[:, 8])
    part = _percent_exact(r[:, 5], r[:, 6])
    read_start = r[:, 0].astype(np.int64)
    read_end = r[:, 1].astype(np.int64) + 1
    full[failed] = 0.0
    part[failed] = 0.0
    read_end[failed] = 0
    return full, part, read_start, read_end


def record_string(rec):
    """The reference result string of one record, formatted in Python (used by tests next to pb200FormatRecord)."""
    rec = [int(x) for x in rec]
    if rec[0] == -1 and rec[4] == SCORE_EMPTY:
        return '-1,0,-1,0,-2147483648,0.000000,0.000000'

    def pct(c, l):
        return '-nan' if l == 0 else '%f' % (100.0 * c / l)
    return '%d,%d,%d,%d,%d,%s,%s' % (rec[0], rec[1], rec[2], rec[3], rec[4], pct(rec[5], rec[6]), pct(rec[7], rec[8]))
