"""
Run-time drop-in for an UNMODIFIED Porechop checkout: batch the three alignment phases without touching its code
(INTEGRATION.md section 3b).

`install()` does two things to the imported `porechop` package:

  1. `porechop.nanopore_read.adapter_alignment` (the name `align_adapter()` calls, nanopore_read.py:17,477) becomes a
     memo lookup keyed by (read window, adapter sequence): a hit returns the reference's result string formatted from
     a prefetched 9-int record; a miss falls through to the engine's per-call `adapterAlignment`.
  2. the three phase drivers of porechop.py (`find_matching_adapter_sets` :286, `find_adapters_at_read_ends` :438,
     `find_adapters_in_read_middles` :533) are wrapped: the wrapper gathers every (window, adapter) pair the phase is
     going to ask for, submits them as ONE batch (cross product) per window kind, fills the memo, then calls the
     ORIGINAL driver -- which therefore keeps its own progress lines, verbose output, thread pool, barcode calling and
     trimming arithmetic, byte for byte, and finds every alignment already computed.

Phase C's sequential '-' masking (nanopore_read.py:210-243) is prefetched with speculative rounds: round 0 is the
cross product of the trimmed reads with the middle adapters; reads with a hit are masked exactly as the reference
will mask them and re-submitted from the hit adapter on, until no read has a hit.

Unlike `phases.py` (which replaces the loops and fills the read objects itself) this keeps every line of the
reference's Python on the path; it is the lowest-risk way to put the engine under the real CLI.
"""
import inspect

import numpy as np

from . import cpp_function_wrappers as W
from .align import record_string, scores_from_records


class _Table:
    """records of a cross product: unique windows x adapters."""

    def __init__(self, windows, adapters, records):
        self.row = {w: i for i, w in enumerate(windows)}
        self.col = {a: j for j, a in enumerate(adapters)}
        self.records = records          # int32[len(windows), len(adapters), 9]

    def get(self, w, a):
        j = self.col.get(a)
        if j is None:
            return None
        i = self.row.get(w)
        if i is None:
            return None
        return self.records[i, j]


class Memo:
    """Prefetched alignment records for one scoring scheme; thread-safe for lookups (read-only while a phase runs)."""

    def __init__(self):
        self.scoring = None
        self.tables = []
        self.pairs = {}
        self.hits = 0
        self.misses = 0
        self.batches = 0

    def clear(self):
        self.tables, self.pairs = [], {}

    def lookup(self, read_seq, adapter_seq, scoring):
        if self.scoring is None or tuple(scoring) != self.scoring:
            return None
        for t in self.tables:
            rec = t.get(read_seq, adapter_seq)
            if rec is not None:
                return rec
        return self.pairs.get((read_seq, adapter_seq))

    # ---- prefetch ----
    def add_cross(self, windows, adapter_seqs, scoring):
        """every distinct window x every distinct adapter, one engine call.  Returns the _Table (or None if empty)."""
        windows = list(dict.fromkeys(windows))
        adapter_seqs = list(dict.fromkeys(adapter_seqs))
        if not windows or not adapter_seqs:
            return None
        self.scoring = tuple(int(x) for x in scoring)
        sbuf, soff = W.pack_sequences(windows, offset_dtype=np.int64)
        abuf, aoff = W.pack_sequences(adapter_seqs, offset_dtype=np.int32)
        rec = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, list(self.scoring))
        self.batches += 1
        t = _Table(windows, adapter_seqs, np.asarray(rec).reshape(len(windows), len(adapter_seqs), 9))
        self.tables.append(t)
        return t

    def add_pairs(self, seqs, adapter_seqs, pair_seq, pair_adapter, scoring):
        """explicit pair list (Phase C re-submissions)."""
        if not len(pair_seq):
            return np.zeros((0, 9), dtype=np.int32)
        self.scoring = tuple(int(x) for x in scoring)
        sbuf, soff = W.pack_sequences(seqs, offset_dtype=np.int64)
        abuf, aoff = W.pack_sequences(adapter_seqs, offset_dtype=np.int32)
        rec = W.adapter_alignment_batch(sbuf, soff, abuf, aoff, list(self.scoring),
                                        np.asarray(pair_seq, dtype=np.int32), np.asarray(pair_adapter, dtype=np.int32))
        self.batches += 1
        rec = np.asarray(rec).reshape(-1, 9)
        for k, (s, a) in enumerate(zip(pair_seq, pair_adapter)):
            self.pairs[(seqs[s], adapter_seqs[a])] = rec[k]
        return rec


def _bound(fn, args, kwargs):
    b = inspect.signature(fn).bind(*args, **kwargs)
    b.apply_defaults()
    return b.arguments


def _trimmed(read):
    # the value of NanoporeRead.get_seq_with_start_end_adapters_trimmed() (nanopore_read.py:57-63)
    return read.get_seq_with_start_end_adapters_trimmed()


def _middle_adapter_list(matching_sets):
    # the (name, sequence) list find_adapters_in_read_middles builds (porechop.py:541-548)
    out = []
    for s in matching_sets:
        if s.start_sequence:
            out.append(s.start_sequence)
        if s.end_sequence and ((not s.start_sequence) or s.end_sequence[1] != s.start_sequence[1]):
            out.append(s.end_sequence)
    return out


def prefetch_middles(memo, reads, adapters, middle_threshold, scoring):
    """Fill `memo` with every alignment find_middle_adapters() will request for `reads` (speculative rounds).
    Round 0 is vectorised over the whole cross product; only reads with a hit are touched in Python afterwards."""
    order = [a[1] for a in adapters]                 # request order (duplicates keep their place)
    ad_seqs = list(dict.fromkeys(order))
    trimmed = list(dict.fromkeys(_trimmed(r) for r in reads))
    t = memo.add_cross(trimmed, ad_seqs, scoring)
    if t is None:
        return
    n, m = len(trimmed), len(ad_seqs)
    full, _, rs, re_ = (x.reshape(n, m) for x in scores_from_records(t.records.reshape(-1, 9)))
    col = {a: j for j, a in enumerate(ad_seqs)}
    cols = [col[a] for a in order]
    hit = full[:, cols] >= middle_threshold          # [n, len(order)]; NaN never hits, as in the reference
    first_pos = hit.argmax(axis=1)
    active = []                                      # [masked sequence, position in `order` to resume from]
    for i in np.flatnonzero(hit.any(axis=1)):
        pos = int(first_pos[i])
        a_, b_ = int(rs[i, cols[pos]]), int(re_[i, cols[pos]])
        active.append([trimmed[i][:a_] + '-' * (b_ - a_) + trimmed[i][b_:], pos])
    while active:
        ps, pa, where = [], [], {}
        for k, st in enumerate(active):
            for a in dict.fromkeys(order[st[1]:]):
                where[(k, a)] = len(ps)
                ps.append(k)
                pa.append(col[a])
        rec = memo.add_pairs([st[0] for st in active], ad_seqs, ps, pa, scoring)
        full, _, rs, re_ = scores_from_records(rec)
        still = []
        for k, st in enumerate(active):
            for pos in range(st[1], len(order)):
                p = where[(k, order[pos])]
                if full[p] >= middle_threshold:
                    a_, b_ = int(rs[p]), int(re_[p])
                    still.append([st[0][:a_] + '-' * (b_ - a_) + st[0][b_:], pos])
                    break
        active = still


def install(porechop_pkg=None):
    """Patch the imported Porechop package in place.  Returns the Memo (hit / miss / batch counters) so a caller can
    check that the phases really ran batched.  `uninstall(memo)` restores the original functions."""
    if porechop_pkg is None:
        import porechop as porechop_pkg             # the user's checkout, wherever it is on sys.path
    import importlib
    P = importlib.import_module(porechop_pkg.__name__ + '.porechop')
    NR = importlib.import_module(porechop_pkg.__name__ + '.nanopore_read')
    memo = Memo()
    orig = {'adapter_alignment': NR.adapter_alignment, 'A': P.find_matching_adapter_sets,
            'B': P.find_adapters_at_read_ends, 'C': P.find_adapters_in_read_middles}

    def adapter_alignment(read_sequence, adapter_sequence, scoring_scheme_vals):
        rec = memo.lookup(read_sequence, adapter_sequence, scoring_scheme_vals)
        if rec is None:
            memo.misses += 1
            return W.adapter_alignment(read_sequence, adapter_sequence, scoring_scheme_vals)
        memo.hits += 1
        return record_string(rec)

    def find_matching_adapter_sets(*args, **kwargs):
        a = _bound(orig['A'], args, kwargs)
        reads, end_size, scoring = a['check_reads'], a['end_size'], a['scoring_scheme_vals']
        search = [s for s in P.ADAPTERS if '(full sequence)' not in s.name]       # porechop.py:296
        memo.clear()
        memo.add_cross([r.seq[:end_size] for r in reads], [s.start_sequence[1] for s in search if s.start_sequence], scoring)
        memo.add_cross([r.seq[-end_size:] for r in reads], [s.end_sequence[1] for s in search if s.end_sequence], scoring)
        try:
            return orig['A'](*args, **kwargs)
        finally:
            memo.clear()

    def find_adapters_at_read_ends(*args, **kwargs):
        a = _bound(orig['B'], args, kwargs)
        reads, sets, end_size, scoring = a['reads'], a['matching_sets'], a['end_size'], a['scoring_scheme_vals']
        memo.clear()
        memo.add_cross([r.seq[:end_size] for r in reads], [s.start_sequence[1] for s in sets if s.start_sequence], scoring)
        memo.add_cross([r.seq[-end_size:] for r in reads], [s.end_sequence[1] for s in sets if s.end_sequence], scoring)
        try:
            return orig['B'](*args, **kwargs)
        finally:
            memo.clear()

    def find_adapters_in_read_middles(*args, **kwargs):
        a = _bound(orig['C'], args, kwargs)
        memo.clear()
        prefetch_middles(memo, a['reads'], _middle_adapter_list(a['matching_sets']), a['middle_threshold'],
                         a['scoring_scheme_vals'])
        try:
            return orig['C'](*args, **kwargs)
        finally:
            memo.clear()

    NR.adapter_alignment = adapter_alignment
    P.find_matching_adapter_sets = find_matching_adapter_sets
    P.find_adapters_at_read_ends = find_adapters_at_read_ends
    P.find_adapters_in_read_middles = find_adapters_in_read_middles
    memo._orig, memo._modules = orig, (P, NR)
    return memo


def uninstall(memo):
    P, NR = memo._modules
    NR.adapter_alignment = memo._orig['adapter_alignment']
    P.find_matching_adapter_sets = memo._orig['A']
    P.find_adapters_at_read_ends = memo._orig['B']
    P.find_adapters_in_read_middles = memo._orig['C']
