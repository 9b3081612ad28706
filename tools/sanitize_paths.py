#!/usr/bin/env python3
"""Small runs of every engine path (default + opt-in) for compute-sanitizer on the GPU box:

    compute-sanitizer --tool memcheck  python tools/sanitize_paths.py
    compute-sanitizer --tool racecheck python tools/sanitize_paths.py      (shared-memory hazards: staging rings, profiles)

Inputs are tiny (the tools slow kernels down by 10-100x); results are compared with the oracle so a run also fails on
wrong records.  The host simulation (tests/sim) already runs the same paths under AddressSanitizer and with shuffled lane
order; this is the hardware-side counterpart."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import numpy as np                                  # noqa: E402
from helpers import oracle_batch                    # noqa: E402
from porechop_b200 import cpp_function_wrappers as W, workloads as wl    # noqa: E402

DEFAULTS = {'h2d_pack': 0, 'tight_window': 1, 'profile': 1, 'direct_max': 160, 'chunk_tasks': 131072,
            'hbuf': 'auto'}
yt, yb = wl.nsk007()
_, sw, ew = wl.synth_end_windows(600, yt, yb, seed=1)
sbuf, soff = wl.windows_to_batch(sw)
a1, o1 = wl.pack_adapters([yt])
a2, o2 = wl.pack_adapters([yt, yb])
lbuf, loff = wl.synth_reads(6, yt, yb, seed=2, chimera_p=0.5, max_len=4000)
starts, ends = wl.demux_adapters()
bad = 0


def run(label, opts, fn):
    global bad
    try:
        for k, v in opts.items():
            W.set_option(k, v)
        got, exp = fn()
    finally:
        for k in opts:
            W.set_option(k, DEFAULTS[k])
    ok = np.array_equal(got, exp)
    bad += 0 if ok else 1
    print('%-10s %-70s %s' % (label, opts, 'ok' if ok else 'DIFFERENT'), flush=True)


def cross(buf, off, ab, ao):
    return lambda: (W.adapter_alignment_batch(buf, off, ab, ao, wl.DEFAULT_SCORING), oracle_batch(buf, off, ab, ao, wl.DEFAULT_SCORING))


for opts in ({}, {'direct_max': 100}, {'direct_max': 100, 'tight_window': 0}, {'h2d_pack': 1, 'chunk_tasks': 200},
             {'hbuf': 'global'}):
    run('windows', opts, cross(sbuf, soff, a1, o1))
    run('windows2', opts, cross(sbuf, soff, a2, o2))
for opts in ({}, {'profile': 0, 'tight_window': 0}, {'profile': 0}, {'direct_max': 100000, 'hbuf': 'global'}):
    run('long', opts, cross(lbuf, loff, a2, o2))
a3, o3 = wl.pack_adapters(starts)
run('demux', {}, cross(sbuf[:150 * 12], soff[:13], a3, o3))
run('demux', {'direct_max': 100}, cross(sbuf[:150 * 12], soff[:13], a3, o3))
run('demux', {'hbuf': 'global'}, cross(sbuf[:150 * 12], soff[:13], a3, o3))
outs = W.adapter_end_decisions([(sbuf, soff, a2, o2, True, [0, 1])], wl.DEFAULT_SCORING, 150, 2, 75.0, 4)
print('decisions', outs[0][0][:6].tolist(), flush=True)
# barcode ranking on the device (top2) over a many-column class, checked against the host ranking of the same records
from porechop_b200 import hostio
from porechop_b200.fastq import Top2Scores, top2_from_scores
cols = list(range(0, len(starts), 7))
(trim, top2, rec), = W.adapter_end_decisions([(sbuf[:150 * 40], soff[:41], a3, o3, True, cols)], wl.DEFAULT_SCORING, 150, 2, 75.0, 4,
                                             want_top2=True, want_records=True)
exp = top2_from_scores(hostio.full_scores(rec.reshape(40, len(starts), 9), cols)) if hostio.LIB is not None else None
got = Top2Scores([str(c) for c in cols], top2).ranked()
ok = exp is None or all(np.array_equal(g, e) for g, e in zip(got, exp))
bad += 0 if ok else 1
print('top2      %s' % ('ok' if ok else 'DIFFERENT'), flush=True)
got = W.adapter_alignment_batch_multi([(sbuf, soff, a1, o1), (lbuf, loff, a2, o2)], wl.DEFAULT_SCORING)
bad += 0 if np.array_equal(got[1], oracle_batch(lbuf, loff, a2, o2, wl.DEFAULT_SCORING)) else 1
print('SANITIZE RUN COMPLETE, %d wrong' % bad, flush=True)
sys.exit(1 if bad else 0)
