#!/usr/bin/env python3
"""Time the flat FASTQ pipeline (porechop_b200/fastq.py) end to end on one GPU: synthetic FASTQ bytes -> parse -> end
trim (Phase B) -> middle scan (Phase C) -> emit, with the per-stage seconds `trim_fastq` reports.  Not a bench.py
metric (bench.py measures the alignment path itself); this is the tool for finding what the CLI-level run is bound by.

    python tools/flat_pipeline_bench.py --reads 200000 --repeat 3
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_fastq(n_reads, seed, mean_len=8000):
    """FASTQ bytes of the bench.py 'middle' read model (log-normal lengths, adapters at the ends, 5 % chimeras)."""
    import numpy as np
    from porechop_b200 import hostio, workloads as wl
    yt, yb = wl.nsk007()
    buf, off = wl.synth_reads(n_reads, yt, yb, seed=seed, chimera_p=0.05)
    n = len(off) - 1
    names = ['@read_%d ch=%d\n' % (i, i % 512) for i in range(n)]
    nbuf = np.frombuffer(''.join(names).encode(), dtype=np.uint8)
    noff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(x) for x in names], out=noff[1:])
    lens = np.diff(off)
    rec = np.diff(noff) + lens + 3 + lens + 1                # name line, bases, '\n+\n', qualities, '\n'
    roff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rec, out=roff[1:])
    out = np.full(int(roff[-1]), ord('5'), dtype=np.uint8)   # qualities are constant '5' 
    for i in range(n):                                       # generator only, not timed
        p = roff[i]
        out[p:p + len(names[i])] = nbuf[noff[i]:noff[i + 1]]
        p += len(names[i])
        out[p:p + lens[i]] = buf[off[i]:off[i + 1]]
        out[p + lens[i]:p + lens[i] + 3] = np.frombuffer(b'\n+\n', dtype=np.uint8)
        out[roff[i + 1] - 1] = 10
    return out.tobytes(), (yt, yb), hostio.LIB is not None


def run(n_reads, repeat, seed=20260923):
    from porechop_b200 import fastq, workloads as wl
    data, (yt, yb), native = synthetic_fastq(n_reads, seed)
    sets = [(('SQK-NSK007_Y_Top', yt), ('SQK-NSK007_Y_Bottom', yb))]
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        out, info = fastq.trim_fastq(data, sets, wl.DEFAULT_SCORING, as_array=True)
        total = time.perf_counter() - t0
        if best is None or total < best['seconds_total']:
            best = {'seconds_total': total, 'seconds': info['seconds'], 'out_bytes': int(len(out)),
                    'split_reads': len(info['middle'])}
    best.update({'reads': n_reads, 'in_bytes': len(data), 'reads_per_s': n_reads / best['seconds_total'],
                 'in_MB_per_s': len(data) / 1e6 / best['seconds_total'], 'native_hostio': native})
    return best


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--reads', type=int, default=100000)
    ap.add_argument('--repeat', type=int, default=3)
    a = ap.parse_args()
    print(json.dumps(run(a.reads, a.repeat)))
