"""The simulated engine (tests/sim) built with AddressSanitizer: default paths and every opt-in path once, small inputs.
Run through tests/test_sim_engine.py::test_sim_engine_under_address_sanitizer (PB200_SIM_ASAN=1), which sets LD_PRELOAD.
"Device" buffers are heap blocks and static shared arrays are globals here, so an out-of-bounds access of a kernel is
an ASan report.  TESTS ONLY."""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), HERE]
import numpy as np              # noqa: E402
import build_sim                # noqa: E402
build_sim.build = lambda force=False: os.environ['PB200_SIM_ASAN_LIB']
import sim_engine
from helpers import oracle_batch
from porechop_b200 import workloads as wl
W = sim_engine.load()
yt, yb = wl.nsk007()
_, sw, ew = wl.synth_end_windows(257, yt, yb, seed=1)
sbuf, soff = wl.windows_to_batch(sw); abuf, aoff = wl.pack_adapters([yt])
def run(opts, fn, label):
    for k,v in opts.items(): W.set_option(k,v)
    got, exp = fn()
    for k in opts: W.set_option(k, 'auto' if k=='hbuf' else (512 if k=='direct_max' else (131072 if k=='chunk_tasks' else (1 if k in ('profile','tight_window') else 0))))
    print(label, opts, 'equal' if np.array_equal(got, exp) else 'DIFFERENT', flush=True)
f = lambda: (W.adapter_alignment_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING), oracle_batch(sbuf, soff, abuf, aoff, wl.DEFAULT_SCORING))
for opts in ({}, {'direct_max':100}, {'direct_max':100,'tight_window':0}, {'h2d_pack':1,'chunk_tasks':100}, {'hbuf':'global'}):
    run(opts, f, 'windows')
buf, off = wl.synth_reads(5, yt, yb, seed=2, chimera_p=0.5, max_len=3000); a2, o2 = wl.pack_adapters([yt, yb])
g = lambda: (W.adapter_alignment_batch(buf, off, a2, o2, wl.DEFAULT_SCORING), oracle_batch(buf, off, a2, o2, wl.DEFAULT_SCORING))
for opts in ({}, {'profile':0,'tight_window':0}, {'profile':0}, {'direct_max':100000,'hbuf':'global'}):
    run(opts, g, 'long')
starts, ends = wl.demux_adapters()
_, sw2, _ = wl.synth_end_windows(9, starts[5], ends[5], seed=5)
sb2, so2 = wl.windows_to_batch(sw2); a3, o3 = wl.pack_adapters(starts)
h = lambda: (W.adapter_alignment_batch(sb2, so2, a3, o3, wl.DEFAULT_SCORING), oracle_batch(sb2, so2, a3, o3, wl.DEFAULT_SCORING))
run({}, h, 'demux')
outs = W.adapter_end_decisions([(sbuf, soff, abuf, aoff, True, [0])], wl.DEFAULT_SCORING, 150, 2, 75.0, 4)
print('decisions', outs[0][0][:5])
ps = np.array([0,3,3,1],dtype=np.int32); pa=np.array([0,1,0,1],dtype=np.int32)
got = W.adapter_alignment_batch(buf, off, a2, o2, wl.DEFAULT_SCORING, ps, pa)
print('pairs', 'equal' if np.array_equal(got, oracle_batch(buf, off, a2, o2, wl.DEFAULT_SCORING, ps, pa)) else 'DIFFERENT')
print('ASAN RUN COMPLETE')
