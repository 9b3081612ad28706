#!/usr/bin/env python3
"""
bench.py -- BASELINE.json metric on B200: reads/s (+ GCUPS) of the adapter-alignment hot path.

Workload (default, BASELINE.json configs[1]): 1 M synthetic ~8 kb ONT reads (seed 20260923), adapter set
SQK-NSK007 (= the LSK109 Y-adapter), end-trim only: per read the 150-nt start window vs Y_Top (28 nt) and the 150-nt
end window vs Y_Bottom (22 nt) -> 2 alignments, 7 500 DP cells, 372 algorithmic bytes per read.  A "step" is one pass
of the hot path over the whole batch (two batched C-ABI calls).  Other workloads: --workload demux (configs[2]) and
--workload middle (configs[3] sample).

  value   reads/s with the windows already resident in HBM (adapterAlignmentBatchDevice), CUDA-event timed
  e2e     reads/s through the host-buffer C-ABI call (adapterAlignmentBatch, pinned host buffers): H2D + kernels +
          D2H inside the timed region, every rank on its own shard and PCIe link; the optional NCCL re-gather of the
          36-byte records to rank 0 is timed separately (e2e.gather_records_to_rank0_ms)
  roofline   dominant kernel (trace_kernel) algorithmic bytes / CUDA-event duration vs the measured HBM peak --
          reported because the north star asks for it; the kernel is integer-ALU (DPX) bound, see `alu`
  cpu_baseline / --impl reference   the reference's own C++ (oracle/_ref/cpp_functions.so, falling back to the
          oracle port) timed by the native harness on all host cores, on a bounded sample of the same pair list

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 via torchrun (one rank per GPU, weak scaling:
every rank processes its own --reads batch).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='endtrim', choices=['endtrim', 'demux', 'middle', 'sweep'])
    ap.add_argument('--sweep-lengths', default='500,1000,2000,5000,10000,20000,50000,100000',
                    help='--workload sweep: read lengths (one batch each, --sweep-bases bases per batch)')
    ap.add_argument('--sweep-bases', type=float, default=2e8, help='--workload sweep: bases per length point and rank')
    ap.add_argument('--reads', type=int, default=0, help='reads per rank (default: 1M endtrim, 32k demux, 256k middle)')
    ap.add_argument('--cpu-sample-reads', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-multi', action='store_true',
                    help="e2e: the step's batches in ONE adapterAlignmentBatchMulti submit instead of one call per batch")
    ap.add_argument('--e2e-decisions', action='store_true',
                    help='e2e (endtrim / demux): adapterEndDecisions -- trim amounts + barcode score pairs decided on the '
                         'device, 4 + 4*adapters bytes per window come back instead of 36 per alignment')
    ap.add_argument('--opt', action='append', default=[], metavar='NAME=VALUE',
                    help='engine option for this run (pb200SetOption), e.g. --opt h2d_pack=1 --opt tight_window=1')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed regions (NVML; nvidia-smi as fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self.stop_flag = False
        self.active = False
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40, 'hw_power_brake': 0x80,
                 'sw_power_cap': 0x4}
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    try:
                        r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.01)

    def summary(self):
        if not self.samples:
            try:
                out = subprocess.check_output(['nvidia-smi', '--query-gpu=clocks.sm,clocks.max.sm', '--format=csv,noheader,nounits',
                                               '-i', str(self.index)], timeout=20).decode().strip().split(',')
                return {'sm_mhz': float(out[0]), 'sm_max_mhz': float(out[1]), 'reasons': ['not sampled under load']}
            except Exception:
                return {'sm_mhz': None, 'sm_max_mhz': self.sm_max, 'reasons': ['unavailable']}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.sm_max, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
def make_workload(args, rank):
    """Returns a list of batches; each batch = (name, window matrix uint8[n, w] or (buf, off), adapter list)."""
    from porechop_b200 import workloads as wl
    seed = wl.SEED + rank
    if args.workload == 'endtrim':
        n = args.reads or 1000000
        yt, yb = wl.nsk007()
        L, sw, ew = wl.synth_end_windows(n, yt, yb, seed=seed)
        batches = [('start', wl.windows_to_batch(sw), [yt]), ('end', wl.windows_to_batch(ew), [yb])]
        desc = '%d synthetic ~8kb reads (lognormal, seed %d), SQK-NSK007 (LSK109 Y-adapter), end-trim: 150x28 + 150x22 per read' % (n, seed)
    elif args.workload == 'demux':
        n = args.reads or 32768
        starts, ends = wl.demux_adapters()
        L, sw, ew = wl.synth_end_windows(n, starts[100], ends[100], seed=seed)
        batches = [('start', wl.windows_to_batch(sw), starts), ('end', wl.windows_to_batch(ew), ends)]
        desc = '%d synthetic reads x 356 adapters (119 sets + 12 native-full + 96 rapid-full), demux end windows' % n
    elif args.workload == 'sweep':
        # BASELINE configs[4]: read-length sweep x the 192 sequences of the 96 forward barcode sets, full-read scan; the same
        # number of bases at every length so the points are comparable (one batch per length; run one length at a time with
        # --sweep-lengths L to get a per-length number)
        bcs = wl.forward_barcode_sequences()
        lengths = [int(x) for x in args.sweep_lengths.split(',')]
        batches, n = [], 0
        for L in lengths:
            k = max(1, int(args.sweep_bases // L))
            batches.append(('L%d' % L, wl.synth_fixed_length_reads(k, L, bcs, seed=seed + L), bcs))
            n += k
        desc = 'read-length sweep %s x 192 forward barcode sequences (24 nt), %.0e bases per length, full-read scan' % (
            args.sweep_lengths, args.sweep_bases)
    else:
        n = args.reads or 262144     # >= ~14 reads per resident group, so one 60-kb read is not the makespan
        yt, yb = wl.nsk007()
        buf, off = wl.synth_reads(n, yt, yb, seed=seed, chimera_p=0.05)
        batches = [('middle', (buf, off), [yt, yb])]
        desc = '%d synthetic full reads (5%% chimeras) x {Y_Top, Y_Bottom}, middle-adapter scan (two-pass)' % n
    return n, batches, desc


def batch_cells(batch):
    (buf, off), ads = batch[1], batch[2]
    return int(off[-1] - off[0]) * sum(len(a) for a in ads)


def batch_alg_bytes(batch):
    (buf, off), ads = batch[1], batch[2]
    n = len(off) - 1
    return int(off[-1] - off[0]) + sum(len(a) for a in ads) + 36 * n * len(ads)


def run_reference_harness(batches, scoring, sample_reads, threads, answers=None):
    """Time the reference CPU path (oracle/_ref/cpp_functions.so via the native harness) on the first
    `sample_reads` reads of every batch.  Returns (seconds, reads, cells, kind).  With `answers` (a list), the result
    strings of every sampled pair (pair order = read-major) are appended per batch: the checker for the parity gate."""
    from porechop_b200 import workloads as wl
    lib = os.path.join(ROOT, 'oracle', '_ref', 'cpp_functions.so')
    kind = 'reference'
    if not os.path.exists(lib):
        lib = os.path.join(ROOT, 'oracle', 'liboracle.so')
        kind = 'port'
    harness = os.path.join(ROOT, 'oracle', '_ref', 'ref_harness')
    if not os.path.exists(harness) or not os.path.exists(lib):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), 'liboracle.so', 'harness', 'ref'])
    sec = 0.0
    cells = 0
    with tempfile.TemporaryDirectory() as d:
        for name, (buf, off), ads in batches:
            k = min(sample_reads, len(off) - 1)
            abuf, aoff = wl.pack_adapters(ads)
            p = os.path.join(d, name + '.bin')
            wl.write_harness_file(p, buf[:off[k]], off[:k + 1], abuf, aoff, scoring)
            cmd = [harness, lib, p, str(threads)]
            if answers is not None:
                cmd.append(os.path.join(d, name + '.answers'))
            info = json.loads(subprocess.check_output(cmd).decode())
            sec += info['seconds']
            cells += info['cells']
            if answers is not None:
                with open(cmd[-1]) as f:
                    answers.append(f.read().split('\n')[:-1])
    return sec, min(sample_reads, len(batches[0][1][1]) - 1), cells, kind


def parity_gate(records_per_batch, answers_per_batch, format_record):
    """SURVEY 8(d) parity gate: the engine's records of the sampled pairs, rendered as the reference's result string,
    must equal the CPU reference's strings.  (An empty alignment is compared on its first field only: the reference
    leaves the others uninitialised.)  Returns {'checked': n, 'mismatches': k}."""
    checked = bad = 0
    for rec, ans in zip(records_per_batch, answers_per_batch):
        for r, a in zip(rec[:len(ans)], ans):
            g = format_record(r)
            ok = (g == a) or (a.startswith('-1,') and g.startswith('-1,'))
            checked += 1
            bad += 0 if ok else 1
    return {'checked': checked, 'mismatches': bad}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    from porechop_b200 import workloads as wl
    scoring = wl.DEFAULT_SCORING

    if args.impl == 'reference':
        if rank != 0:
            return 0
        n, batches, desc = make_workload(args, 0)
        cores = host_cores()
        per_read_cells = sum(batch_cells(b) for b in batches) / n
        # bounded sample: ~3 core-seconds per core per step at ~0.08 GCUPS/core
        sample = args.cpu_sample_reads or int(max(64, min(n, 0.08e9 * 1.5 * cores / per_read_cells)))
        times = []
        cells = 0
        for s in range(args.warmup + args.steps):
            sec, reads, cells, kind = run_reference_harness(batches, scoring, sample, cores)
            if s >= args.warmup:
                times.append(sec)
        t = float(np.mean(times))
        value = sample / t
        line = {'impl': 'reference', 'metric': 'reads/sec', 'value': value, 'unit': 'reads/s', 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic',
                'gcups': cells / t / 1e9,
                'config': {'workload': desc, 'sample': 'first %d reads of the batch per step' % sample},
                'cpu_baseline': {'value': value, 'unit': 'reads/s', 'cores': cores, 'kind': kind,
                                 'sample': 'first %d reads (%d alignments) per step, %d threads, native harness over the reference C-ABI' %
                                           (sample, sample * int(round(per_read_cells / 7500 * 2)) if args.workload == 'endtrim' else sample, cores)},
                'e2e': {'value': value, 'unit': 'reads/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from porechop_b200 import cpp_function_wrappers as W

    n, batches, desc = make_workload(args, rank)
    K, Wm = args.steps, args.warmup
    for kv in args.opt:
        name, _, val = kv.partition('=')
        W.set_option(name, val)

    # ---- host (pinned) and device copies of the inputs ----
    host, dev = [], []
    for name, (buf, off), ads in batches:
        abuf, aoff = wl.pack_adapters(ads)
        hb = torch.empty(len(buf), dtype=torch.uint8, pin_memory=True)
        hb.numpy()[:] = buf
        ho = torch.empty(len(off), dtype=torch.int64, pin_memory=True)
        ho.numpy()[:] = off
        n_pairs = (len(off) - 1) * len(ads)
        hout = torch.empty((n_pairs, 9), dtype=torch.int32, pin_memory=True)
        host.append((hb, ho, abuf, aoff, hout))
        db, do = hb.cuda(non_blocking=True), ho.cuda(non_blocking=True)
        dout = torch.empty((n_pairs, 9), dtype=torch.int32, device='cuda')
        max_len = int(np.max(np.diff(off))) if len(off) > 1 else 0
        dev.append((db, do, abuf, aoff, dout, max_len))
    torch.cuda.synchronize()
    in_bytes = sum(h[0].numel() + h[1].numel() * 8 + len(h[2]) + len(h[3]) * 4 for h in host)
    out_bytes = sum(h[4].numel() * 4 for h in host)
    cells_per_step = sum(batch_cells(b) for b in batches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the device-resident path runs on an explicit (non-default) stream; the CUDA events that time it are recorded on
    # the same stream.  (Handle 0 = "NULL" means "the library's own stream" to the C-ABI.)
    bench_stream = torch.cuda.Stream()

    def step_device():
        s = bench_stream.cuda_stream
        assert s != 0
        for db, do, abuf, aoff, dout, max_len in dev:
            W.adapter_alignment_batch_device(db.data_ptr(), do.data_ptr(), do.numel() - 1, db.numel(), max_len, abuf, aoff,
                                             scoring, dout.data_ptr(), s)

    dec_out = None
    if args.e2e_decisions:
        assert args.workload in ('endtrim', 'demux'), '--e2e-decisions applies to the end-window workloads'
        dec_out = []
        for (name, _, ads), (hb, ho, abuf, aoff, hout) in zip(batches, host):
            nw, na = ho.numel() - 1, len(ads)
            dec_out.append((torch.empty(nw, dtype=torch.int32, pin_memory=True),
                            torch.empty((nw, na, 2), dtype=torch.uint16, pin_memory=True)))
        out_bytes_dec = sum(t.numel() * 4 + p.numel() * 2 for t, p in dec_out)

    def step_e2e():
        # every rank pushes its own shard through the host-buffer C-ABI over its own PCIe link; the records stay
        # rank-local (no data-path collective, prompt (5)); the optional re-gather to a writer rank is timed separately
        if args.e2e_decisions:
            # every adapter is a score column (upper bound of what barcode calling needs)
            W.adapter_end_decisions([(hb.numpy(), ho.numpy(), abuf, aoff, b[0] == 'start', list(range(len(b[2]))))
                                     for b, (hb, ho, abuf, aoff, hout) in zip(batches, host)], scoring, wl.END_SIZE, 2, 75.0, 4,
                                    out_arrays=[(t.numpy(), p.numpy(), None) for t, p in dec_out])
            return
        if args.e2e_multi:
            W.adapter_alignment_batch_multi([(hb.numpy(), ho.numpy(), abuf, aoff, hout.numpy()) for hb, ho, abuf, aoff, hout in host],
                                            scoring)
            return
        for hb, ho, abuf, aoff, hout in host:
            W.adapter_alignment_batch(hb.numpy(), ho.numpy(), abuf, aoff, scoring, out=hout.numpy())

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- value: device-resident ----
    for _ in range(max(Wm, 3)):
        step_device()
    barrier()
    W.synchronize()
    W.timing_enable(True)
    W.timing_read(reset=True)
    l0 = W.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active = True
    barrier()
    ev0.record(bench_stream)
    for _ in range(K):
        step_device()
    ev1.record(bench_stream)
    barrier()
    sampler.active = False
    launches = W.kernel_launches() - l0
    dev_ms = ev0.elapsed_time(ev1)
    W.synchronize()
    dp_ms, dp_n = W.timing_read(reset=True)
    W.timing_enable(False)

    # ---- e2e: host buffers through the C-ABI ----
    for _ in range(max(Wm, 3)):
        step_e2e()
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.active = False
    sampler.stop_flag = True
    if args.e2e_decisions:
        # the in-run parity gate needs the records themselves: one more (untimed) step through the record call, and the
        # decisions of the timed steps are checked against the host rule on those records
        from porechop_b200 import hostio
        out_bytes = out_bytes_dec
        dec_bad = 0
        for b, (hb, ho, abuf, aoff, hout), (t, p) in zip(batches, host, dec_out):
            W.adapter_alignment_batch(hb.numpy(), ho.numpy(), abuf, aoff, scoring, out=hout.numpy())
            rec = hout.numpy().reshape(ho.numel() - 1, len(b[2]), 9)
            exp = hostio.end_trim(rec, b[0] == 'start', wl.END_SIZE, 2, 75.0, 4)
            dec_bad += int(np.count_nonzero(exp != t.numpy()))
            dec_bad += int(np.count_nonzero(p.numpy()[:, :, 0] != rec[:, :, 7].astype(np.uint16)))

    # informational: NCCL re-gather of the 36-byte records of every rank to rank 0 (device -> device over NVLink)
    gather_ms = None
    if world > 1:
        db, do, abuf, aoff, dout, max_len = dev[0]
        bucket = [torch.empty_like(dout) for _ in range(world)] if rank == 0 else None
        dist.gather(dout, bucket, dst=0)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _, _, _, _, d_o, _ in dev:
            dist.gather(d_o, [torch.empty_like(d_o) for _ in range(world)] if rank == 0 else None, dst=0)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)

    t = torch.tensor([dev_ms, e2e_s * 1e3, gather_ms or 0.0], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, gather_ms = float(t[0]), float(t[1]), (float(t[2]) if world > 1 else None)
    total_reads = n * world
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = total_reads * K / (dev_ms / 1e3)
    e2e_value = total_reads * K / (e2e_ms / 1e3)
    # roofline of the dominant kernel (trace_kernel / score_kernel launches timed by CUDA events in the library)
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback'
    alg_bytes_step = sum(batch_alg_bytes(b) for b in batches)
    dp_launch_ms = dp_ms / max(dp_n, 1)
    launches_per_step = dp_n / K
    bytes_per_launch = alg_bytes_step / max(launches_per_step, 1e-9)
    achieved = bytes_per_launch / (dp_launch_ms / 1e3) / 1e9 if dp_n else None
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    clocks = sampler.summary()
    sm_mhz = clocks.get('sm_mhz') or 1965.0
    gcups_kernel = cells_per_step * K / (dp_ms / 1e3) / 1e9 if dp_n else None
    line = {
        'metric': 'reads/sec', 'value': value, 'unit': 'reads/s', 'n_gpus': world, 'steps': K, 'warmup': Wm,
        'ms_per_step': dev_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'int16 (s16x2 DPX)', 'data': 'synthetic',
        'gcups': cells_per_step * world * K / (dev_ms / 1e3) / 1e9,
        'config': {'workload': desc, 'reads_per_gpu': n, 'alignments_per_read': sum(len(b[2]) for b in batches),
                   'cells_per_read': cells_per_step / n, 'scoring': list(scoring),
                   'l2': 'inputs %.0f MB per step exceed the 126 MB L2' % (in_bytes / 1e6)},
        'e2e': {'value': e2e_value, 'unit': 'reads/s', 'h2d_bytes_per_step': in_bytes, 'd2h_bytes_per_step': out_bytes,
                'ms_per_step': e2e_ms / K, 'gcups': cells_per_step * world * K / (e2e_ms / 1e3) / 1e9,
                'path': ('adapterEndDecisions (host buffers, pinned), one submit per step per rank, decisions come back' if args.e2e_decisions else
                         'adapterAlignmentBatchMulti (host buffers, pinned), one submit per step per rank' if args.e2e_multi else
                         'adapterAlignmentBatch (host buffers, pinned), one call per batch per rank') + '; records stay rank-local',
                'options': args.opt,
                'decisions': ({'mismatches_vs_host_rule': dec_bad} if args.e2e_decisions else None),
                'gather_records_to_rank0_ms': gather_ms},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': (achieved / peak) if achieved else None, 'traffic': traffic, 'peak_source': peak_src,
                     'kernel': 'trace_kernel (s16x2 wavefront DP + in-kernel traceback)', 'launch_ms': dp_launch_ms,
                     'launches_per_step': launches_per_step, 'algorithmic_bytes_per_launch': bytes_per_launch,
                     'note': 'the kernel is integer-ALU (DPX) bound by construction (SURVEY 0.7); see alu'},
        'alu': {'gcups_kernel': gcups_kernel, 'sm_mhz': sm_mhz,
                'lane_instr_peak_per_s': 148 * 4 * 32 * sm_mhz * 1e6,
                'cells_per_lane_instr': (gcups_kernel * 1e9) / (148 * 4 * 32 * sm_mhz * 1e6) if gcups_kernel else None},
        'clocks': clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        per_read_cells = cells_per_step / n
        sample = args.cpu_sample_reads or int(max(64, min(n, 0.08e9 * 1.0 * cores * 4 / per_read_cells)))
        answers = []
        sec, reads, ccells, kind = run_reference_harness(batches, scoring, sample, cores, answers)
        try:    # parity gate on the same sample: records of the last e2e step vs the CPU reference's strings
            line['parity'] = parity_gate([h[4].numpy() for h in host], answers, W.format_record)
            line['parity']['against'] = kind
        except Exception as e:      # never lose the bench line to the checker
            line['parity'] = {'error': repr(e)}
        line['cpu_baseline'] = {'value': sample / sec, 'unit': 'reads/s', 'cores': cores, 'kind': kind,
                                'gcups': ccells / sec / 1e9,
                                'sample': 'first %d reads of the same batch (all their alignments), %d threads, native harness '
                                          'over the reference C-ABI (adapterAlignment+freeCString)' % (sample, cores)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
