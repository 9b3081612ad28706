"""Long-running differential fuzz (not collected by pytest; authoring container only): random FASTQ / FASTA inputs and
random Porechop options through the UNMODIFIED reference CLI and through `porechop_b200.flat_cli` (oracle as the
engine), comparing every output file byte for byte.

    python tests/fuzz/fuzz_flat_cli.py <seed> <iterations>

Round 1: see DESIGN.md section 2 for the counts."""
import json
import os
import random
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np  # noqa: E402

from helpers import oracle_batch  # noqa: E402
from test_flat_cli import _flat_cli, _plain  # noqa: E402
from test_patch_cli import REF, _run_cli, load_reference  # noqa: E402

porechop, P, A = load_reference()
sys.path.insert(0, REF)
from porechop_b200 import cpp_function_wrappers as W  # noqa: E402

W.adapter_alignment_batch = lambda sb, so, ab, ao, sc, ps=None, pa=None, out=None: oracle_batch(
    np.asarray(sb), np.asarray(so), np.asarray(ab), np.asarray(ao), list(sc), ps, pa)

AD = json.load(open(os.path.join(os.path.dirname(HERE), 'golden', 'adapters.json')))
SETS = {d['name']: d for d in AD['sets']}
FULL = {d['name']: d for d in AD['full_barcode_sets']}


def mutate(rng, s, p):
    out = []
    for c in s:
        r = rng.random()
        if r < p * 0.3:
            continue
        if r < p * 0.7:
            out.append(rng.choice('ACGT'))
            continue
        out.append(c)
        if r > 1 - p * 0.3:
            out.append(rng.choice('ACGT'))
    return ''.join(out)


def make_input(rng, barcoded, fasta):
    def rand(n):
        return ''.join(rng.choice('ACGT') for _ in range(n))
    kit = rng.choice(['SQK-NSK007', 'SQK-MAP006', 'PCR adapters 1', 'SQK-NSK007'])
    start, end = SETS[kit]['start'], SETS[kit]['end']
    recs = []
    for i in range(rng.randint(3, 14)):
        body = rand(rng.choice([0, 10, 60, 149, 150, 151, 400, 1500, 3000]) + rng.randint(0, 30))
        p = rng.choice([0.0, 0.05, 0.12, 0.25])
        if barcoded:
            k = rng.randint(1, 6)
            nb, nb2 = FULL['Native barcoding %d (full sequence)' % k], FULL['Native barcoding %d (full sequence)' % rng.randint(1, 6)]
            s = (mutate(rng, nb['start'][1], p) if rng.random() < 0.8 else '') + body + \
                (mutate(rng, (nb if rng.random() < 0.8 else nb2)['end'][1], p) if rng.random() < 0.7 else '')
        else:
            s = (mutate(rng, start[1], p) if start and rng.random() < 0.8 else '') + body
            if rng.random() < 0.3 and start:
                s += mutate(rng, (end or start)[1], p) + mutate(rng, start[1], p) + rand(rng.randint(50, 1500))
            s += mutate(rng, end[1], p) if end and rng.random() < 0.6 else ''
        if rng.random() < 0.1:
            s = s.lower()
        if rng.random() < 0.05:
            s = s.replace('T', 'U')
        name = 'r%d' % i + (' some description' if rng.random() < 0.4 else '')
        if fasta:
            width = rng.choice([60, 70, 10000])
            recs.append('>' + name + '\n' + '\n'.join(s[k:k + width] for k in range(0, len(s), width)) + '\n')
        else:
            q = ''.join(chr(rng.randint(35, 73)) for _ in range(len(s)))
            recs.append('@' + name + '\n' + s + '\n+\n' + q + '\n')
    return ''.join(recs)


def random_args(rng, barcoded, out):
    a = ['-v', '0', '-t', '1', '--check_reads', str(rng.choice([2, 5, 10000]))]
    if barcoded:
        a += ['-b', out + '/bins']
        if rng.random() < 0.3:
            a += ['--require_two_barcodes']
        if rng.random() < 0.3:
            a += ['--discard_unassigned']
        if rng.random() < 0.3:
            a += ['--untrimmed']
        if rng.random() < 0.5:
            a += ['--barcode_threshold', str(rng.choice([60, 70, 80])), '--barcode_diff', str(rng.choice([0, 1, 5, 10]))]
    else:
        a += ['-o', out + '/o.' + rng.choice(['fastq', 'fasta', 'fastq.gz'])]
    if rng.random() < 0.5:
        a += ['--end_size', str(rng.choice([50, 100, 150, 200]))]
    if rng.random() < 0.5:
        a += ['--end_threshold', str(rng.choice([60, 75, 85])), '--min_trim_size', str(rng.choice([1, 4, 10])),
              '--extra_end_trim', str(rng.choice([0, 2, 7]))]
    if rng.random() < 0.6:
        a += ['--middle_threshold', str(rng.choice([70, 85, 95])), '--min_split_read_size', str(rng.choice([1, 100, 1000])),
              '--extra_middle_trim_good_side', str(rng.choice([0, 10, 50])), '--extra_middle_trim_bad_side', str(rng.choice([0, 100]))]
    if rng.random() < 0.2:
        a += ['--no_split']
    if rng.random() < 0.2:
        a += ['--discard_middle']
    if rng.random() < 0.3:
        a += ['--adapter_threshold', str(rng.choice([70, 90, 99]))]
    return a


def main():
    seed, iters = int(sys.argv[1]), int(sys.argv[2])
    rng = random.Random(seed)
    bad = 0
    for it in range(iters):
        barcoded, fasta = rng.random() < 0.4, rng.random() < 0.25
        text = make_input(rng, barcoded, fasta)
        d = tempfile.mkdtemp()
        try:
            inp = os.path.join(d, 'in.' + ('fasta' if fasta else 'fastq'))
            with open(inp, 'w') as f:
                f.write(text)
            os.environ['PB200_FLAT_CHUNK_BYTES'] = str(rng.choice([2000, 50000, 1 << 28]))
            st = rng.getstate()
            res = []
            for which in ('a', 'b'):
                rng.setstate(st)
                args = ['-i', inp] + random_args(rng, barcoded, os.path.join(d, which))
                try:
                    if which == 'a':
                        res.append(_plain(_run_cli(P, A, args, os.path.join(d, which))[1]))
                    else:
                        res.append(_plain(_flat_cli(A, args, os.path.join(d, which))))
                except SystemExit as e:        # e.g. "no barcodes were found": both must exit the same way
                    res.append(('exit', str(e)))
            if res[0] != res[1]:
                bad += 1
                keep = os.path.join(tempfile.gettempdir(), 'flat_fuzz_fail_%d_%d' % (seed, it))
                shutil.copytree(d, keep, dirs_exist_ok=True)
                print('MISMATCH', seed, it, args, 'kept in', keep, flush=True)
                if bad > 3:
                    break
        finally:
            shutil.rmtree(d, ignore_errors=True)
    print('seed', seed, 'iters', iters, 'bad', bad, flush=True)


if __name__ == '__main__':
    main()
