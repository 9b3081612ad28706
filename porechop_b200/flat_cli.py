"""
`python -m porechop_b200.flat_cli <porechop arguments>`: Porechop's run on flat buffers (no per-read Python objects).

The argument parser, the adapter table and every policy decision are the user's Porechop checkout's own code, called
as is (`porechop.porechop.get_arguments`, `ADAPTERS`, `fix_up_1d2_sets`, `choose_barcoding_kit`,
`add_full_barcode_adapter_sets`; porechop.py:33-79).  What is replaced is the data path between them
(porechop_b200/fastq.py): FASTQ bytes -> flat buffers -> batched alignments on the engine -> trim / split / barcode
decisions on arrays -> output bytes.  Output files are byte-identical to the reference CLI's
(tests/test_flat_cli.py); the progress / verbose report is not reproduced -- use `python -m porechop_b200`
(porechop_b200/patch.py) when that is wanted.

Limits (the run exits with a message instead of guessing): input must be one FASTQ file (plain or .gz); FASTA input and
Albacore directories go through `python -m porechop_b200`.
"""
import gzip
import os
import sys

from . import fastq


def _read_input(path):
    with open(path, 'rb') as f:
        magic = f.read(2)
    opener = gzip.open if magic == b'\x1f\x8b' else open
    with opener(path, 'rb') as f:
        data = f.read()
    # Python's text mode (misc.py:157 `open(..., 'rt')`) turns a lone '\r' into a line break as well; not supported here
    first = data.lstrip()[:1]
    if first != b'@':
        sys.exit('porechop_b200.flat_cli: input is not FASTQ (FASTA / directories: use `python -m porechop_b200`)')
    return data


def _out_format(args, read_type='fastq'):
    # the format rules of output_reads (porechop.py:624-650)
    fmt = args.format
    if fmt == 'auto':
        if args.output is None:
            fmt = read_type
            if args.barcode_dir is not None and args.input.lower().endswith('.gz'):
                fmt += '.gz'
        else:
            low = args.output.lower()
            fmt = next((f for f in ('fasta.gz', 'fastq.gz', 'fasta', 'fastq') if '.' + f in low), read_type)
    gz = fmt.endswith('.gz') and (args.barcode_dir is not None or args.output is not None)
    return (fmt[:-3] if fmt.endswith('.gz') else fmt), gz


def _write(path, payload, gz):
    if gz:
        with gzip.open(path, 'wb') as f:        # same bytes inside as the reference's `gzip -c`, different container timestamp
            f.write(payload)
    else:
        with open(path, 'wb') as f:
            f.write(payload)


def main():
    try:
        from porechop import porechop as P
    except ImportError:
        sys.exit('porechop_b200.flat_cli: no `porechop` package on sys.path (point PYTHONPATH at a Porechop checkout)')
    args = P.get_arguments()
    if os.path.isdir(args.input):
        sys.exit('porechop_b200.flat_cli: directory input is not supported (use `python -m porechop_b200`)')
    scoring = args.scoring_scheme_vals
    batch = fastq.parse_fastq(_read_input(args.input))

    # Phase A on flat buffers; the scores land on Porechop's own Adapter objects so that its policy code runs unchanged
    search = [a for a in P.ADAPTERS if '(full sequence)' not in a.name]                     # porechop.py:296
    as_tuple = lambda a: (a.name, tuple(a.start_sequence) or None, tuple(a.end_sequence) or None)   # noqa: E731
    best_s, best_e = fastq.search_adapter_sets(batch, [as_tuple(a) for a in search], scoring, args.check_reads, args.end_size)
    for a, s, e in zip(search, best_s, best_e):
        a.best_start_score, a.best_end_score = max(a.best_start_score, float(s)), max(a.best_end_score, float(e))
    matching = [a for a in search if a.best_start_or_end_score() >= args.adapter_threshold]  # porechop.py:327
    matching = P.fix_up_1d2_sets(matching)
    null = open(os.devnull, 'w')
    direction = P.choose_barcoding_kit(matching, 0, null) if args.barcode_dir else None
    matching = P.add_full_barcode_adapter_sets(matching)
    sets = [as_tuple(a) for a in matching]

    fmt, gz = _out_format(args)
    common = dict(end_size=args.end_size, extra_end_trim=args.extra_end_trim, end_threshold=args.end_threshold,
                  min_trim_size=args.min_trim_size, no_split=args.no_split, middle_threshold=args.middle_threshold,
                  extra_middle_trim_good_side=args.extra_middle_trim_good_side,
                  extra_middle_trim_bad_side=args.extra_middle_trim_bad_side,
                  min_split_read_size=args.min_split_read_size, discard_middle=args.discard_middle, fmt=fmt, as_array=True)
    if args.barcode_dir is not None:
        os.makedirs(args.barcode_dir, exist_ok=True)
        if sets:
            bins, _ = fastq.demux_fastq(batch, sets, scoring, direction, barcode_threshold=args.barcode_threshold,
                                        barcode_diff=args.barcode_diff, require_two_barcodes=args.require_two_barcodes,
                                        discard_unassigned=args.discard_unassigned, untrimmed=args.untrimmed, **common)
        else:
            out = fastq.emit(batch, fmt=fmt, untrimmed=args.untrimmed, as_array=True)
            bins = {} if (args.discard_unassigned or not len(out)) else {'none': out}
        for name, payload in bins.items():
            _write(os.path.join(args.barcode_dir, name + '.' + fmt + ('.gz' if gz else '')), payload, gz)
    else:
        if sets:
            out, _ = fastq.trim_fastq(batch, sets, scoring, **common)
        else:                       # "No adapters found - output reads are unchanged from input reads"
            out = fastq.emit(batch, fmt=fmt, as_array=True)
        if args.output is None:
            sys.stdout.buffer.write(out)
        else:
            _write(args.output, out, gz)
    return 0


if __name__ == '__main__':
    sys.exit(main())
