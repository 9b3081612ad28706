"""GPU tier: the batched Phase A/B/C drivers (porechop_b200/phases.py) against golden vectors produced by the
reference's own Python + C++ (tests/golden/make_golden_phases.py): adapter-set scores, trim amounts, recorded
alignments, barcode score dicts (incl. insertion order) and the sequentially-masked middle-adapter hits."""
import pytest

from helpers import load_golden

pytestmark = pytest.mark.gpu
SC = [3, -6, -5, -2]


class Read:
    """the NanoporeRead fields the drivers fill (reference nanopore_read.py:21-55)"""

    def __init__(self, name, seq):
        self.name, self.seq = name, seq.upper()
        self.start_trim_amount = self.end_trim_amount = 0
        self.start_adapter_alignments, self.end_adapter_alignments = [], []
        self.middle_adapter_positions, self.middle_trim_positions = set(), set()
        self.middle_hit_str = ''
        self.start_barcode_scores, self.end_barcode_scores = {}, {}


class AdapterSet:
    """the Adapter interface the drivers use (reference adapters.py:18-52)"""

    def __init__(self, d):
        self.name = d['name']
        self.start_sequence, self.end_sequence = d['start'], d['end']
        self.best_start_score = self.best_end_score = 0.0

    def best_start_or_end_score(self):
        return max(self.best_start_score, self.best_end_score)

    def is_barcode(self):
        return self.name.startswith('Barcode ')

    def barcode_direction(self):
        return 'reverse' if '_rev' in self.start_sequence[0] else 'forward'

    def get_barcode_name(self):
        names = [self.name] + ([self.start_sequence[0]] if self.start_sequence else []) + \
                ([self.end_sequence[0]] if self.end_sequence else [])
        return sorted(names, key=len)[0].replace(' ', '_')


def feq(a, b):
    return a == b or (a != a and b != b)


@pytest.mark.parametrize('case_index', [0, 1, 2, 3])
def test_phases_match_reference(case_index):
    from porechop_b200 import phases
    case = load_golden('golden_phases.json')[case_index]
    ad = load_golden('adapters.json')
    reads = [Read(r['name'], r['seq']) for r in load_golden('fixture_reads.json') if r['file'] == case['file']]
    table = [AdapterSet(d) for d in ad['sets']]
    full = {d['name']: AdapterSet(d) for d in ad['full_barcode_sets']}

    # Phase A
    phases.align_adapter_sets(reads, table, 150, SC)
    got = [[s.name, s.best_start_score, s.best_end_score] for s in table]
    for g, e in zip(got, case['set_scores']):
        assert g[0] == e[0] and feq(g[1], e[1]) and feq(g[2], e[2]), (g, e)
    matching = [s for s in table if s.best_start_or_end_score() >= 90.0]
    by_name = {s.name: s for s in table}
    by_name.update(full)
    matching = [by_name[n] for n in case['matching_sets']]        # incl. the synthesised full-barcode sets, in order
    assert set(s.name for s in matching if s.name in {t.name for t in table}) == \
        set(s.name for s in table if s.best_start_or_end_score() >= 90.0)

    # Phase B
    phases.find_adapters_at_read_ends(reads, matching, 150, 2, 75.0, SC, 4, case['barcodes'], case['forward_or_reverse'])
    # Phase C
    adapters = []
    for m in matching:
        if m.start_sequence:
            adapters.append(tuple(m.start_sequence))
        if m.end_sequence and ((not m.start_sequence) or m.end_sequence[1] != m.start_sequence[1]):
            adapters.append(tuple(m.end_sequence))
    start_names = {m.start_sequence[0] for m in matching if m.start_sequence}
    end_names = {m.end_sequence[0] for m in matching if m.end_sequence}
    phases.find_adapters_in_read_middles(reads, adapters, case['middle_threshold'], 10, 100, SC, start_names, end_names)

    for r, e in zip(reads, case['reads']):
        assert r.name == e['name']
        assert (r.start_trim_amount, r.end_trim_amount) == (e['start_trim_amount'], e['end_trim_amount']), r.name
        for got_list, exp_list in ((r.start_adapter_alignments, e['start_adapter_alignments']),
                                   (r.end_adapter_alignments, e['end_adapter_alignments'])):
            assert len(got_list) == len(exp_list)
            for g, x in zip(got_list, exp_list):
                assert g[0].name == x[0] and feq(g[1], x[1]) and feq(g[2], x[2]) and g[3] == x[3] and g[4] == x[4]
        assert [[k, v] for k, v in r.start_barcode_scores.items()] == e['start_barcode_scores']
        assert [[k, v] for k, v in r.end_barcode_scores.items()] == e['end_barcode_scores']
        assert sorted(r.middle_adapter_positions) == e['middle_adapter_positions']
        assert sorted(r.middle_trim_positions) == e['middle_trim_positions']
        assert r.middle_hit_str == e['middle_hit_str']
