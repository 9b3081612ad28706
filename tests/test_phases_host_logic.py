"""CPU tier: the host-side logic of the batched phase drivers (gather -> submit -> scatter, speculative masking rounds)
with the engine call replaced by the oracle -- so the scatter/ordering/masking logic is checked against the reference's
own Python outputs (tests/golden/golden_phases.json) even without a GPU.  The GPU tier runs the same comparison through
the real engine (tests/test_gpu_phases.py)."""
import pytest

from helpers import load_golden, oracle_batch
from test_gpu_phases import SC, AdapterSet, Read, feq


@pytest.fixture
def oracle_engine(monkeypatch):
    from porechop_b200 import phases

    def fake(seq_buf, seq_off, ad_buf, ad_off, scoring, pair_seq=None, pair_adapter=None, out=None):
        return oracle_batch(seq_buf, seq_off, ad_buf, ad_off, scoring, pair_seq, pair_adapter)
    monkeypatch.setattr(phases.W, 'adapter_alignment_batch', fake)
    return phases


@pytest.mark.parametrize('case_index', [0, 2, 3])
def test_phase_drivers_scatter_logic(oracle_engine, case_index):
    phases = oracle_engine
    case = load_golden('golden_phases.json')[case_index]
    ad = load_golden('adapters.json')
    reads = [Read(r['name'], r['seq']) for r in load_golden('fixture_reads.json') if r['file'] == case['file']]
    table = [AdapterSet(d) for d in ad['sets']]
    by_name = {s.name: s for s in table}
    by_name.update({d['name']: AdapterSet(d) for d in ad['full_barcode_sets']})
    # Phase A on a subset of the table (the full 236-sequence search is exercised on the GPU tier)
    subset = [by_name[n] for n in ('SQK-NSK007', 'SQK-MAP006', 'Rapid', 'Barcode 1 (reverse)', 'Barcode 2 (forward)')]
    phases.align_adapter_sets(reads, subset, 150, SC)
    exp = {e[0]: e for e in case['set_scores']}
    for s in subset:
        assert feq(s.best_start_score, exp[s.name][1]) and feq(s.best_end_score, exp[s.name][2])
    matching = [by_name[n] for n in case['matching_sets']]
    phases.find_adapters_at_read_ends(reads, matching, 150, 2, 75.0, SC, 4, case['barcodes'], case['forward_or_reverse'])
    adapters = []
    for m in matching:
        if m.start_sequence:
            adapters.append(tuple(m.start_sequence))
        if m.end_sequence and ((not m.start_sequence) or m.end_sequence[1] != m.start_sequence[1]):
            adapters.append(tuple(m.end_sequence))
    phases.find_adapters_in_read_middles(reads, adapters, case['middle_threshold'], 10, 100, SC,
                                         {m.start_sequence[0] for m in matching if m.start_sequence},
                                         {m.end_sequence[0] for m in matching if m.end_sequence})
    for r, e in zip(reads, case['reads']):
        assert (r.start_trim_amount, r.end_trim_amount) == (e['start_trim_amount'], e['end_trim_amount'])
        assert [[k, v] for k, v in r.start_barcode_scores.items()] == e['start_barcode_scores']
        assert [[k, v] for k, v in r.end_barcode_scores.items()] == e['end_barcode_scores']
        assert sorted(r.middle_adapter_positions) == e['middle_adapter_positions']
        assert sorted(r.middle_trim_positions) == e['middle_trim_positions']
        assert r.middle_hit_str == e['middle_hit_str']
