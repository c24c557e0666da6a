"""Pins the CPU oracle (oracle/kbbq_oracle.c) to the reference: its own known-answer tests and
the golden vectors produced by running the unmodified Python reference (tests/golden/make_golden.py).
Runs without a GPU."""
import numpy as np

from conftest import DELTA_KEYS, TABLE_KEYS, load_case


def test_oracle_tables_deltas_apply_match_reference(golden_case, oracle_mod):
    g = golden_case
    L, R = int(g["L"]), int(g["R"])
    t = oracle_mod.covariate_arrays(g["seq"], g["qual"], g["corr"], g["rg"], g["second"], L, R)
    for got, key in zip(t, TABLE_KEYS):
        assert np.array_equal(got, g[key]), key
    d = oracle_mod.get_delta_qs(*t)
    for got, key in zip(d, DELTA_KEYS):
        assert np.array_equal(got, g[key]), key
    out = oracle_mod.apply(g["seq"], g["qual"], g["rg"], g["second"], L, R, t[0], *d)
    assert np.array_equal(out, g["outq"])
    out2 = oracle_mod.recalibrate(g["seq"], g["qual"], g["corr"], g["rg"], g["second"], L, R, threads=3)
    assert np.array_equal(out2, g["outq"])


def test_oracle_matches_reference_at_baseline_shapes(synth_golden_case, oracle_mod):
    """BASELINE configs 2, 3, 4 (150 bp x 1 and 8 read groups, 250 bp x 32) at 4-5 k reads: the unmodified reference's
    tables, deltas and output qualities on the reads kbbq.synth generates (tests/golden/make_golden.py)."""
    g = synth_golden_case
    L, R = int(g["L"]), int(g["R"])
    t = oracle_mod.covariate_arrays(g["seq"], g["qual"], g["corr"], g["rg"], g["second"], L, R)
    for got, key in zip(t, TABLE_KEYS):
        assert np.array_equal(got, g[key]), key
    d = oracle_mod.get_delta_qs(*t)
    for got, key in zip(d, DELTA_KEYS):
        assert np.array_equal(got, g[key]), key
    assert np.array_equal(oracle_mod.apply(g["seq"], g["qual"], g["rg"], g["second"], L, R, t[0], *d), g["outq"])


def test_oracle_delta_grid(oracle_mod):
    g = load_case("delta_grid")
    dq = oracle_mod.gatk_delta_q(g["prior"], g["errs"], g["total"])
    assert np.array_equal(dq, g["dq"])


def test_oracle_near_ties_match_reference(oracle_mod):
    """Cells constructed so that the two best candidates nearly tie (oracle.near_tie_cells), answered by the
    unmodified reference (tests/golden/make_golden_ties.py): pins `first maximum of the long-double sums`."""
    g = load_case("delta_near_ties")
    pq, errs, tot = g["prior"].astype(np.int64), g["errs"], g["total"]
    assert np.array_equal(oracle_mod.gatk_delta_q(pq, errs, tot), g["dq"].astype(np.int64))
    # the generator is reproducible: the stored cells are what near_tie_cells constructs
    p2, e2, t2 = oracle_mod.near_tie_cells(75_000, 31, 3.0, 10.0)
    assert np.array_equal(p2, pq) and np.array_equal(e2, errs) and np.array_equal(t2, tot)
    gap = oracle_mod.delta_q_top2_gap(pq, errs, tot)
    assert (gap < 2.0 ** -40).sum() >= 20      # a blind random grid has none
    # documentation of the one known difference: at >= 1e10 observations per cell scipy's
    # candidate-independent gammaln term, added before rounding, can reorder two candidates that tie in the restatement
    assert len(g["known_diff_prior"]) <= 10 and int(g["known_diff_searched"]) >= 2_000_000
    assert (g["known_diff_total"] >= 10 ** 10).all() and (g["known_diff_gap"] < 2.0 ** -50).all()


def test_oracle_reference_kats(oracle_mod):
    # tests/test_compare_reads.py:141-151 (+ the survey's probed values)
    dq = oracle_mod.gatk_delta_q([10, 20, 30], [10, 200, 0], [1000, 1000, 50000])
    assert dq.tolist() == [3, -8, 2]
    # tests/test_gatk_applybqsr.py:105-121
    one = np.array
    d = oracle_mod.get_delta_qs(one([10]), one([0]), one([1000]), one([[0] * 43]), one([[1000] * 43]),
                                np.zeros((1, 43, 2), int), np.full((1, 43, 2), 1000),
                                np.zeros((1, 43, 16), int), np.full((1, 43, 16), 1000))
    assert d[0].tolist() == [3] and d[1][0, 0] == 2 and d[2][0, 0, 0] == 1
    assert d[3][0, 0, 0] == 1 and d[3][0, 0, 16] == 0
    # tests/test_compare_reads.py:153-161
    assert oracle_mod.p_to_q([.2, .3, .4, .1, .01, .001]).tolist() == [6, 5, 3, 10, 20, 30]
    # tests/test_recalibrate.py:19-99: ATG / ACG, quals 7 7 2
    seq = np.frombuffer(b"ATG", np.uint8)
    corr = np.frombuffer(b"ACG", np.uint8)
    qual = np.array([7, 7, 2], np.uint8)
    t = oracle_mod.covariate_arrays(seq, qual, corr, None, None, 3, 1)
    assert t[0].tolist() == [6] and t[1].tolist() == [1] and t[2].tolist() == [2]
    assert t[3][0, 7] == 1 and t[4][0, 7] == 2 and t[3].sum() == 1 and t[4].sum() == 2
    assert t[5].shape == (1, 43, 6) and t[5][0, 7, 1] == 1 and t[5].sum() == 1
    assert t[6][0, 7, 0] == 1 and t[6][0, 7, 1] == 1 and t[6].sum() == 2
    assert t[7][0, 7, 1] == 1 and t[7].sum() == 1 and t[8][0, 7, 1] == 1 and t[8].sum() == 1  # 'AT' = 1
    out = oracle_mod.recalibrate(seq, qual, corr, None, None, 3, 1)
    assert out.tolist() == [[6, 6, 2]]


def test_oracle_constants(oracle_mod):
    s = load_case("scalars")
    p, lnp, ln1mp, prior = oracle_mod.constants()
    assert np.array_equal(p, s["q_to_p"])
    assert np.array_equal(prior, s["prior_dist"])
    # header == the reference's expressions evaluated by THIS numpy
    with np.errstate(divide="ignore"):
        assert np.array_equal(lnp, np.log(np.power(10.0, -(np.arange(43) / 10.0))))
        assert np.array_equal(ln1mp, np.log1p(-np.power(10.0, -(np.arange(43) / 10.0))))
    assert np.array_equal(oracle_mod.p_to_q(s["p_samples"]), s["p_to_q_samples"])
    # the meanq path (long double) reproduces the reference's truncation artefact q -> p -> q
    for q in range(6, 43):
        pt = np.zeros((1, 43, 2), np.int64)
        pt[0, q, 0] = 5
        mq = oracle_mod.marginals(np.zeros_like(pt), pt)[0]
        assert mq[0] == s["p_to_q_roundtrip"][q]


def test_oracle_input_errors(oracle_mod):
    import pytest
    seq = np.frombuffer(b"ACGTACGT", np.uint8)
    with pytest.raises(oracle_mod.OracleError):
        oracle_mod.build_tables(seq, np.full(8, 43, np.uint8), seq, None, None, 8, 1)
    with pytest.raises(oracle_mod.OracleError):
        oracle_mod.build_tables(np.frombuffer(b"ACGTXCGT", np.uint8), np.full(8, 30, np.uint8), seq, None, None, 8, 1)
