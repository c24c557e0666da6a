#!/usr/bin/env python3
"""Generate the golden vectors in tests/golden/ by running the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference; imports it through oracle/ref_shim):

    python tests/golden/make_golden.py

For every case it writes real FASTQ files, calls the reference's own entry points
  kbbq.recalibrate.fastq_to_covariate_arrays   (kbbq/recalibrate.py:22-121)
  kbbq.gatk.applybqsr.get_delta_qs             (kbbq/gatk/applybqsr.py:80-103)
  kbbq.recalibrate.recalibrate_fastq           (kbbq/recalibrate.py:123-156, stdout captured)
and stores inputs (packed SoA arrays + read names) and outputs in one .npz per case.
`delta_grid.npz` holds kbbq.compare_reads.gatk_delta_q (:235-260) on a grid of random cells.
The files are small fixtures; the reference itself never travels to the GPU box.
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

rc, cr, ab = ref_shim.load()
BASES = np.frombuffer(b"ATGC", dtype=np.uint8)


def make_case(name, N, L, R, seed, paired=True, infer_rg=True, p_n=0.01, p_err=0.03,
              qmode="uniform", all_second=False, lowq_reads=0):
    rng = np.random.default_rng(seed)
    seq = BASES[rng.integers(0, 4, size=(N, L))]
    if qmode == "uniform":
        qual = rng.integers(0, 43, size=(N, L)).astype(np.uint8)
    else:  # Illumina-shaped: high start, decaying, noisy, some Q2 tails
        mu = rng.normal(36, 3, size=(N, 1))
        cyc = (np.arange(L) / L) ** 2
        qual = np.clip(np.rint(mu - 8 * cyc + rng.normal(0, 2, size=(N, L))), 2, 41).astype(np.uint8)
        tails = rng.random(N) < 0.1
        for r in np.nonzero(tails)[0]:
            qual[r, L - rng.integers(1, max(2, L // 4)):] = 2
    isn = rng.random((N, L)) < p_n
    seq = np.where(isn, np.uint8(ord("N")), seq)
    qual = np.where(isn, np.uint8(2) if qmode != "uniform" else qual, qual).astype(np.uint8)
    for r in range(lowq_reads):  # whole reads below minscore
        qual[(7 * r + 3) % N, :] = rng.integers(0, 6, size=L)
    corr = seq.copy()
    e = rng.random((N, L)) < p_err
    shift = rng.integers(1, 4, size=(N, L))
    code = np.zeros(256, int)
    code[BASES] = np.arange(4)
    corr[e] = BASES[(code[seq[e]] + shift[e]) % 4]  # N -> some base also counts as an error
    if paired:
        second = (np.arange(N) % 2).astype(np.uint8)
        rgk = rng.integers(0, R, size=(N + 1) // 2).repeat(2)[:N]
    else:
        second = np.zeros(N, np.uint8)
        rgk = rng.integers(0, R, size=N)
    if all_second:
        second[:] = 1
    if infer_rg:
        names = ["r%d/%d_RG:Z:g%d" % (i // 2 if paired else i, second[i] + 1, rgk[i]) for i in range(N)]
    else:
        rgk[:] = 0
        names = ["r%d/%d" % (i // 2 if paired else i, second[i] + 1) for i in range(N)]
    # first-seen order -> rg ints, exactly what recalibrate.py:59-64 does
    seen = {}
    rg = np.array([seen.setdefault(k, len(seen)) for k in rgk], dtype=np.uint16)
    nrg = len(seen)

    with tempfile.TemporaryDirectory() as td:
        fu, fc = os.path.join(td, "u.fq"), os.path.join(td, "c.fq")
        with open(fu, "w") as hu, open(fc, "w") as hc:
            for i in range(N):
                q = (qual[i] + 33).tobytes().decode()
                hu.write("@%s\n%s\n+\n%s\n" % (names[i], seq[i].tobytes().decode(), q))
                hc.write("@%s corrected by lighter\n%s\n+\n%s\n" % (names[i], corr[i].tobytes().decode(), q))
        tables = rc.fastq_to_covariate_arrays((fu, fc), infer_rg=infer_rg)
        dqs = ab.get_delta_qs(*tables)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            rc.recalibrate_fastq((fu, fc), infer_rg=infer_rg)
        with open(fu) as hu:
            fastq_in = hu.read()
        with open(fc) as hc:
            fastq_corr = hc.read()
    lines = buf.getvalue().split("\n")
    outq = np.array([[ord(ch) - 33 for ch in lines[4 * i + 3]] for i in range(N)], dtype=np.int16)
    assert all(lines[4 * i] == "@" + names[i] for i in range(N))
    keys = ("meanq", "rg_errs", "rg_total", "q_errs", "q_total", "pos_errs", "pos_total",
            "dinuc_errs", "dinuc_total")
    out = dict(seq=seq, qual=qual, corr=corr, rg=rg, second=second, L=L, R=nrg,
               infer_rg=infer_rg, names=np.array("\n".join(names)), outq=outq,
               rgdq=dqs[0], qdq=dqs[1], posdq=dqs[2], dindq=dqs[3])
    out.update({k: np.asarray(v, dtype=np.int64) for k, v in zip(keys, tables)})
    if N <= 64:
        out.update(fastq_in=np.array(fastq_in), fastq_corr=np.array(fastq_corr),
                   fastq_out=np.array(buf.getvalue()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "N", N, "L", L, "R", nrg, "meanq", tables[0], "out range", outq.min(), outq.max(),
          "changed", float((outq != qual).mean()))


def make_synth_case(name, seed, N, L, R):
    """BASELINE-shaped case: the reads come from kbbq.synth (the numpy twin of the device generator bench.py and the
    GPU tests use), so the .npz only keeps (seed, N, L, R), a digest of the inputs, the first-seen read-group
    numbers and the reference's outputs; tests regenerate the same bytes."""
    import hashlib
    import importlib.util   # by path: the name `kbbq` is taken by the reference package here
    spec = importlib.util.spec_from_file_location("kbbq_b200_synth", os.path.join(ROOT, "kbbq-py_b200", "kbbq", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    seq, qual, corr, rgk, second = synth.synth_reads(seed, 0, N, L, R)
    infer_rg = R > 1
    names = ["r%d/%d" % (i // 2, second[i] + 1) + ("_RG:Z:g%d" % rgk[i] if infer_rg else "") for i in range(N)]
    seen = {}
    rg = np.array([seen.setdefault(int(k), len(seen)) for k in rgk], dtype=np.uint16) if infer_rg else np.zeros(N, np.uint16)
    nrg = max(1, len(seen))
    with tempfile.TemporaryDirectory() as td:
        fu, fc = os.path.join(td, "u.fq"), os.path.join(td, "c.fq")
        with open(fu, "w") as hu, open(fc, "w") as hc:
            for i in range(N):
                q = (qual[i] + 33).tobytes().decode()
                hu.write("@%s\n%s\n+\n%s\n" % (names[i], seq[i].tobytes().decode(), q))
                hc.write("@%s\n%s\n+\n%s\n" % (names[i], corr[i].tobytes().decode(), q))
        tables = rc.fastq_to_covariate_arrays((fu, fc), infer_rg=infer_rg)
        dqs = ab.get_delta_qs(*tables)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            rc.recalibrate_fastq((fu, fc), infer_rg=infer_rg)
    lines = buf.getvalue().split("\n")
    outq = np.array([[ord(ch) - 33 for ch in lines[4 * i + 3]] for i in range(N)], dtype=np.int16)
    assert all(lines[4 * i] == "@" + names[i] for i in range(N)) and outq.min() >= 0 and outq.max() < 256
    digest = hashlib.sha256(seq.tobytes() + qual.tobytes() + corr.tobytes() + rgk.tobytes() + second.tobytes()).hexdigest()
    keys = ("meanq", "rg_errs", "rg_total", "q_errs", "q_total", "pos_errs", "pos_total",
            "dinuc_errs", "dinuc_total")
    out = dict(seed=seed, N=N, L=L, R=nrg, R_synth=R, infer_rg=infer_rg, input_sha256=np.array(digest), rg=rg,
               outq=outq.astype(np.uint8), rgdq=dqs[0], qdq=dqs[1], posdq=dqs[2], dindq=dqs[3])
    out.update({k: np.asarray(v, dtype=np.int64) for k, v in zip(keys, tables)})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "N", N, "L", L, "R", nrg, "meanq", tables[0], "changed", float((outq != qual).mean()))


def delta_grid(seed=7, n=24000):
    rng = np.random.default_rng(seed)
    prior = rng.integers(0, 43, size=n)
    scale = 10 ** rng.uniform(0, 11, size=n)
    tot = (rng.random(n) * scale).astype(np.int64)
    rate = 10 ** rng.uniform(-5, 0, size=n)
    errs = np.minimum(tot, (tot * rate * rng.uniform(0.2, 5, size=n)).astype(np.int64))
    # hand-placed corners: empty cells, all errors, no errors, priors at both ends
    corner = np.array([[p, e, t] for p in (0, 1, 6, 20, 41, 42)
                       for (e, t) in ((0, 0), (0, 1), (1, 1), (0, 10 ** 6), (10 ** 6, 10 ** 6),
                                      (5, 10), (10 ** 9, 10 ** 11), (123456, 987654321))])
    prior = np.concatenate([prior, corner[:, 0]])
    errs = np.concatenate([errs, corner[:, 1]])
    tot = np.concatenate([tot, corner[:, 2]])
    dq = np.concatenate([cr.gatk_delta_q(prior[i:i + 4096], errs[i:i + 4096], tot[i:i + 4096])
                         for i in range(0, prior.size, 4096)])
    np.savez_compressed(os.path.join(HERE, "delta_grid.npz"), prior=prior, errs=errs, total=tot, dq=dq)
    print("delta_grid", prior.size, "dq range", dq.min(), dq.max())


def scalar_kats():
    q = np.arange(43)
    p = cr.q_to_p(q)
    rt = cr.p_to_q(p)
    ps = np.concatenate([[0.0, 1.0, 0.5, 1e-5, 0.0999999, 0.1000001, 2.0], 10 ** -(np.linspace(0, 5, 101) + 0.003)])
    np.savez_compressed(os.path.join(HERE, "scalars.npz"), q_to_p=p.astype(np.float64),
                        p_to_q_roundtrip=rt, p_samples=ps, p_to_q_samples=cr.p_to_q(ps),
                        prior_dist=np.asarray(cr.RescaledNormal.prior_dist, dtype=np.float64))
    print("scalars ok", rt)


if __name__ == "__main__":
    make_case("mixed_r3", N=600, L=60, R=3, seed=11)
    make_case("illumina_r1", N=800, L=75, R=1, seed=12, infer_rg=False, qmode="illumina", p_err=0.01, p_n=0.002)
    make_case("tails_r2_second", N=300, L=40, R=2, seed=13, qmode="illumina", all_second=True, lowq_reads=9)
    make_case("odd_len_unpaired", N=257, L=33, R=4, seed=14, paired=False)
    make_case("tiny_r2", N=24, L=17, R=2, seed=15)
    if "--no-synth" not in sys.argv:   # the BASELINE shapes (configs 2, 3, 4), ~1 minute each
        make_synth_case("c2_l150_r1", 1002, 5000, 150, 1)
        make_synth_case("c3_l150_r8", 1003, 5000, 150, 8)
        make_synth_case("c4_l250_r32", 1004, 4000, 250, 32)
    delta_grid()
    scalar_kats()
