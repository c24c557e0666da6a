"""Native FASTQ ingest / egress (csrc/fastq_io.cpp) against the Python tokenizer
that mirrors pysam.FastxFile (kbbq/fastx.py) and against the reference's naming rules
(kbbq/compare_reads.py:304-318, kbbq/recalibrate.py:17,59-64,152-156).  Host only: no GPU needed."""
import gzip
import os

import numpy as np
import pytest

from kbbq import fastx


def _write(path, records, eol="\n", final_newline=True):
    text = eol.join("@%s%s\n%s\n+\n%s".replace("\n", eol) % (n, (" " + c) if c else "", s, q) for n, c, s, q in records)
    if final_newline:
        text += eol
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wt", newline="") as fh:
        fh.write(text)


def _random_records(rng, n, L, n_rg=5, comments=True):
    recs = []
    for i in range(n):
        pair = i // 2
        name = "r%d/%d_RG:Z:g%d" % (pair, 1 + (i & 1), rng.integers(0, n_rg))
        seq = "".join(rng.choice(list("ACGTN"), size=L))
        qual = "".join(chr(33 + int(q)) for q in rng.integers(2, 42, size=L))
        recs.append((name, "comment %d" % i if comments and i % 3 == 0 else "", seq, qual))
    return recs


@pytest.mark.parametrize("eol,final,suffix", [("\n", True, ".fq"), ("\r\n", True, ".fq"), ("\n", False, ".fq"),
                                              ("\n", True, ".fq.gz")])
def test_pack_matches_python_tokenizer(tmp_path, eol, final, suffix):
    rng = np.random.default_rng(3)
    recs = _random_records(rng, 1237, 41)
    path = tmp_path / ("x" + suffix)
    _write(path, recs, eol, final)
    f = fastx.NativeFastq(path, threads=5)
    assert f.N == len(recs) and f.L == 41
    seq, qual = f.pack()
    for i, rec in enumerate(fastx.FastxFile(str(path))):
        assert rec.name == recs[i][0] == f.name(i)
        assert seq[i].tobytes().decode() == rec.sequence == recs[i][2]
        assert list(qual[i]) == rec.get_quality_array()
    part_s, part_q = f.pack(100, 50)
    assert np.array_equal(part_s, seq[100:150]) and np.array_equal(part_q, qual[100:150])
    f.close()


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_infer_rg_first_seen_order_and_second(tmp_path, threads):
    rng = np.random.default_rng(4)
    recs = _random_records(rng, 5000, 8, n_rg=7)
    path = tmp_path / "x.fq"
    _write(path, recs)
    f = fastx.NativeFastq(path, threads=threads)
    rg, second, keys = f.infer(True)
    seen = {}
    for i, (name, _, _, _) in enumerate(recs):  # kbbq/recalibrate.py:59-64 with fastq_infer_rg
        key = name.split("_")[1].split(":")[-1]
        assert rg[i] == seen.setdefault(key, len(seen))
        assert second[i] == (name.split("_")[0][-2:] == "/2")
    assert keys == list(seen)
    rg0, second0, keys0 = f.infer(False)
    assert not rg0.any() and keys0 == [0] and np.array_equal(second0, second)


def test_name_rules_raise_like_the_reference(tmp_path):
    a, b = tmp_path / "a.fq", tmp_path / "b.fq"
    _write(a, [("foo/1", "", "ACG", "III"), ("bar/2_XX:Z:g", "", "ACG", "III")])
    f = fastx.NativeFastq(a)
    with pytest.raises(IndexError):       # no second '_' field
        f.infer(True)
    _write(a, [("bar/2_XX:Z:g", "", "ACG", "III")])
    with pytest.raises(AssertionError):   # second field does not start with RG
        fastx.NativeFastq(a).infer(True)
    _write(a, [("foo/1", "", "ACG", "III"), ("bar/2", "", "ACG", "III")])
    _write(b, [("foo/1", "corrected", "ACG", "III"), ("baz/2", "", "ACG", "III")])
    with pytest.raises(AssertionError):   # find_corrected_sites name check
        fastx.NativeFastq(a).check_names(fastx.NativeFastq(b), 2)
    fastx.NativeFastq(a).check_names(fastx.NativeFastq(b), 1)


def test_malformed_and_ragged_inputs(tmp_path):
    p = tmp_path / "x.fq"
    p.write_text("@a\nACG\n+\nIII\n@b\nAC\n")
    with pytest.raises(ValueError):
        fastx.NativeFastq(p)
    p.write_text("@a\nACG\n+\nII\n")
    with pytest.raises(ValueError):
        fastx.NativeFastq(p)
    p.write_text("@a\nACG\n+\nIII\n@b\nAC\n+\nII\n")
    f = fastx.NativeFastq(p)
    assert f.N == 2 and f.L == -1
    with pytest.raises(ValueError):
        f.pack()
    p.write_text("")
    f = fastx.NativeFastq(p)
    assert f.N == 0
    names, seq, qual = fastx.read_packed(p)
    assert names == [] and seq.size == 0
    with pytest.raises(OSError):
        fastx.NativeFastq(tmp_path / "missing.fq")


def test_write_round_trip(tmp_path):
    rng = np.random.default_rng(6)
    recs = _random_records(rng, 70_001, 23)
    src, dst = tmp_path / "in.fq", tmp_path / "out.fq"
    _write(src, recs)
    f = fastx.NativeFastq(src, threads=4)
    newq = rng.integers(0, 60, size=(f.N, f.L)).astype(np.uint8)
    fd = os.open(dst, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    f.write(fd, newq)
    os.close(fd)
    g = fastx.NativeFastq(dst)
    seq, qual = g.pack()
    seq0, _ = f.pack()
    assert np.array_equal(seq, seq0) and np.array_equal(qual, newq)
    with open(dst) as fh:  # header without the comment, '+' line bare (kbbq/recalibrate.py:152-156)
        head = [fh.readline() for _ in range(4)]
    assert head[0] == "@" + recs[0][0] + "\n" and head[2] == "+\n"


def test_write_to_pipe_append_and_offset_agree(tmp_path):
    """The formatter writes at per-thread offsets into seekable files and serially into pipes and
    O_APPEND files; the bytes must be the same, and a seekable descriptor continues where it stood."""
    import threading
    rng = np.random.default_rng(9)
    recs = _random_records(rng, 20_003, 31)
    src = tmp_path / "in.fq"
    _write(src, recs)
    f = fastx.NativeFastq(src, threads=4)
    newq = rng.integers(0, 60, size=(f.N, f.L)).astype(np.uint8)
    plain = tmp_path / "plain.fq"
    fd = os.open(plain, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    os.write(fd, b"# kept\n")          # the descriptor does not start at 0
    f.write(fd, newq)
    os.write(fd, b"# after\n")         # ... and is left behind the records
    os.close(fd)
    want = plain.read_bytes()
    assert want.startswith(b"# kept\n@") and want.endswith(b"\n# after\n")
    body = want[len(b"# kept\n"):-len(b"# after\n")]
    app = tmp_path / "append.fq"
    fd = os.open(app, os.O_WRONLY | os.O_CREAT | os.O_APPEND)
    f.write(fd, newq)
    os.close(fd)
    assert app.read_bytes() == body
    r, w = os.pipe()
    got = []
    t = threading.Thread(target=lambda: got.append(b"".join(iter(lambda: os.read(r, 1 << 20), b""))))
    t.start()
    f.write(w, newq)
    os.close(w)
    t.join()
    os.close(r)
    assert got[0] == body
