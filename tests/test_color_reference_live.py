"""Movi Color against the LIVE reference (where it is compiled: oracle/_ref, the build container): a twelve-document pangenome,
bigger and more tangled than the committed fixture (three clades, shared taxon ids, ties between documents), is indexed, coloured and
sampled by the unmodified reference in a scratch directory; the reference's own answers to 2 000 reads are then compared with the
oracle's restatement and with the lanes of the device kernel run on the CPU.  Skipped where the reference binaries are absent."""
import os
import random
import subprocess

import numpy as np
import pytest

from _color import HostColors, HostSsa, emulate_color, emulate_sa
from conftest import parse_bpf_records
from movi_b200 import HostIndex, Reads
from oracle.port import OracleIndex

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref")
MOVI = os.path.join(BIN, "movi-regular-thresholds")
BWTGEN = os.path.join(ROOT, "movi_b200", "datagen", "_build", "bwt_thresholds")
pytestmark = pytest.mark.skipif(not (os.path.exists(MOVI) and os.path.exists(os.path.join(BIN, "movi-prepare-ref")) and os.path.exists(BWTGEN)),
                                reason="the reference is not compiled here (oracle/_ref)")


def sh(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, **kw).stdout


@pytest.fixture(scope="module")
def live(tmp_path_factory):
    w = tmp_path_factory.mktemp("color_live")
    idx = w / "idx"
    idx.mkdir()
    rng = random.Random(23)

    def mut(s, rate):
        s = list(s)
        for i in range(len(s)):
            if rng.random() < rate:
                s[i] = rng.choice("ACGT")
        return "".join(s)

    def rc(s):
        return s[::-1].translate(str.maketrans("ACGT", "TGCA"))

    base = "".join(rng.choice("ACGT") for _ in range(40000))
    clades = [mut(base, 0.05) for _ in range(3)]
    docs = [mut(clades[i // 4], 0.008) for i in range(12)]
    docs[5] = docs[4]                                    # two identical documents: every match counts for both (ties)
    with open(w / "ref.fasta", "w") as f:
        for i, d in enumerate(docs):
            f.write(">doc%d\n" % i)
            for k in range(0, len(d), 70):
                f.write(d[k:k + 70] + "\n")
    reads = w / "reads.fa"
    with open(reads, "w") as f:
        for i in range(2000):
            L = rng.randint(20, 2500)
            if i % 8 == 7:
                s = "".join(rng.choice("ACGT") for _ in range(L))
            else:
                d = docs[rng.randrange(12)]
                p = rng.randrange(0, len(d) - L)
                s = mut(d[p:p + L], rng.choice((0.0, 0.01, 0.05)))
                if rng.random() < 0.5:
                    s = rc(s)
            if i % 29 == 0:
                s = s[: L // 3] + "NN" + s[L // 3 + 2:]
            f.write(">q%d%s\n%s\n" % (i, "\tt" if i % 5 == 0 else "", s))
    fa = str(idx / "ref.fa")
    sh([os.path.join(BIN, "movi-prepare-ref"), str(w / "ref.fasta"), fa])
    (idx / "ref.fa.doc_ids").write_text("".join("%d\n" % t for t in (11, 11, 12, 13, 21, 21, 22, 23, 31, 32, 33, 33)))
    sh([BWTGEN, fa, fa, "-q"])
    sh([MOVI, "build", "--index", str(idx), "--fasta", fa])
    sh([MOVI, "color", "--index", str(idx)])
    sh([MOVI, "build-SA", "--index", str(idx)])
    return dict(w=w, idx=str(idx), reads=str(reads))


@pytest.mark.parametrize("flags,kw", [([], dict(min_len=1, pvalue=False)), (["--min-len", "5"], dict(min_len=5, pvalue=False)),
                                      (["--pvalue-scoring", "--min-len", "0"], dict(min_len=0, pvalue=True))])
def test_multi_classify_live(live, flags, kw):
    out = str(live["w"] / ("mc_%s.csv" % "_".join(f.strip("-") for f in flags)))
    sh([MOVI, "query", "--index", live["idx"], "--read", live["reads"], "--pml", "--multi-classify", "--no-prefetch", "-t1", "-o", out, "--no-output"] + flags)
    want = open(out).read()
    rd = Reads(live["reads"])
    orc = OracleIndex(live["idx"])
    assert orc.load_colors() == 9
    taxa, raw, score, _ = orc.query_multi_classify(rd.bases, rd.offsets, **kw)
    assert "".join("%s,%d,%d\n" % (n, t[0], t[1]) for n, t in zip(rd.ids(), taxa)) == want
    assert len({l.split(",", 1)[1] for l in want.split("\n") if l}) > 8            # many different answers, two-taxon lines among them
    hi = HostIndex(live["idx"])
    colors = HostColors(live["idx"], hi.r)
    res, _ = emulate_color(hi, colors, rd.bases, rd.offsets, lanes=37, **kw)
    assert colors.text(rd, res, **kw) == want


def test_sa_entries_live(live):
    w = live["w"]
    scratch = str(w / "sa_reads.fa")
    with open(live["reads"]) as f, open(scratch, "w") as g:
        g.write("".join(f.readlines()[:600]))                                       # 300 reads
    sh([MOVI, "query", "--index", live["idx"], "--read", scratch, "--pml", "--sa-entries", "--no-prefetch", "-t1"])
    want = parse_bpf_records(open(scratch + ".regular-thresholds.pml.sa_entries.bpf", "rb").read(), 8, header=False)
    rd = Reads(scratch)
    orc = OracleIndex(live["idx"])
    _, sa = orc.query_pml_sa(rd.bases, rd.offsets)
    hi = HostIndex(live["idx"])
    _, esa = emulate_sa(hi, HostSsa(live["idx"]), rd.bases, rd.offsets, lanes=11)
    assert len(want) == rd.count
    for i, (_, v) in enumerate(want):
        s, e = int(rd.offsets[i]), int(rd.offsets[i + 1])
        assert np.array_equal(v, sa[s:e]) and np.array_equal(v, esa[s:e]), i
