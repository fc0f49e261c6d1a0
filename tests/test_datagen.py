"""The bench data generators: BWT/threshold generator pinned by the reference's golden index sizes, index writer
checked byte for byte against index files written by the unmodified reference (tests/golden/c1)."""
import lzma
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD, TYPES
from movi_b200.datagen import synth


@pytest.fixture(scope="module")
def c1_bwt(tmp_path_factory):
    """ref.fasta -> cleaned text (sequence + reverse complement, reference src/prepare_ref.cpp:36-71) -> BWT + thresholds."""
    d = tmp_path_factory.mktemp("c1bwt")
    bwtgen, builder = synth.tools()
    txt = lzma.open(os.path.join(GOLD, "ref.fasta.xz")).read().decode()
    seq = "".join(l.strip() for l in txt.split("\n") if not l.startswith(">")).upper()
    seq = "".join(c if c in "ACGT" else "A" for c in seq)
    a = np.frombuffer(seq.encode(), dtype=np.uint8)
    with open(d / "ref.txt", "wb") as f:
        f.write(a.tobytes()); f.write(synth.revcomp(a).tobytes())
    subprocess.check_call([bwtgen, "--raw", str(d / "ref.txt"), str(d / "ref"), "-q"])
    return str(d / "ref"), builder


@pytest.mark.parametrize("typ", TYPES)
def test_index_writer_byte_identical_to_reference(c1_bwt, tmp_path, typ):
    prefix, builder = c1_bwt
    subprocess.check_call([builder, prefix, str(tmp_path / typ), typ], stderr=subprocess.DEVNULL)
    got = (tmp_path / typ / "index.movi").read_bytes()
    want = lzma.open(os.path.join(GOLD, "c1", typ, "index.movi.xz")).read()
    assert len(got) == len(want) == {"regular-thresholds": 948119, "blocked-thresholds": 711733, "sampled-thresholds": 475326}[typ]
    # bytes 9..15 are struct padding of MoviHeader (uninitialised stack bytes in the reference)
    assert got[:9] == want[:9] and got[16:] == want[16:]


def test_rle_and_raw_inputs_agree(c1_bwt, tmp_path):
    prefix, builder = c1_bwt
    bwtgen, _ = synth.tools()
    subprocess.check_call([bwtgen, "--raw", prefix + ".txt", str(tmp_path / "rle"), "--rle", "-q"])
    assert open(prefix + ".thr_pos", "rb").read() == open(tmp_path / "rle.thr_pos", "rb").read()
    subprocess.check_call([builder, str(tmp_path / "rle"), str(tmp_path / "a"), "regular-thresholds", "--rle"], stderr=subprocess.DEVNULL)
    subprocess.check_call([builder, prefix, str(tmp_path / "b"), "regular-thresholds"], stderr=subprocess.DEVNULL)
    assert (tmp_path / "a" / "index.movi").read_bytes() == (tmp_path / "b" / "index.movi").read_bytes()


def test_synthetic_pangenome_is_seeded_and_shaped():
    haps = synth.pangenome(20000, 4)
    again = synth.pangenome(20000, 4)
    assert all(np.array_equal(a, b) for a, b in zip(haps, again))
    base = synth.base_genome(20000)
    for h in haps:
        assert abs(len(h) - 20000) < 30 and set(np.unique(h)) <= set(b"ACGT")
        m = min(len(h), 20000)
    bases, offsets = synth.make_reads(haps, 500, 150)
    assert len(bases) == 500 * 150 and offsets[-1] == len(bases) and set(np.unique(bases)) <= set(b"ACGT")
    b2, _ = synth.make_reads(haps, 500, 150)
    assert np.array_equal(bases, b2)
