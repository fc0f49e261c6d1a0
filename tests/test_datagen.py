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


def test_synthetic_index_writer_is_read_by_the_reference_and_the_oracle(index_dirs, tmp_path):
    """datagen/index_file.py writes the index.movi of tables synthesised in HBM (bench workload c4).  Here it re-writes the
    reference-built c1 index from its parsed fields: the unmodified reference (when compiled here) and the oracle must give
    the same PML / ZML / count output on the re-written file as on the original."""
    import ctypes as C
    from movi_b200 import HostIndex
    from movi_b200.datagen.index_file import write_regular_thresholds
    from oracle.port import OracleIndex, pack_reads
    src = index_dirs[("regular-thresholds", False)]
    hix = HostIndex(src)
    d = hix.desc
    rows = np.ctypeslib.as_array(C.cast(d.rows, C.POINTER(C.c_uint32)), shape=(hix.r * 2,)).copy()
    half = (hix.r // 2) * 2
    out = write_regular_thresholds(str(tmp_path / "rewritten"), hix.n, hix.r, hix.end_bwt_idx, list(d.end_bwt_idx_thresholds),
                                   [rows[:half], rows[half:]], d.last_runs, d.last_offsets, d.first_runs, d.first_offsets,
                                   original_r=hix.original_r)
    reads = os.path.join(GOLD, "reads", "synth300.fasta")
    a, b = OracleIndex(src), OracleIndex(out)
    from movi_b200 import Reads
    rd = Reads(reads)
    assert np.array_equal(a.query_pml(rd.bases, rd.offsets), b.query_pml(rd.bases, rd.offsets))
    assert np.array_equal(a.query_zml(rd.bases, rd.offsets), b.query_zml(rd.bases, rd.offsets))
    ca, cb = a.query_count(rd.bases, rd.offsets), b.query_count(rd.bases, rd.offsets)
    assert np.array_equal(ca[0], cb[0]) and np.array_equal(ca[1], cb[1])
    ref_bin = os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "movi-regular-thresholds")
    if not os.path.exists(ref_bin):
        pytest.skip("oracle/_ref not built here")
    for q in ("--pml", "--zml", "--count"):
        outs = []
        for idx in (src, out):
            r = subprocess.run([ref_bin, "query", "--index", idx, "--read", reads, q, "--stdout", "--no-prefetch", "-t1"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            assert r.returncode == 0
            outs.append(r.stdout)
        assert outs[0] == outs[1] and len(outs[0]) > 1000
        # the multithreaded prefetch path (what the bench's CPU arm runs) loads it as well
        r = subprocess.run([ref_bin, "query", "--index", out, "--read", reads, q, "--no-output", "-t", "4", "-s", "16"],
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        assert r.returncode == 0 and b"processing the reads" in r.stderr
