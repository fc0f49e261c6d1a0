"""Movi Color on the CPU side: the host loaders / writers of include/movi_host.h and the lane state machine of the device kernel
(movi_b200/csrc/ext_kernel.cuh, compiled for the host by tests/native/ext_emulate.cu) against the reference's outputs and the oracle."""
import lzma
import os

import numpy as np
import pytest

from _color import HostColors, HostSsa, emulate_color, emulate_sa, random_reads, result_scores
from conftest import COLOR_VARIANTS, GOLD, color_expected, parse_bpf_records
from movi_b200 import HostIndex, Reads
from movi_b200 import _native as N
from movi_b200.api import COLOR_RESULT_DTYPE
from oracle.port import OracleIndex, pack_reads


@pytest.fixture(scope="module")
def case(color_index_dir):
    hi = HostIndex(color_index_dir)
    return dict(dir=color_index_dir, hi=hi, colors=HostColors(color_index_dir, hi.r), ssa=HostSsa(color_index_dir),
                reads=Reads(os.path.join(GOLD, "color", "reads.fa")), orc=OracleIndex(color_index_dir))


def test_host_loaders_match_the_oracle(case):
    assert case["colors"].num_species == case["orc"].load_colors() == 5
    assert [case["colors"].taxon(s) for s in range(5)] == [case["orc"].taxon(s) for s in range(5)] == [562, 1280, 9606, 10090, 28901]
    assert case["ssa"].rate == 100 and case["ssa"].rows == case["hi"].r


def test_vectors_of_vectors_table_flattens_to_the_reference_flat_table(case):
    """doc_sets.bin (`color --color-vectors`) loaded and flattened on the host == doc_sets_flat.bin as the reference wrote it; the
    kernel's lanes on it reproduce the reference's `--multi-classify --color-vectors` file."""
    import ctypes as C
    vec = HostColors(case["dir"], case["hi"].r, kind=1)
    flat = case["colors"]
    assert vec.num_species == flat.num_species and vec.desc.flat_size == flat.desc.flat_size
    n = int(flat.desc.flat_size)
    assert C.string_at(vec.desc.flat_colors, 2 * n) == C.string_at(flat.desc.flat_colors, 2 * n)
    assert C.string_at(vec.desc.doc_set_flat_inds, 5 * case["hi"].r) == C.string_at(flat.desc.doc_set_flat_inds, 5 * case["hi"].r)
    rd = case["reads"]
    res, _ = emulate_color(case["hi"], vec, rd.bases, rd.offsets)
    assert vec.text(rd, res) == open(os.path.join(GOLD, "color", "expected.multi_classify.vectors.csv")).read()


@pytest.mark.parametrize("tag", sorted(COLOR_VARIANTS))
def test_host_writer_reproduces_the_reference_file_from_oracle_results(case, tag):
    """movi_host_write_multi_classify: the oracle's per-read results in, the reference's -o file out."""
    v = COLOR_VARIANTS[tag]
    rd = case["reads"]
    taxa, raw, score, _ = case["orc"].query_multi_classify(rd.bases, rd.offsets, **v)
    res = np.zeros(rd.count, COLOR_RESULT_DTYPE)
    res["best_doc"], res["second_best_doc"], res["sum_matching_lengths"] = raw[:, 0], raw[:, 1], raw[:, 2]
    if v["pvalue"]:
        res["best_raw"], res["second_raw"] = score[:, 0].view(np.uint64), score[:, 1].view(np.uint64)
    else:
        res["best_raw"], res["second_raw"] = score[:, 0].astype(np.uint64), score[:, 1].astype(np.uint64)
    assert case["colors"].text(rd, res, **v) == color_expected(tag)


@pytest.mark.parametrize("tag", sorted(COLOR_VARIANTS))
@pytest.mark.parametrize("lanes,wide", [(1, False), (7, False), (64, True)])
def test_kernel_lanes_on_the_cpu_multi_classify(case, tag, lanes, wide):
    v = COLOR_VARIANTS[tag]
    rd = case["reads"]
    res, ml = emulate_color(case["hi"], case["colors"], rd.bases, rd.offsets, lanes=lanes, wide=wide, **v)
    assert case["colors"].text(rd, res, **v) == color_expected(tag)
    taxa, raw, score, oml = case["orc"].query_multi_classify(rd.bases, rd.offsets, **v)
    assert np.array_equal(np.stack([res["best_doc"], res["second_best_doc"], res["sum_matching_lengths"]], axis=1), raw)
    assert np.array_equal(result_scores(res, v["pvalue"]), score)
    assert np.array_equal(ml, oml)


@pytest.mark.parametrize("lanes,wide", [(1, False), (5, False), (33, True)])
def test_kernel_lanes_on_the_cpu_sa_entries(case, lanes, wide):
    rd = case["reads"]
    ml, sa = emulate_sa(case["hi"], case["ssa"], rd.bases, rd.offsets, lanes=lanes, wide=wide)
    want = parse_bpf_records(lzma.open(os.path.join(GOLD, "color", "expected.sa_entries.bpf.xz")).read(), 8, header=False)
    for i, (_, v) in enumerate(want):
        assert np.array_equal(v, sa[int(rd.offsets[i]):int(rd.offsets[i + 1])]), i
    assert np.array_equal(ml, case["orc"].query_pml(rd.bases, rd.offsets))


def test_kernel_lanes_on_the_cpu_ragged_reads(case):
    """3000 ragged reads (empty, 1-base, with N, random) -- every lane slot is reused many times."""
    b, o = pack_reads(random_reads(case["dir"], 3000, seed=5))
    for v in (dict(min_len=1, pvalue=False), dict(min_len=0, pvalue=True)):
        res, ml = emulate_color(case["hi"], case["colors"], b, o, lanes=48, **v)
        taxa, raw, score, oml = case["orc"].query_multi_classify(b, o, **v)
        assert np.array_equal(np.stack([res["best_doc"], res["second_best_doc"], res["sum_matching_lengths"]], axis=1), raw)
        assert np.array_equal(result_scores(res, v["pvalue"]), score) and np.array_equal(ml, oml)
    ml, sa = emulate_sa(case["hi"], case["ssa"], b, o, lanes=48)
    oml, osa = case["orc"].query_pml_sa(b, o)
    assert np.array_equal(ml, oml) and np.array_equal(sa, osa)


def test_sa_writer_reproduces_the_reference_file(case, tmp_path):
    rd = case["reads"]
    _, sa = case["orc"].query_pml_sa(rd.bases, rd.offsets)
    f = tmp_path / "x.sa_entries.bpf"
    fd = os.open(f, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        assert N.host_lib().movi_host_sa_append(fd, rd._h, sa.ctypes.data) == 0
    finally:
        os.close(fd)
    assert f.read_bytes() == lzma.open(os.path.join(GOLD, "color", "expected.sa_entries.bpf.xz")).read()


def test_filter_echo_pieces_on_the_cpu(tmp_path):
    """filter_kernels.cuh's per-read pieces (keep rule, sizes, the strided copy) against movi_host_classify_bins in filter mode,
    which restates the reference's writer: same text, for --filter and --filter --invert."""
    import ctypes as C
    import threading
    from _color import emulator
    rng = np.random.default_rng(8)
    n = 500
    path = tmp_path / "r.fa"
    with open(path, "w") as f:
        for i in range(n):
            L = int(rng.integers(0, 300)) if i % 9 else 0
            f.write(">read%d%s\n%s\n" % (i, " x" if i % 2 else "", "".join(rng.choice(list("ACGTN"), L))))
    rd = Reads(str(path))
    assert rd.count == n
    above = rng.integers(0, 5, n).astype(np.uint32); below = rng.integers(0, 5, n).astype(np.uint32)
    above[::9] = 0; below[::9] = 0                                  # empty reads have no bins
    total = np.zeros(n, np.uint64)
    ids, id_off = rd.id_blob
    H = N.host_lib()
    L = emulator()
    L.movi_filter_emulate.restype = C.c_uint64
    L.movi_filter_emulate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    for invert in (0, 1):
        rfd, wfd = os.pipe()
        buf = []
        t = threading.Thread(target=lambda: buf.append(os.fdopen(rfd, "rb").read()))
        t.start()
        assert H.movi_host_classify_bins(wfd, 2 if invert else 1, rd._h, 5, above.ctypes.data, below.ctypes.data, total.ctypes.data, None) == 0
        os.close(wfd); t.join()
        text = np.zeros(len(buf[0]) + 64, np.uint8)
        got = L.movi_filter_emulate(above.ctypes.data, below.ctypes.data, invert, rd.bases.ctypes.data, rd.offsets.ctypes.data,
                                    ids.ctypes.data, id_off.ctypes.data, n, text.ctypes.data, len(text))
        assert got == len(buf[0]) and text[:got].tobytes() == buf[0]
        assert 0 < got


@pytest.mark.parametrize("sep", [False, True])
def test_kernel_lanes_on_the_cpu_separators_index(index_dirs, tmp_path, sep):
    """The walk of the colour kernel on an index WITH separators ('%' rows, shifted threshold indices, the side table): a made-up
    colour table (three documents spread over the rows) next to the reference-built c1 index; lengths == the oracle's PML, the
    counters == the oracle's restatement, on the read sets that hold '%', N and lower-case characters."""
    import shutil
    import struct
    src = index_dirs[("regular-thresholds", sep)]
    d = tmp_path / "idx"
    d.mkdir()
    shutil.copy(os.path.join(src, "index.movi"), d / "index.movi")
    hi = HostIndex(str(d))
    r = hi.r
    rng = np.random.default_rng(3)
    colours = [[0], [1], [2], [0, 1], [1, 2], [0, 1, 2], [2, 0]]
    flat, starts = [], []
    for c in colours:
        starts.append(len(flat)); flat += [len(c)] + c
    flat = np.array(flat, dtype="<u2")
    pick = rng.integers(0, len(colours) + 1, r)                       # some rows point beyond the table: "dropped by compression"
    offs = np.array(starts + [len(flat) + 5], dtype=np.uint64)[pick]
    inds = np.zeros((r, 5), np.uint8)
    for k in range(5):
        inds[:, k] = (offs >> np.uint64(8 * k)) & np.uint64(0xFF)
    with open(d / "doc_sets_flat.bin", "wb") as f:
        f.write(struct.pack("<Q", len(flat)) + flat.tobytes() + inds.tobytes())
    (d / "ref.fa.doc_offsets").write_text("1000\n5000\n5600\n")
    colors = HostColors(str(d), r)
    orc = OracleIndex(str(d))
    assert orc.load_colors() == colors.num_species == 3
    for read_set in ("edge.fasta", "synth300.fasta", "sample.fastq"):
        rd = Reads(os.path.join(GOLD, "reads", read_set))
        for v in (dict(min_len=1, pvalue=False), dict(min_len=4, pvalue=False), dict(min_len=0, pvalue=True)):
            res, ml = emulate_color(hi, colors, rd.bases, rd.offsets, lanes=9, **v)
            taxa, raw, score, oml = orc.query_multi_classify(rd.bases, rd.offsets, **v)
            assert np.array_equal(ml, orc.query_pml(rd.bases, rd.offsets)), (read_set, v)
            assert np.array_equal(np.stack([res["best_doc"], res["second_best_doc"], res["sum_matching_lengths"]], axis=1), raw), (read_set, v)
            assert np.array_equal(result_scores(res, v["pvalue"]), score), (read_set, v)
            got = colors.text(rd, res, **v)
            want = "".join("%s,%d,%d\n" % (name, t[0], t[1]) for name, t in zip(rd.ids(), taxa))
            assert got == want, (read_set, v)


@pytest.mark.parametrize("sep", [False, True])
def test_kernel_lanes_on_the_cpu_sa_entries_separators_index(index_dirs, sep):
    """SA entries on the reference-built c1 indexes (with and without separators), the sampled suffix array built by the oracle's
    restatement of `build-SA` (pinned against the reference's ssa.movi in test_color_oracle.py), two sample rates."""
    import ctypes as C
    d = index_dirs[("regular-thresholds", sep)]
    hi = HostIndex(d)
    orc = OracleIndex(d)
    for rate in (100, 7):
        r_, sampled, all_p = orc.build_ssa(rate)

        class S:                       # what emulate_sa takes (HostSsa's fields)
            pass
        s = S(); s.rate = rate; s.sampled = sampled.ctypes.data; s.all_p = all_p.ctypes.data
        for read_set in ("edge.fasta", "synth300.fasta"):
            rd = Reads(os.path.join(GOLD, "reads", read_set))
            ml, sa = emulate_sa(hi, s, rd.bases, rd.offsets, lanes=6)
            oml, osa = orc.query_pml_sa(rd.bases, rd.offsets)
            assert np.array_equal(ml, oml) and np.array_equal(sa, osa), (read_set, rate)
            assert np.array_equal(ml, orc.query_pml(rd.bases, rd.offsets))
