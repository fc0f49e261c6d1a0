"""Parity of the CUDA path (through the C ABI) with the reference outputs and with the CPU oracle.
Bit-exact: PML / ZML are uint16 per base, counts are uint64 per read."""
import numpy as np
import pytest

from conftest import READ_SETS, TYPES, expected_text, reads_path
from _fmt import count_text, ml_text

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def structures(index_dirs):
    from movi_b200 import MoveStructure
    ms = {k: MoveStructure(v) for k, v in index_dirs.items()}
    yield ms
    for m in ms.values():
        m.close()


@pytest.fixture(scope="module")
def read_sets():
    from movi_b200 import Reads
    return {s: Reads(reads_path(s)) for s in READ_SETS}


_DEFAULTS = {}


@pytest.fixture(autouse=True)
def _default_kernels_after_each_test(structures):
    """Tests that walk through the kernel variants leave the shared contexts on the library's defaults."""
    if not _DEFAULTS:
        ms = next(iter(structures.values()))
        _DEFAULTS.update(pml_sector=ms.get_option("pml_sector"))
    yield
    for ms in structures.values():
        for k, v in _DEFAULTS.items():
            ms.set_option(k, v)


@pytest.mark.parametrize("sector", [0, 1])       # plain row fetch / the sector-cached fetch tables beyond the L2 get
@pytest.mark.parametrize("sep", [False, True])
@pytest.mark.parametrize("typ", TYPES)
@pytest.mark.parametrize("read_set", READ_SETS)
def test_pml_matches_reference(structures, read_sets, typ, read_set, sep, sector):
    rd = read_sets[read_set]
    ms = structures[(typ, sep)]
    ms.set_option("pml_sector", sector)
    ml = ms.query_pml(rd.bases, rd.offsets)
    assert ml_text(rd.ids(), rd.offsets, ml) == expected_text(read_set, "pml", sep)


@pytest.mark.parametrize("sep", [False, True])
@pytest.mark.parametrize("typ", TYPES[:2])
@pytest.mark.parametrize("read_set", READ_SETS)
def test_zml_matches_reference(structures, read_sets, typ, read_set, sep):
    rd = read_sets[read_set]
    ml = structures[(typ, sep)].query_zml(rd.bases, rd.offsets)
    assert ml_text(rd.ids(), rd.offsets, ml) == expected_text(read_set, "zml", sep)


@pytest.mark.parametrize("sep", [False, True])
@pytest.mark.parametrize("typ", TYPES[:2])
@pytest.mark.parametrize("read_set", READ_SETS)
def test_count_matches_reference(structures, read_sets, typ, read_set, sep):
    rd = read_sets[read_set]
    counts, pos = structures[(typ, sep)].query_backward_search(rd.bases, rd.offsets)
    assert count_text(rd.ids(), rd.offsets, counts, pos) == expected_text(read_set, "count", sep)


def test_sampled_rejects_zml_and_count(structures, read_sets):
    """reference src/movi.cpp:228-233"""
    from movi_b200 import MoviGpuError
    rd = read_sets["sample.fasta"]
    ms = structures[("sampled-thresholds", False)]
    with pytest.raises(MoviGpuError, match="not suppported by the sampled modes"):
        ms.query_zml(rd.bases, rd.offsets)
    with pytest.raises(MoviGpuError, match="not suppported by the sampled modes"):
        ms.query_backward_search(rd.bases, rd.offsets)


def _random_reads(seed, n_reads, max_len, ref_seq, p_illegal=0.01):
    rng = np.random.default_rng(seed)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    both = ref_seq + ref_seq.translate(comp)[::-1]
    reads = []
    for i in range(n_reads):
        L = int(rng.integers(0 if i % 97 == 0 else 1, max_len + 1))
        if rng.random() < 0.15:
            s = bytearray(rng.choice(np.frombuffer(b"ACGT", np.uint8), L).tobytes())
        else:
            p = int(rng.integers(0, len(both) - L))
            s = bytearray(both[p:p + L])
            for j in np.nonzero(rng.random(L) < 0.02)[0]:
                s[j] = b"ACGT"[int(rng.integers(0, 4))]
        for j in np.nonzero(rng.random(L) < p_illegal)[0]:
            s[j] = b"NnRa%$"[int(rng.integers(0, 6))]
        reads.append(bytes(s))
    return reads


@pytest.fixture(scope="module")
def ref_seq():
    import lzma, os
    from conftest import GOLD
    txt = lzma.open(os.path.join(GOLD, "ref.fasta.xz")).read().decode()
    return "".join(l.strip() for l in txt.split("\n") if not l.startswith(">")).upper().encode()


@pytest.mark.parametrize("sep", [False, True])
@pytest.mark.parametrize("typ", TYPES)
def test_random_reads_vs_oracle(structures, index_dirs, ref_seq, typ, sep):
    """20k ragged reads (empty, 1-base, illegal characters, up to 600 bases) -- every value must equal the oracle's."""
    from oracle.port import OracleIndex, pack_reads
    reads = _random_reads(7, 20000, 600, ref_seq)
    bases, offsets = pack_reads(reads)
    orc = OracleIndex(index_dirs[(typ, sep)])
    ms = structures[(typ, sep)]
    want = orc.query_pml(bases, offsets, threads=0)
    for sector in (0, 1):
        ms.set_option("pml_sector", sector)
        assert np.array_equal(ms.query_pml(bases, offsets), want)
    ms.set_option("pml_sector", -1)
    if typ != "sampled-thresholds":
        # (reads of length 0 are outside the reference's domain for zml/count -- it indexes R[-1]; both sides emit nothing / 0)
        want_z = orc.query_zml(bases, offsets, threads=0)
        c2, p2 = orc.query_count(bases, offsets, threads=0)
        for sector in (0, 1):      # 1: the sector-cached form big tables use
            ms.set_option("pml_sector", sector)
            try:
                assert np.array_equal(ms.query_zml(bases, offsets), want_z)
                c1, p1 = ms.query_backward_search(bases, offsets)
                assert np.array_equal(c1, c2) and np.array_equal(p1, p2)
            finally:
                ms.set_option("pml_sector", -1)


def test_small_pipeline_slices(structures, index_dirs, ref_seq):
    """The host entry point slices big batches; force many tiny slices and compare with one big slice."""
    from oracle.port import pack_reads
    reads = _random_reads(11, 5000, 300, ref_seq)
    bases, offsets = pack_reads(reads)
    ms = structures[("regular-thresholds", False)]
    big = ms.query_pml(bases, offsets)
    old = ms.get_option("batch_bases")
    ms.set_option("batch_bases", 65536)
    try:
        small = ms.query_pml(bases, offsets)
        z_small = ms.query_zml(bases, offsets)
        c_small = ms.query_backward_search(bases, offsets)
    finally:
        ms.set_option("batch_bases", old)
    assert np.array_equal(big, small)
    assert np.array_equal(z_small, ms.query_zml(bases, offsets))
    c_big = ms.query_backward_search(bases, offsets)
    assert np.array_equal(c_small[0], c_big[0]) and np.array_equal(c_small[1], c_big[1])


def test_device_entry_points_with_torch(structures, index_dirs, ref_seq):
    """The *_device entry points on torch-owned HBM buffers, plus device-side classification."""
    import torch
    from oracle.port import OracleIndex, pack_reads, classify
    reads = [r for r in _random_reads(13, 4000, 500, ref_seq) if len(r) > 0]
    bases, offsets = pack_reads(reads)
    ms = structures[("regular-thresholds", False)]
    orc = OracleIndex(index_dirs[("regular-thresholds", False)])
    dev = torch.device("cuda:0")
    pad = (-len(bases)) % 32 + 32
    d_bases = torch.from_numpy(np.concatenate([bases, np.zeros(pad, np.uint8)])).to(dev)
    d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    d_out = torch.zeros(len(bases) + 16, dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms.query_pml_device(d_bases.data_ptr(), d_off.data_ptr(), len(reads), d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    ms.check_device_status()
    got = d_out.cpu().numpy().view(np.uint16)[: len(bases)]
    want = orc.query_pml(bases, offsets, threads=0)
    assert np.array_equal(got, want)
    n = len(reads)
    d_above = torch.zeros(n, dtype=torch.int32, device=dev)
    d_below = torch.zeros(n, dtype=torch.int32, device=dev)
    d_sum = torch.zeros(n, dtype=torch.int64, device=dev)
    ms.classify_device(d_out.data_ptr(), d_off.data_ptr(), n, 150, 7, d_above.data_ptr(), d_below.data_ptr(), d_sum.data_ptr(), stream)
    torch.cuda.synchronize()
    a, b, s = d_above.cpu().numpy(), d_below.cpu().numpy(), d_sum.cpu().numpy()
    for i in range(0, n, 37):
        f, ea, eb, es, _ = classify(want[int(offsets[i]):int(offsets[i + 1])], 150, 7)
        assert (a[i], b[i], s[i]) == (ea, eb, es)
    d_cnt = torch.zeros(n, dtype=torch.int64, device=dev)
    d_pos = torch.zeros(n, dtype=torch.int32, device=dev)
    ms.query_count_device(d_bases.data_ptr(), d_off.data_ptr(), n, d_cnt.data_ptr(), d_pos.data_ptr(), stream)
    ms.query_zml_device(d_bases.data_ptr(), d_off.data_ptr(), n, d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    ms.check_device_status()
    c2, p2 = orc.query_count(bases, offsets, threads=0)
    assert np.array_equal(d_cnt.cpu().numpy().view(np.uint64), c2) and np.array_equal(d_pos.cpu().numpy(), p2)
    assert np.array_equal(d_out.cpu().numpy().view(np.uint16)[: len(bases)], orc.query_zml(bases, offsets, threads=0))


@pytest.mark.parametrize("sep", [False, True])
def test_blocked_compact_device_layout(index_dirs, read_sets, ref_seq, sep):
    """MOVI_GPU_LAYOUT_COMPACT: blocked-thresholds rows stay 6 bytes in HBM, the id is completed per step from id_blocks."""
    from movi_b200 import MoveStructure
    from movi_b200.api import LAYOUT_COMPACT, LAYOUT_EXPANDED8
    from oracle.port import OracleIndex, pack_reads
    d = index_dirs[("blocked-thresholds", sep)]
    ms = MoveStructure(d, layout=LAYOUT_COMPACT)
    try:
        assert ms.device_layout == LAYOUT_COMPACT
        assert ms.index_device_bytes < 6.1 * ms.r + 4096          # 6 bytes per row (+ id_blocks), not 8
        for read_set in READ_SETS:
            rd = read_sets[read_set]
            assert ml_text(rd.ids(), rd.offsets, ms.query_pml(rd.bases, rd.offsets)) == expected_text(read_set, "pml", sep)
            assert ml_text(rd.ids(), rd.offsets, ms.query_zml(rd.bases, rd.offsets)) == expected_text(read_set, "zml", sep)
            c, p = ms.query_backward_search(rd.bases, rd.offsets)
            assert count_text(rd.ids(), rd.offsets, c, p) == expected_text(read_set, "count", sep)
        bases, offsets = pack_reads(_random_reads(17, 20000, 600, ref_seq))
        orc = OracleIndex(d)
        for wide in (0, 1):
            ms.set_option("force_wide", wide)
            assert np.array_equal(ms.query_pml(bases, offsets), orc.query_pml(bases, offsets, threads=0))
            assert np.array_equal(ms.query_zml(bases, offsets), orc.query_zml(bases, offsets, threads=0))
            c1, p1 = ms.query_backward_search(bases, offsets)
            c2, p2 = orc.query_count(bases, offsets, threads=0)
            assert np.array_equal(c1, c2) and np.array_equal(p1, p2)
    finally:
        ms.close()
    # the same index expanded: 8 bytes per row
    ms8 = MoveStructure(d, layout=LAYOUT_EXPANDED8)
    assert ms8.device_layout == LAYOUT_EXPANDED8 and ms8.index_device_bytes >= 8 * ms8.r
    ms8.close()


@pytest.mark.parametrize("sep", [False, True])
def test_sampled_compact_device_layout(index_dirs, read_sets, ref_seq, sep):
    """MOVI_GPU_LAYOUT_COMPACT for sampled-thresholds: 4 bytes per row (the 3-byte row + its walk-back distance from the id
    checkpoint) + the checkpoints = 5 bytes per row against 8; MoveStructure::get_id is one checkpoint read per LF step."""
    from movi_b200 import MoveStructure, MoviGpuError
    from movi_b200.api import LAYOUT_COMPACT
    from oracle.port import OracleIndex, pack_reads
    d = index_dirs[("sampled-thresholds", sep)]
    ms = MoveStructure(d, layout=LAYOUT_COMPACT)
    try:
        assert ms.device_layout == LAYOUT_COMPACT
        assert ms.index_device_bytes < (4 + 5 * (5 if sep else 4) / 20 + 0.05) * ms.r + 4096       # rows + sigma checkpoint tables of r / 20 entries
        for read_set in READ_SETS:
            rd = read_sets[read_set]
            assert ml_text(rd.ids(), rd.offsets, ms.query_pml(rd.bases, rd.offsets)) == expected_text(read_set, "pml", sep)
        bases, offsets = pack_reads(_random_reads(19, 20000, 600, ref_seq))
        want = OracleIndex(d).query_pml(bases, offsets, threads=0)
        for wide in (0, 1):
            ms.set_option("force_wide", wide)
            assert np.array_equal(ms.query_pml(bases, offsets), want)
        with pytest.raises(MoviGpuError, match="not suppported by the sampled modes"):
            ms.query_zml(bases, offsets)
    finally:
        ms.close()


def test_fused_query_classify_matches_oracle(structures, index_dirs, read_sets, ref_seq):
    """movi_gpu_query_classify: PML / ZML binned on the device; the three report columns per read must equal
    Classifier::classify on the oracle's matching lengths (and the lengths themselves, when asked for, are unchanged)."""
    from oracle.port import OracleIndex, classify, pack_reads
    ms = structures[("regular-thresholds", False)]
    orc = OracleIndex(index_dirs[("regular-thresholds", False)])
    sets = [(read_sets[s].bases, read_sets[s].offsets) for s in ("long6.fasta", "synth300.fasta", "edge.fasta")]
    sets.append(pack_reads(_random_reads(23, 5000, 900, ref_seq)))
    old = ms.get_option("batch_bases")
    for bases, offsets in sets:
        for zml in (False, True):
            want = orc.query_zml(bases, offsets, threads=0) if zml else orc.query_pml(bases, offsets, threads=0)
            for batch in (old, 65536):           # one slice and many small slices
                ms.set_option("batch_bases", batch)
                try:
                    a, b, s, ml = ms.query_classify(bases, offsets, 150, 7, zml=zml, want_ml=True)
                    a2, b2, s2 = ms.query_classify(bases, offsets, 150, 7, zml=zml)
                finally:
                    ms.set_option("batch_bases", old)
                assert np.array_equal(ml, want)
                assert np.array_equal(a, a2) and np.array_equal(b, b2) and np.array_equal(s, s2)
                for i in range(0, len(offsets) - 1, max(1, (len(offsets) - 1) // 400)):
                    v = want[int(offsets[i]):int(offsets[i + 1])]
                    if len(v) == 0:
                        assert (a[i], b[i], s[i]) == (0, 0, 0)
                        continue
                    _, ea, eb, es, _ = classify(v, 150, 7)
                    assert (a[i], b[i], s[i]) == (ea, eb, es), i


@pytest.mark.parametrize("typ", ["regular-thresholds", "blocked-thresholds", "sampled-thresholds"])
def test_byte_wide_results_for_short_reads(structures, index_dirs, read_sets, ref_seq, typ):
    """movi_gpu_query_ml8: reads of <= 255 bases come back as uint8, equal to the oracle's uint16 values; a longer read is
    refused (its lengths may not fit one byte)."""
    from movi_b200 import MoviGpuError
    from oracle.port import OracleIndex, pack_reads
    ms = structures[(typ, False)]
    orc = OracleIndex(index_dirs[(typ, False)])
    sets = [(read_sets["sample.fastq"].bases, read_sets["sample.fastq"].offsets), pack_reads(_random_reads(31, 6000, 255, ref_seq))]
    old = ms.get_option("batch_bases")
    for bases, offsets in sets:
        for zml in ((False,) if typ == "sampled-thresholds" else (False, True)):
            want = orc.query_zml(bases, offsets, threads=0) if zml else orc.query_pml(bases, offsets, threads=0)
            assert want.max() <= 255
            for batch in (old, 65536):
                ms.set_option("batch_bases", batch)
                try:
                    got = ms.query_ml8(bases, offsets, zml=zml)
                finally:
                    ms.set_option("batch_bases", old)
                assert got.dtype == np.uint8 and np.array_equal(got, want.astype(np.uint8))
    bases, offsets = pack_reads([b"ACGT" * 64])          # 256 bases
    with pytest.raises(MoviGpuError, match="longer than 255"):
        ms.query_ml8(bases, offsets)


@pytest.mark.parametrize("typ", TYPES)
def test_matching_lengths_saturate_at_65535(structures, index_dirs, ref_seq, typ):
    """A 75 kbp read that is a substring of the indexed text: the running match length passes 2^16; the stored value saturates
    at 65535 (MoveQuery::add_ml, include/move_query.hpp:26-38) while the count goes on (tests/test_basics.cpp:304-315)."""
    from oracle.port import OracleIndex, pack_reads
    read = ref_seq[2000:77000]
    bases, offsets = pack_reads([read, read[:300], read[-70000:]])
    ms = structures[(typ, False)]
    orc = OracleIndex(index_dirs[(typ, False)])
    want = orc.query_pml(bases, offsets, threads=0)
    assert want.max() == 65535 and (want == 65535).sum() > 5000
    for sector in (0, 1):
        ms.set_option("pml_sector", sector)
        assert np.array_equal(ms.query_pml(bases, offsets), want)
    if typ != "sampled-thresholds":
        want_z = orc.query_zml(bases, offsets, threads=0)
        assert want_z.max() == 65535
        assert np.array_equal(ms.query_zml(bases, offsets), want_z)
        c, p = ms.query_backward_search(bases, offsets)
        c2, p2 = orc.query_count(bases, offsets, threads=0)
        assert np.array_equal(c, c2) and np.array_equal(p, p2) and p[0] == 0 and c[0] >= 1


@pytest.mark.parametrize("sep", [False, True])
def test_device_row_format_round_trips_to_the_file_format(structures, index_dirs, sep):
    """Upload permutes the bit fields of the regular-thresholds rows into the device format (threshold mask per read character,
    character at the top of the word); movi_gpu_index_export_rows must give back the file's rows bit for bit."""
    import ctypes as C
    from movi_b200 import HostIndex
    d = index_dirs[("regular-thresholds", sep)]
    hix = HostIndex(d)
    r = hix.r
    file_rows = np.ctypeslib.as_array(C.cast(hix.desc.rows, C.POINTER(C.c_uint32)), shape=(r, 2)).copy()
    got = structures[("regular-thresholds", sep)].export_rows()
    assert got.shape == (r, 2)
    assert np.array_equal(got[:, 0], file_rows[:, 0])
    diff = np.nonzero(got[:, 1] != file_rows[:, 1])[0]
    # the only freedom: threshold bits of rows whose character lies outside the four query characters ('$', '%')
    for i in diff:
        c = (int(file_rows[i, 1]) >> 13) & 7
        assert i == hix.end_bwt_idx or (sep and c == 0) or c >= 4 + (1 if sep else 0), (int(i), hex(int(file_rows[i, 1])), hex(int(got[i, 1])))
        assert (int(got[i, 1]) ^ int(file_rows[i, 1])) & ~((1 << 11) | (1 << 12) | (1 << 27)) == 0
    assert len(diff) <= 16


@pytest.mark.parametrize("sep", [False, True])
@pytest.mark.parametrize("typ", TYPES)
def test_rpml_matches_reference_and_oracle(structures, index_dirs, read_sets, ref_seq, typ, sep):
    """MOVI_GPU_FLAG_RPML = `movi query --rpml` (reposition_randomly): the reference's own output on every read set, the oracle
    on 20 k ragged random reads; expanded and compact device layouts; host-buffer, device and byte-wide entry points."""
    import torch
    from movi_b200 import MoveStructure
    from oracle.port import OracleIndex, pack_reads
    ms = structures[(typ, sep)]
    for read_set in READ_SETS:
        rd = read_sets[read_set]
        ml = ms.query_pml(rd.bases, rd.offsets, rpml=True)
        assert ml_text(rd.ids(), rd.offsets, ml) == expected_text(read_set, "rpml", sep)
    orc = OracleIndex(index_dirs[(typ, sep)])
    bases, offsets = pack_reads(_random_reads(13, 20000, 600, ref_seq))
    want = orc.query_rpml(bases, offsets, threads=0)
    assert not np.array_equal(want, orc.query_pml(bases, offsets, threads=0))
    assert np.array_equal(ms.query_pml(bases, offsets, rpml=True), want)
    assert np.array_equal(ms.query_pml(bases, offsets), orc.query_pml(bases, offsets, threads=0))     # the flag does not stick
    if typ != "regular-thresholds" and not sep:
        compact = MoveStructure(index_dirs[(typ, sep)], layout=2)
        try:
            assert np.array_equal(compact.query_pml(bases, offsets, rpml=True), want)
        finally:
            compact.close()
    dev = torch.device("cuda:0")
    pad = np.zeros(64, np.uint8)
    d_b = torch.from_numpy(np.concatenate([bases, pad])).to(dev)
    d_o = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    d_out = torch.zeros(len(bases) + 64, dtype=torch.int16, device=dev)
    ms.query_pml_device(d_b.data_ptr(), d_o.data_ptr(), len(offsets) - 1, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream, rpml=True)
    torch.cuda.synchronize()
    ms.check_device_status()
    assert np.array_equal(d_out[: len(bases)].cpu().numpy().view(np.uint16), want)
    # the fused-classify and byte-wide entry points take the flag as well
    from oracle.port import classify
    a, b, s = ms.query_classify(bases, offsets, 150, 7, rpml=True)
    for i in range(0, len(offsets) - 1, 211):
        v = want[int(offsets[i]):int(offsets[i + 1])]
        if len(v):
            _, ea, eb, es, _ = classify(v, 150, 7)
            assert (a[i], b[i], s[i]) == (ea, eb, es)


@pytest.mark.parametrize("sep", [False, True])
@pytest.mark.parametrize("typ", ["regular", "blocked", "sampled"])
def test_index_types_without_thresholds(nothr_index_dirs, read_sets, ref_seq, typ, sep):
    """The non-threshold index types (MODE 3 "regular": 12-bit run lengths; MODE 2 "blocked": id_blocks + two more id bits;
    MODE 5 "sampled": id checkpoints, 10-bit lengths, 4-bit character): PML (which repositions with reposition_randomly there),
    ZML and count against the unmodified reference's output and, on 20 k ragged random reads, against the oracle; host-buffer,
    byte-wide and fused-classify entry points; for "regular" the rows export back to the file's bytes."""
    import ctypes as C
    from movi_b200 import HostIndex, MoveStructure, MoviGpuError
    from oracle.port import OracleIndex, classify, pack_reads
    d = nothr_index_dirs[(typ, sep)]
    ms = MoveStructure(d)
    try:
        assert ms.host_index.type_name == typ
        rd0 = read_sets["sample.fastq"]
        l0 = ms.launch_count
        ms.query_pml(rd0.bases, rd0.offsets)
        assert ms.launch_count == l0 + 1          # exactly one kernel per PML call (one pipeline slice)
        for read_set in READ_SETS:
            rd = read_sets[read_set]
            assert ml_text(rd.ids(), rd.offsets, ms.query_pml(rd.bases, rd.offsets)) == expected_text("%s.%s" % (read_set, typ), "pml", sep)
            if typ != "sampled":
                assert ml_text(rd.ids(), rd.offsets, ms.query_zml(rd.bases, rd.offsets)) == expected_text(read_set, "zml", sep)
                c, p = ms.query_backward_search(rd.bases, rd.offsets)
                assert count_text(rd.ids(), rd.offsets, c, p) == expected_text(read_set, "count", sep)
        orc = OracleIndex(d)
        bases, offsets = pack_reads(_random_reads(17, 20000, 600, ref_seq))
        want = orc.query_pml(bases, offsets, threads=0)
        assert np.array_equal(ms.query_pml(bases, offsets), want)
        assert np.array_equal(ms.query_pml(bases, offsets, rpml=True), want)          # the flag changes nothing here
        if typ != "sampled":
            assert np.array_equal(ms.query_zml(bases, offsets), orc.query_zml(bases, offsets, threads=0))
            c1, p1 = ms.query_backward_search(bases, offsets)
            c2, p2 = orc.query_count(bases, offsets, threads=0)
            assert np.array_equal(c1, c2) and np.array_equal(p1, p2)
        else:
            with pytest.raises(MoviGpuError, match="not suppported by the sampled modes"):
                ms.query_zml(bases, offsets)
        short = pack_reads(_random_reads(19, 4000, 255, ref_seq))
        assert np.array_equal(ms.query_ml8(*short), orc.query_pml(*short, threads=0).astype(np.uint8))
        a, b, s = ms.query_classify(bases, offsets, 150, 7)
        for i in range(0, len(offsets) - 1, 97):
            v = want[int(offsets[i]):int(offsets[i + 1])]
            if len(v):
                _, ea, eb, es, _ = classify(v, 150, 7)
                assert (a[i], b[i], s[i]) == (ea, eb, es)
        if typ == "regular":
            hix = HostIndex(d)
            file_rows = np.ctypeslib.as_array(C.cast(hix.desc.rows, C.POINTER(C.c_uint32)), shape=(hix.r, 2)).copy()
            got = ms.export_rows()
            mask = ~np.uint32(1 << 12)                 # bit 12 of the n field is not used by the format (LENGTH_BITS 12, SHIFT_C 13)
            assert np.array_equal(got[:, 0], file_rows[:, 0]) and np.array_equal(got[:, 1] & mask, file_rows[:, 1] & mask)
    finally:
        ms.close()


@pytest.mark.parametrize("typ", ["regular-thresholds", "sampled-thresholds"])
def test_device_side_bpf_records(structures, read_sets, ref_seq, tmp_path, typ):
    """movi_gpu_query_bpf: the .bpf records {u16 id_len; id; u64 L; u16 ml[L]} assembled on the device equal, byte for byte, what
    the host writer produces from the same matching lengths (= the reference's file, tests/test_cli_gpu.py) -- every read set
    (odd and even id lengths, 1-base reads), one slice and many tiny slices, PML and ZML, and the device-only entry point."""
    import torch
    from movi_b200 import _native as N
    ms = structures[(typ, False)]
    H = N.host_lib()
    old = ms.get_option("batch_bases")
    for read_set in READ_SETS:
        rd = read_sets[read_set]
        ids, id_off = rd.id_blob
        for zml in ((False,) if typ == "sampled-thresholds" else (False, True)):
            ml = ms.query_zml(rd.bases, rd.offsets) if zml else ms.query_pml(rd.bases, rd.offsets)
            f = str(tmp_path / ("%s.%d.bpf" % (read_set, zml)))
            assert H.movi_host_write_bpf(f.encode(), rd._h, ml.ctypes.data) == 0
            want = open(f, "rb").read()[12:]
            for batch in (old, 65536):
                ms.set_option("batch_bases", batch)
                try:
                    got = ms.query_bpf(rd.bases, rd.offsets, ids, id_off, zml=zml)
                finally:
                    ms.set_option("batch_bases", old)
                assert got.tobytes() == want, (read_set, zml, batch)
    # device-only form on torch buffers
    rd = read_sets["synth300.fasta"]
    ids, id_off = rd.id_blob
    dev = torch.device("cuda:0")
    ml = ms.query_pml(rd.bases, rd.offsets)
    d_ml = torch.from_numpy(np.concatenate([ml, np.zeros(64, np.uint16)]).view(np.int16)).to(dev)
    d_off = torch.from_numpy(rd.offsets.astype(np.int64)).to(dev)
    d_ids = torch.from_numpy(np.concatenate([ids, np.zeros(16, np.uint8)])).to(dev)
    d_ido = torch.from_numpy(id_off.astype(np.int64)).to(dev)
    size = 10 * rd.count + int(id_off[-1]) + 2 * int(rd.offsets[-1])
    d_img = torch.zeros(size + 16, dtype=torch.uint8, device=dev)
    ms._check(ms._L.movi_gpu_bpf_assemble_device(ms._ctx, d_ml.data_ptr(), d_off.data_ptr(), rd.count, d_ids.data_ptr(), d_ido.data_ptr(),
                                                 d_img.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    f = str(tmp_path / "dev.bpf")
    assert H.movi_host_write_bpf(f.encode(), rd._h, ml.ctypes.data) == 0
    assert d_img[:size].cpu().numpy().tobytes() == open(f, "rb").read()[12:]


def _variants(text):
    """The same records in other dresses: CRLF line ends, no final newline, sequences folded to 37 columns (FASTA only)."""
    out = {"as is": text, "crlf": text.replace(b"\n", b"\r\n"), "no final newline": text.rstrip(b"\n")}
    if text[:1] == b">":
        folded = []
        for rec in text.split(b">")[1:]:
            head, _, seq = rec.partition(b"\n")
            seq = seq.replace(b"\n", b"")
            folded.append(b">" + head + b"\n" + b"".join(seq[i:i + 37] + b" \t\n" for i in range(0, len(seq), 37)))
        out["folded"] = b"".join(folded)
    return out


@pytest.mark.parametrize("read_set", READ_SETS)
def test_device_side_ingest_matches_the_host_parser(structures, read_set, tmp_path):
    """movi_gpu_parse_chunk (FASTA / FASTQ parsed on the device) against movi_host_reads_load (the restatement of the reference's
    BatchLoader): whole files, and files cut into two chunks at many byte positions -- the first chunk leaves its last record to
    the second one."""
    from movi_b200 import Reads
    ms = structures[("regular-thresholds", False)]
    text = open(reads_path(read_set), "rb").read()
    fastq = text[:1] == b"@"
    for name, t in _variants(text).items():
        f = tmp_path / "v.fa"
        f.write_bytes(t)
        want = Reads(str(f))
        wids, wido = want.id_blob
        bases, offsets, ids, ido, consumed, flags = ms.parse_chunk(t, fastq, final=True)
        assert flags == 0, name
        assert np.array_equal(offsets, want.offsets) and np.array_equal(bases, want.bases), name
        assert np.array_equal(ido, wido) and np.array_equal(ids, wids), name
        rng = np.random.default_rng(len(t))
        cuts = sorted(set(int(c) for c in rng.integers(0, len(t), 24)) | {0, 1, len(t) - 1, len(t)})
        for c in cuts:
            b1, o1, i1, io1, used, fl = ms.parse_chunk(t[:c], fastq, final=False)
            assert fl == 0 and used <= c
            b2, o2, i2, io2, used2, fl2 = ms.parse_chunk(t[used:], fastq, final=True)
            assert fl2 == 0, (name, c)
            n1 = len(o1) - 1 if len(o1) else 0
            got_b = np.concatenate([b1, b2])
            got_o = np.concatenate([o1[:n1], o2 + (o1[n1] if len(o1) else 0)]) if len(o2) else o1
            got_i = np.concatenate([i1, i2])
            got_io = np.concatenate([io1[:n1], io2 + (io1[n1] if len(io1) else 0)]) if len(io2) else io1
            assert np.array_equal(got_o, want.offsets) and np.array_equal(got_b, want.bases), (name, c)
            assert np.array_equal(got_io, wido) and np.array_equal(got_i, wids), (name, c)


def test_device_side_ingest_flags_what_the_reference_rejects(structures):
    ms = structures[("regular-thresholds", False)]
    assert ms.parse_chunk(b"ACGT\n>a\nAC\n", False)[5] & 1                 # does not start with '>'
    assert ms.parse_chunk(b"@a\nAC\n+\nII\nxx\nAC\n+\nII\n", True)[5] & 1    # FASTQ record not starting with '@'
    assert ms.parse_chunk(b">a\nAC\n>\nGG\n", False)[5] & 2                # header without an id
    assert ms.parse_chunk(b"@a\nAC\n+\nII\n\n", True)[5] & 4               # empty line where a header is due
    assert ms.parse_chunk(b"", False) [5] == 0


def test_file_bytes_to_bpf_records_on_the_device(structures, read_sets, tmp_path):
    """movi_gpu_query_file_chunk: raw file bytes in, .bpf records out, nothing but two copies on the host path."""
    from movi_b200 import _native as N
    H = N.host_lib()
    for typ in ("regular-thresholds", "blocked-thresholds"):
        ms = structures[(typ, False)]
        for read_set in READ_SETS:
            rd = read_sets[read_set]
            text = open(reads_path(read_set), "rb").read()
            for zml in (False, True):
                ml = ms.query_zml(rd.bases, rd.offsets) if zml else ms.query_pml(rd.bases, rd.offsets)
                f = str(tmp_path / "w.bpf")
                assert H.movi_host_write_bpf(f.encode(), rd._h, ml.ctypes.data) == 0
                want = open(f, "rb").read()[12:]
                image, res = ms.query_file_chunk(text, text[:1] == b"@", final=True, zml=zml)
                assert res.parse_flags == 0 and res.n_reads == rd.count and res.consumed_bytes == len(text)
                assert image.tobytes() == want, (typ, read_set, zml)
                # two chunks
                image1, r1 = ms.query_file_chunk(text[: len(text) // 2], text[:1] == b"@", final=False, zml=zml)
                image2, r2 = ms.query_file_chunk(text[r1.consumed_bytes:], text[:1] == b"@", final=True, zml=zml)
                assert image1.tobytes() + image2.tobytes() == want


def test_two_contexts_share_one_uploaded_index(structures, read_sets, ref_seq):
    """movi_gpu_share_index: a second context on the same GPU walks the table the first one uploaded (what lets the command line keep
    two file chunks in flight per GPU); closing the borrower leaves the owner's index alone."""
    import threading
    from movi_b200 import MoveStructure
    from oracle.port import pack_reads
    for key in (("regular-thresholds", False), ("blocked-thresholds", True)):
        owner = structures[key]
        bases, offsets = pack_reads(_random_reads(29, 8000, 400, ref_seq))
        want = owner.query_pml(bases, offsets)
        want_c = owner.query_backward_search(bases, offsets)
        other = MoveStructure.sharing(owner)
        try:
            assert other.index_device_bytes == 0
            got = {}
            t = threading.Thread(target=lambda: got.update(a=other.query_pml(bases, offsets), c=other.query_backward_search(bases, offsets)))
            t.start()
            mine = owner.query_pml(bases, offsets)          # both contexts at once
            t.join()
            assert np.array_equal(got["a"], want) and np.array_equal(mine, want)
            assert np.array_equal(got["c"][0], want_c[0]) and np.array_equal(got["c"][1], want_c[1])
        finally:
            other.close()
        assert np.array_equal(owner.query_pml(bases, offsets), want)
