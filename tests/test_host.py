"""Host-side helpers (libmovi_host.so, pure C++) against the oracle and the reference outputs -- CPU only.
Also: both native libraries load and export every symbol their headers declare."""
import lzma
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLD, READ_SETS, ROOT, TYPES, expected_text, reads_path
from movi_b200 import HostIndex, Reads
from movi_b200 import _native as N
from oracle.port import OracleIndex


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(movi_(?:gpu|host)_[a-z0-9_]+)\s*\(", txt)))


def test_gpu_library_exports_every_declared_symbol():
    """The C ABI library loads without a GPU and exports all of include/movi_gpu.h (no compute calls here)."""
    lib = N.gpu_lib()
    names = _declared("movi_gpu.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in N.GPU_SYMBOLS, "ctypes signature missing for " + n
    assert lib.movi_gpu_abi_version() == 1


def test_host_library_exports_every_declared_symbol():
    lib = N.host_lib()
    for n in _declared("movi_host.h"):
        assert hasattr(lib, n), n
        assert n in N.HOST_SYMBOLS, n


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from movi_b200 import MoveStructure, MoviGpuError
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "index.movi"), "wb").write(lzma.open(os.path.join(GOLD, "c1", "regular-thresholds", "index.movi.xz")).read())
        with pytest.raises(MoviGpuError, match="no CUDA device"):
            MoveStructure(td)


@pytest.mark.parametrize("sep", [False, True])
@pytest.mark.parametrize("typ", TYPES)
def test_index_reader_matches_oracle_reader(index_dirs, typ, sep):
    d = index_dirs[(typ, sep)]
    h, o = HostIndex(d), OracleIndex(d)
    assert (h.mode, h.r, h.n, h.end_bwt_idx) == (o.mode, o.r, o.n, o.end_bwt_idx)
    assert h.type_name == typ
    assert h.alphabet == (b"%ACGT" if sep else b"ACGT")
    assert (h.desc.sep_count > 0) == sep
    if typ == "blocked-thresholds":
        assert h.desc.block_size == 1048576 and h.desc.id_blocks_size == 1
    if typ == "sampled-thresholds":
        assert h.desc.tally_checkpoints == 20 and h.desc.tally_len == h.r // 20 + 2


def test_index_reader_rejects_garbage(tmp_path):
    p = tmp_path / "index.movi"
    p.write_bytes(b"\0" * 4096)
    with pytest.raises(RuntimeError, match="magic"):
        HostIndex(str(tmp_path))
    with pytest.raises(RuntimeError, match="Failed to open"):
        HostIndex(str(tmp_path / "nope"))


def test_reads_loader_ids_and_sequences():
    """id = header from its 2nd character through and including the first blank/tab (src/batch_loader.cpp:117-119)."""
    rd = Reads(reads_path("synth300.fasta"))
    assert rd.count == 300
    assert rd.id(0) == "syn.0 " and rd.id(1) == "syn.1\t" and rd.id(2) == "syn.2"
    # multi-line FASTA is concatenated
    txt = open(reads_path("synth300.fasta")).read().split(">")[1:]
    for i in (0, 1, 2, 150, 299):
        assert rd.read(i) == "".join(txt[i].split("\n")[1:])
    fq = Reads(reads_path("sample.fastq"))
    assert fq.count == 25 and fq.id(0) == "simulated.1" and len(fq.read(0)) == 150
    rev = Reads(reads_path("sample.fastq"), reverse=True)
    assert rev.read(3) == fq.read(3)[::-1]


def test_reads_loader_inflates_gzip(tmp_path):
    """The reference reads its query files through zlib (open_kseq, src/utils.cpp:69-80): .gz inputs, multi-member too."""
    import gzip
    for name in ("sample.fastq", "synth300.fasta"):
        raw = open(reads_path(name), "rb").read()
        plain = Reads(reads_path(name))
        one = tmp_path / (name + ".gz")
        one.write_bytes(gzip.compress(raw))
        half = raw.rfind(b"\n@" if name.endswith("q") else b"\n>", 0, len(raw) // 2) + 1
        two = tmp_path / (name + ".2.gz")
        two.write_bytes(gzip.compress(raw[:half]) + gzip.compress(raw[half:]))
        for f in (one, two):
            z = Reads(str(f))
            assert z.count == plain.count and z.ids() == plain.ids()
            assert np.array_equal(z.bases, plain.bases) and np.array_equal(z.offsets, plain.offsets)
    bad = tmp_path / "bad.fa.gz"
    bad.write_bytes(b"\x1f\x8b" + b"garbage" * 10)
    with pytest.raises(RuntimeError):
        Reads(str(bad))


def test_batch_reader_equals_whole_file_loader():
    """movi_host_reader_*: the same records as movi_host_reads_load, whatever the batch size (the CLI streams with it)."""
    import ctypes as C
    L = N.host_lib()
    for name in ("synth300.fasta", "sample.fastq", "edge.fasta"):
        whole = Reads(reads_path(name))
        for max_bases in (1, 700, 10**9):
            h = C.c_void_p()
            assert L.movi_host_reader_open(os.fsencode(reads_path(name)), 0, C.byref(h)) == 0
            ids, seqs, batches = [], [], 0
            while True:
                b = C.c_void_p()
                assert L.movi_host_reader_next(h, max_bases, C.byref(b)) == 0
                if not b.value:
                    break
                batches += 1
                n = L.movi_host_reads_count(b)
                off = np.ctypeslib.as_array(C.cast(L.movi_host_reads_offsets(b), N.u64p), shape=(n + 1,))
                data = C.string_at(L.movi_host_reads_bases(b), int(off[-1]))
                for i in range(n):
                    ln = C.c_uint32()
                    p = L.movi_host_reads_id(b, i, C.byref(ln))
                    ids.append(C.string_at(p, ln.value).decode())
                    seqs.append(data[int(off[i]):int(off[i + 1])].decode())
                L.movi_host_reads_free(b)
            L.movi_host_reader_close(h)
            assert ids == whole.ids() and seqs == [whole.read(i) for i in range(whole.count)]
            assert batches == 1 if max_bases == 10**9 else batches > 1


def _drain(path, max_bases):
    import ctypes as C
    L = N.host_lib()
    h = C.c_void_p()
    assert L.movi_host_reader_open(os.fsencode(path), 0, C.byref(h)) == 0
    ids, seqs, batches = [], [], 0
    while True:
        b = C.c_void_p()
        assert L.movi_host_reader_next(h, max_bases, C.byref(b)) == 0, L.movi_host_last_error()
        if not b.value:
            break
        batches += 1
        n = L.movi_host_reads_count(b)
        off = np.ctypeslib.as_array(C.cast(L.movi_host_reads_offsets(b), N.u64p), shape=(n + 1,)).copy()
        data = C.string_at(L.movi_host_reads_bases(b), int(off[-1]))
        for i in range(n):
            ln = C.c_uint32()
            p = L.movi_host_reads_id(b, i, C.byref(ln))
            ids.append(C.string_at(p, ln.value))
            seqs.append(data[int(off[i]):int(off[i + 1])])
        L.movi_host_reads_free(b)
    L.movi_host_reader_close(h)
    return ids, seqs, batches


def test_big_files_one_batch_or_many(tmp_path):
    """20 MB FASTQ (quality lines that start with '@' and '>') and multi-line FASTA: one big batch, many small batches and a
    Python parse give the same records.  (A multi-threaded variant of the parser was measured on the B200 box: 0.076 s against
    0.088 s single-threaded for 161 MB -- memory bound, not kept.)"""
    rng = np.random.default_rng(5)
    n = 60000
    lens = rng.integers(1, 400, n)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), int(l))) for l in lens]
    fq = tmp_path / "big.fastq"
    with open(fq, "wb") as f:
        for i, s_ in enumerate(seqs):
            q = (b"@" if i % 3 == 0 else b">" if i % 3 == 1 else b"I") + b"I" * (len(s_) - 1)
            f.write(b"@r%d some text\n" % i + s_ + b"\n+\n" + q[:len(s_)] + b"\n")
    fa = tmp_path / "big.fasta"
    with open(fa, "wb") as f:
        for i, s_ in enumerate(seqs):
            f.write(b">r%d\tx\n" % i + b"\n".join(s_[k:k + 61] for k in range(0, len(s_), 61)) + b"\n")
    assert os.path.getsize(fq) > (20 << 20) and os.path.getsize(fa) > (9 << 20)
    for path, sep in ((fq, b" "), (fa, b"\t")):
        ids_p, seqs_p, nb_p = _drain(str(path), 10**12)          # one batch
        ids_s, seqs_s, nb_s = _drain(str(path), 1 << 20)         # many small batches
        assert nb_p == 1 and nb_s > 5
        assert ids_p == ids_s == [b"r%d" % i + sep for i in range(n)]
        assert seqs_p == seqs_s == seqs
    # a truncated last FASTQ record is dropped
    data = open(fq, "rb").read()
    cut = tmp_path / "cut.fastq"
    cut.write_bytes(data[:data.rfind(b"\n+\n") + 3])
    ids_c, seqs_c, _ = _drain(str(cut), 10**12)
    assert len(ids_c) == n - 1 and seqs_c == seqs[:-1]


def _write_via(fn, *args):
    import tempfile
    with tempfile.TemporaryFile() as f:
        assert fn(f.fileno(), *args) == 0, N.host_lib().movi_host_last_error()
        f.seek(0)
        return f.read().decode()


@pytest.mark.parametrize("read_set", READ_SETS)
def test_text_writers_reproduce_reference_stdout(index_dirs, read_set):
    """oracle values -> host writers == the reference's --stdout bytes (src/utils.cpp:204-209, 248-256)."""
    L = N.host_lib()
    rd = Reads(reads_path(read_set))
    orc = OracleIndex(index_dirs[("regular-thresholds", False)])
    pml = orc.query_pml(rd.bases, rd.offsets)
    assert _write_via(L.movi_host_write_ml_text, rd._h, pml.ctypes.data) == expected_text(read_set, "pml")
    counts, pos = orc.query_count(rd.bases, rd.offsets)
    assert _write_via(L.movi_host_write_counts, rd._h, counts.ctypes.data, pos.ctypes.data) == expected_text(read_set, "count")


def test_bpf_writer_is_byte_identical_to_reference(index_dirs, tmp_path):
    """<reads>.regular-thresholds.pml.bpf written by the reference for sample.fasta (header padding bytes excepted: 10, 11)."""
    L = N.host_lib()
    rd = Reads(reads_path("sample.fasta"))
    pml = OracleIndex(index_dirs[("regular-thresholds", False)]).query_pml(rd.bases, rd.offsets)
    out = tmp_path / "x.bpf"
    assert L.movi_host_write_bpf(str(out).encode(), rd._h, pml.ctypes.data) == 0
    got = out.read_bytes()
    want = lzma.open(os.path.join(GOLD, "expected", "sample.fasta.pml.bpf.xz")).read()
    assert len(got) == len(want) and got[:10] == want[:10] and got[12:] == want[12:]


@pytest.mark.parametrize("read_set", ["long6.fasta", "sample.fasta", "synth300.fasta"])
def test_classifier_report_and_filter_match_reference(index_dirs, read_set):
    """--classify report (incl. header and number formatting) and --filter / --filter -v outputs."""
    L = N.host_lib()
    d = index_dirs[("regular-thresholds", False)]
    thr = N.C.c_uint64()
    assert L.movi_host_nulldb_threshold(d.encode(), b"pml", N.C.byref(thr)) == 0
    assert thr.value == 7                        # "thr=7", tests_data/reads.fasta.pmls.report.sorted
    rd = Reads(reads_path(read_set))
    pml = OracleIndex(d).query_pml(rd.bases, rd.offsets)
    report = _write_via(L.movi_host_classify, 0, rd._h, pml.ctypes.data, 150, thr.value, None)
    assert report == open(os.path.join(GOLD, "expected", read_set + ".pml.report")).read()
    for mode, name in ((1, "filtered"), (2, "filtered_notfound")):
        got = _write_via(L.movi_host_classify, mode, rd._h, pml.ctypes.data, 150, thr.value, None)
        assert got == lzma.open(os.path.join(GOLD, "expected", "%s.pml.%s.xz" % (read_set, name))).read().decode()


def test_cli_fails_loudly_without_gpu(index_dirs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from movi_b200 import build
    build.build_host()
    r = subprocess.run([build.CLI, "query", "-i", index_dirs[("regular-thresholds", False)], "-r", reads_path("sample.fastq"), "--pml", "--stdout"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and r.stdout == ""


def test_index_reader_opens_the_types_without_thresholds(nothr_index_dirs):
    from oracle.port import OracleIndex
    for typ, mode, row_bytes in (("regular", 3, 8), ("blocked", 2, 6), ("sampled", 5, 3)):
        for sep in (False, True):
            hix = HostIndex(nothr_index_dirs[(typ, sep)])
            orc = OracleIndex(nothr_index_dirs[(typ, sep)])
            assert hix.type_name == typ and hix.mode == mode and hix.r == orc.r
            assert hix.alphabet == (b"%ACGT" if sep else b"ACGT")
            size = os.path.getsize(os.path.join(nothr_index_dirs[(typ, sep)], "index.movi"))
            assert hix.r * row_bytes < size


def test_index_reader_survives_corrupt_sizes(index_dirs, tmp_path):
    """Sizes read from a file are not trusted: a row count / table length that would overflow or run past the file is a clean
    error, not an out-of-range read."""
    import struct
    for typ in TYPES:
        raw = bytearray(open(os.path.join(index_dirs[(typ, False)], "index.movi"), "rb").read())
        for r_bad in (1 << 61, (1 << 64) - 1, len(raw)):
            bad = bytearray(raw)
            struct.pack_into("<Q", bad, 24, r_bad)            # MoviHeader.r
            d = tmp_path / ("%s_%d" % (typ, r_bad % 1000))
            d.mkdir(exist_ok=True)
            (d / "index.movi").write_bytes(bad)
            with pytest.raises(RuntimeError, match="truncated|inconsistent|unexpected"):
                HostIndex(str(d))
        cut = tmp_path / (typ + "_cut")
        cut.mkdir()
        (cut / "index.movi").write_bytes(raw[: len(raw) // 2])
        with pytest.raises(RuntimeError, match="truncated|unexpected"):
            HostIndex(str(cut))


def test_gzip_with_trailing_garbage_is_read_like_gzread(tmp_path):
    """Bytes after a complete gzip member that are not another member end the input (the reference's gzopen / gzread does the same)."""
    import gzip
    raw = open(reads_path("synth300.fasta"), "rb").read()
    plain = Reads(reads_path("synth300.fasta"))
    f = tmp_path / "padded.fa.gz"
    f.write_bytes(gzip.compress(raw) + b"\0" * 512 + b"trailing junk")
    z = Reads(str(f))
    assert z.count == plain.count and np.array_equal(z.bases, plain.bases)


def test_record_boundary_finder():
    """movi_host_record_boundary: where the device-side ingest may cut a file (FASTA: a line starting with '>'; FASTQ: '@' line whose
    next-but-one line starts with '+', so that a quality line starting with '@' is not taken for a header)."""
    import ctypes as C
    H = N.host_lib()
    fa = b">a\nACGT\nAC\n>b x\nGG\n>c\nT\n"
    starts = [0, 11, 19]
    for frm in range(len(fa) + 1):
        want = min([s for s in starts if s >= frm] + [len(fa)])
        assert H.movi_host_record_boundary(fa, len(fa), frm, 0) == want, frm
    fq = b"@r1\nACGT\n+\n@@II\n@r2\nAC\n+\n@I\n@r3\nA\n+\nI\n"
    starts = [0, 16, 28]
    for frm in range(len(fq) + 1):
        want = min([s for s in starts if s >= frm] + [len(fq)])
        assert H.movi_host_record_boundary(fq, len(fq), frm, 1) == want, frm
