"""`movi-b200 query ...` end to end on the GPU against the reference's own stdout / files (tests/golden/expected)."""
import lzma
import os
import shutil
import subprocess

import pytest

from conftest import GOLD, READ_SETS, TYPES, expected_text, reads_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli():
    from movi_b200 import build
    build.build_host()
    return build.CLI


def run(cli, *args, cwd=None):
    r = subprocess.run([cli] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=cwd)
    assert r.returncode == 0, r.stderr.decode()[-1000:]
    return r.stdout.decode()


@pytest.mark.parametrize("typ", TYPES)
@pytest.mark.parametrize("read_set", READ_SETS)
def test_pml_stdout(cli, index_dirs, typ, read_set):
    # attached short-flag values as in the reference's tests (tests/test_pml.cpp:59-86)
    out = run(cli, "query", "--index", index_dirs[(typ, False)], "--read", reads_path(read_set), "--pml", "--no-prefetch", "-t1", "--stdout")
    assert out == expected_text(read_set, "pml")


@pytest.mark.parametrize("q", ["zml", "count"])
@pytest.mark.parametrize("sep", [False, True])
def test_zml_count_stdout(cli, index_dirs, q, sep):
    for read_set in ("edge.fasta", "synth300.fasta"):
        out = run(cli, "query", "-i", index_dirs[("blocked-thresholds", sep)], "-r", reads_path(read_set), "--" + q, "-s16", "-t", "4", "--stdout")
        assert out == expected_text(read_set, q, sep)


def test_sampled_rejects_count(cli, index_dirs):
    r = subprocess.run([cli, "query", "-i", index_dirs[("sampled-thresholds", False)], "-r", reads_path("sample.fasta"), "--count", "--stdout"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"not suppported by the sampled modes" in r.stderr


def test_classify_report_filter_and_bpf(cli, index_dirs, tmp_path):
    d = index_dirs[("regular-thresholds", False)]
    for read_set in ("long6.fasta", "sample.fasta"):
        reads = tmp_path / read_set
        shutil.copy(reads_path(read_set), reads)
        # --classify --stdout: report on stdout AND the .bpf file next to the reads
        out = run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--classify", "--stdout")
        assert out == open(os.path.join(GOLD, "expected", read_set + ".pml.report")).read()
        bpf = str(reads) + ".regular-thresholds.pml.bpf"
        assert os.path.exists(bpf)
        if read_set == "sample.fasta":
            got = open(bpf, "rb").read()
            want = lzma.open(os.path.join(GOLD, "expected", "sample.fasta.pml.bpf.xz")).read()
            assert got[:10] == want[:10] and got[12:] == want[12:]
            # view prints the lengths back in read order
            assert run(cli, "view", "--bpf", bpf) == expected_text(read_set, "pml")
        os.remove(bpf)
        # --classify without --stdout: report file, named after the reads and the index type
        run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--classify")
        rep = str(reads) + ".regular-thresholds.pml.report"
        assert open(rep).read() == open(os.path.join(GOLD, "expected", read_set + ".pml.report")).read()
        for inv, name in ((False, "filtered"), (True, "filtered_notfound")):
            out = run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--filter", *(["-v"] if inv else []))
            assert out == lzma.open(os.path.join(GOLD, "expected", "%s.pml.%s.xz" % (read_set, name))).read().decode()


def test_out_prefix_count_file_and_stdin(cli, index_dirs, tmp_path):
    d = index_dirs[("regular-thresholds", False)]
    prefix = str(tmp_path / "res")
    run(cli, "query", "-i", d, "-r", reads_path("synth300.fasta"), "--count", "-o", prefix)
    assert open(prefix + ".count.matches").read() == expected_text("synth300.fasta", "count")
    r = subprocess.run([cli, "query", "-i", d, "-r", "-", "--zml", "--stdout"], input=open(reads_path("sample.fastq"), "rb").read(),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and r.stdout.decode() == expected_text("sample.fastq", "zml")


def test_reverse_and_ignore_illegal(cli, index_dirs):
    """--reverse reverses (not complements) each read; --ignore-illegal-chars 1 substitutes 'A' before the walk."""
    import numpy as np
    from movi_b200 import Reads
    from oracle.port import OracleIndex, pack_reads
    from _fmt import ml_text
    d = index_dirs[("regular-thresholds", False)]
    rd = Reads(reads_path("edge.fasta"))
    orc = OracleIndex(d)
    rev = [rd.read(i)[::-1] for i in range(rd.count)]
    b, o = pack_reads(rev)
    assert run(cli, "query", "-i", d, "-r", reads_path("edge.fasta"), "--pml", "--reverse", "--stdout") == ml_text(rd.ids(), o, orc.query_pml(b, o))
    sub = ["".join(c if c in "ACGT" else "A" for c in rd.read(i)) for i in range(rd.count)]
    b, o = pack_reads(sub)
    assert run(cli, "query", "-i", d, "-r", reads_path("edge.fasta"), "--pml", "--ignore-illegal-chars", "1", "--stdout") == ml_text(rd.ids(), o, orc.query_pml(b, o))


def test_streaming_batches_give_the_same_files(cli, index_dirs, tmp_path):
    """--chunk-bases: the read file goes through in many small batches (parse-ahead, appended outputs, two GPUs' worth of
    sharding logic on one device); every output must be byte-identical to the single-batch run."""
    d = index_dirs[("regular-thresholds", False)]
    for read_set, q in (("synth300.fasta", "pml"), ("synth300.fasta", "zml"), ("synth300.fasta", "count"), ("sample.fastq", "pml")):
        out = run(cli, "query", "-i", d, "-r", reads_path(read_set), "--" + q, "--stdout", "--chunk-bases", "3000")
        assert out == expected_text(read_set, q)
    reads = tmp_path / "long6.fasta"
    shutil.copy(reads_path("long6.fasta"), reads)
    out = run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--classify", "--stdout", "--chunk-bases", "1")    # one read per batch
    assert out == open(os.path.join(GOLD, "expected", "long6.fasta.pml.report")).read()
    reads = tmp_path / "sample.fasta"
    shutil.copy(reads_path("sample.fasta"), reads)
    run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--chunk-bases", "500")
    got = open(str(reads) + ".regular-thresholds.pml.bpf", "rb").read()
    want = lzma.open(os.path.join(GOLD, "expected", "sample.fasta.pml.bpf.xz")).read()
    assert got[:10] == want[:10] and got[12:] == want[12:]
    for inv, name in ((False, "filtered"), (True, "filtered_notfound")):
        out = run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--filter", "--chunk-bases", "700", *(["-v"] if inv else []))
        assert out == lzma.open(os.path.join(GOLD, "expected", "sample.fasta.pml.%s.xz" % name)).read().decode()
    # an empty read file still yields its (empty) outputs
    empty = tmp_path / "none.fasta"
    empty.write_bytes(b"")
    assert run(cli, "query", "-i", d, "-r", str(empty), "--pml", "--stdout") == ""
    run(cli, "query", "-i", d, "-r", str(empty), "--pml")
    assert len(open(str(empty) + ".regular-thresholds.pml.bpf", "rb").read()) == 12


@pytest.mark.parametrize("typ", TYPES)
def test_rpml_stdout_and_file_names(cli, index_dirs, typ, tmp_path):
    """--rpml: the reference's output (tests/golden/make_golden_rpml.py) and its file naming (query_type() = "rpml")."""
    for read_set in ("sample.fastq", "edge.fasta"):
        out = run(cli, "query", "--index", index_dirs[(typ, False)], "--read", reads_path(read_set), "--pml", "--rpml", "--stdout")
        assert out == expected_text(read_set, "rpml")
    reads = tmp_path / "sample.fasta"
    shutil.copy(reads_path("sample.fasta"), reads)
    run(cli, "query", "-i", index_dirs[(typ, False)], "-r", str(reads), "--rpml")
    bpf = "%s.%s.rpml.bpf" % (reads, typ)
    assert os.path.exists(bpf)
    assert run(cli, "view", "--bpf", bpf) == expected_text("sample.fasta", "rpml")


@pytest.mark.parametrize("typ", ["regular", "blocked", "sampled"])
def test_index_types_without_thresholds(cli, nothr_index_dirs, tmp_path, typ):
    """The non-threshold index types through the command line: the reference's output and its file naming."""
    d = nothr_index_dirs[(typ, False)]
    for read_set in ("sample.fastq", "edge.fasta"):
        out = run(cli, "query", "--index", d, "--read", reads_path(read_set), "--pml", "--stdout")
        assert out == expected_text("%s.%s" % (read_set, typ), "pml")
        if typ != "sampled":
            out = run(cli, "query", "--index", d, "--read", reads_path(read_set), "--count", "--stdout")
            assert out == expected_text(read_set, "count")
    reads = tmp_path / "sample.fasta"
    shutil.copy(reads_path("sample.fasta"), reads)
    run(cli, "query", "-i", d, "-r", str(reads), "--pml")
    assert run(cli, "view", "--bpf", "%s.%s.pml.bpf" % (reads, typ)) == expected_text("sample.fasta." + typ, "pml")


def test_ftab_flags_are_accepted(cli, index_dirs):
    """--ftab-k seeds the backward search from a k-mer table in the reference: an accelerator with identical results (checked with
    the unmodified reference: `query --zml/--count --ftab-k 6` after `movi ftab --ftab-k 6` prints byte-identical output on
    synth300.fasta and edge.fasta).  The flags are accepted; no table is needed."""
    d = index_dirs[("blocked-thresholds", False)]
    for q in ("zml", "count"):
        out = run(cli, "query", "-i", d, "-r", reads_path("edge.fasta"), "--" + q, "--ftab-k", "6", "--multi-ftab", "--stdout")
        assert out == expected_text("edge.fasta", q)


@pytest.mark.parametrize("typ", ["regular-thresholds", "blocked-thresholds"])
def test_device_side_ingest_bpf_files(cli, index_dirs, tmp_path, typ):
    """`movi-b200 query --pml` whose only output is the .bpf file: the read file's bytes go to the GPU as they are (parse + walk +
    record assembly there).  Whatever the chunking, and with the host parser instead (--host-parse), the file is the same, and
    its lengths are the reference's."""
    d = index_dirs[(typ, False)]
    for read_set in READ_SETS:
        for q in ("pml", "zml"):
            files = {}
            for name, extra in (("device", []), ("tiny chunks", ["--chunk-bytes", "4096"]), ("host", ["--host-parse"])):
                reads = tmp_path / ("%s.%s" % (name.replace(" ", "_"), read_set))
                shutil.copy(reads_path(read_set), reads)
                run(cli, "query", "-i", d, "-r", str(reads), "--" + q, *extra)
                bpf = "%s.%s.%s.bpf" % (reads, typ, q)
                files[name] = open(bpf, "rb").read()
                if name == "device":
                    assert run(cli, "view", "--bpf", bpf) == expected_text(read_set, q)
            assert files["device"][:10] == files["host"][:10] and files["device"][12:] == files["host"][12:]
            assert files["device"] == files["tiny chunks"]
    # --no-output goes through the same path and writes nothing
    reads = tmp_path / "quiet.fasta"
    shutil.copy(reads_path("synth300.fasta"), reads)
    run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--no-output")
    assert not os.path.exists("%s.%s.pml.bpf" % (reads, typ))


def test_device_side_ingest_hands_odd_inputs_to_the_host_parser(cli, index_dirs, tmp_path):
    """An empty line where a FASTQ header is due ends the input in the reference; the device-side ingest flags it and the host
    parser takes over from the start: same file as with --host-parse."""
    d = index_dirs[("regular-thresholds", False)]
    text = open(reads_path("sample.fastq"), "rb").read()
    lines = text.split(b"\n")
    odd = b"\n".join(lines[:8]) + b"\n\n" + b"\n".join(lines[8:])
    out = {}
    for name, extra in (("device", []), ("host", ["--host-parse"])):
        reads = tmp_path / (name + ".fastq")
        reads.write_bytes(odd)
        run(cli, "query", "-i", d, "-r", str(reads), "--pml", *extra)
        out[name] = open("%s.regular-thresholds.pml.bpf" % reads, "rb").read()
    assert out["device"][12:] == out["host"][12:] and len(out["host"]) > 12


# ---- Movi Color flows: the command lines tests/golden/make_golden_color.py gave the unmodified reference
@pytest.mark.parametrize("gpus", [1, 2])
def test_multi_classify_and_sa_entries_files(cli, color_index_dir, tmp_path, gpus):
    import torch
    if gpus > torch.cuda.device_count():
        pytest.skip("needs %d GPUs" % gpus)
    from conftest import COLOR_VARIANTS, color_expected
    reads = tmp_path / "reads.fa"
    shutil.copy(os.path.join(GOLD, "color", "reads.fa"), reads)
    flags = {"": [], ".minlen3": ["--min-len", "3"], ".pvalue": ["--pvalue-scoring"], ".minlen0": ["--min-len", "0"],
             ".pvalue.minlen0": ["--pvalue-scoring", "--min-len", "0"]}
    want_bpf = lzma.open(os.path.join(GOLD, "color", "expected.pml.bpf.xz")).read()
    for tag in COLOR_VARIANTS:
        o = str(tmp_path / ("mc%s.csv" % tag))
        run(cli, "query", "--index", color_index_dir, "--read", str(reads), "--pml", "--multi-classify", "--no-prefetch", "-t1", "-o", o,
            "--gpus", str(gpus), "--chunk-reads", "64", "--chunk-bases", "20000", *flags[tag])       # several batches
        assert open(o).read() == color_expected(tag), tag
        got = open(o + ".pml.bpf", "rb").read()
        assert got[:10] == want_bpf[:10] and got[12:] == want_bpf[12:]
    o = str(tmp_path / "mc.vectors.csv")        # the vectors-of-vectors table (doc_sets.bin)
    run(cli, "query", "--index", color_index_dir, "--read", str(reads), "--pml", "--multi-classify", "--color-vectors", "-o", o, "--gpus", str(gpus), "--no-output")
    assert open(o).read() == open(os.path.join(GOLD, "color", "expected.multi_classify.vectors.csv")).read()
    assert not os.path.exists(o + ".pml.bpf")
    run(cli, "query", "--index", color_index_dir, "--read", str(reads), "--pml", "--sa-entries", "--no-prefetch", "-t1", "--gpus", str(gpus))
    assert open(str(reads) + ".regular-thresholds.pml.sa_entries.bpf", "rb").read() == lzma.open(os.path.join(GOLD, "color", "expected.sa_entries.bpf.xz")).read()
    got = open(str(reads) + ".regular-thresholds.pml.bpf", "rb").read()
    assert got[12:] == want_bpf[12:]


def test_multi_classify_argument_errors(cli, color_index_dir, index_dirs):
    reads = os.path.join(GOLD, "color", "reads.fa")
    r = subprocess.run([cli, "query", "-i", color_index_dir, "-r", reads, "--pml", "--multi-classify"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"Please include the out file for multi-classify." in r.stderr
    r = subprocess.run([cli, "query", "-i", index_dirs[("regular-thresholds", False)], "-r", reads, "--pml", "--multi-classify", "-o", "/tmp/x.csv"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"doc_sets_flat.bin" in r.stderr
    r = subprocess.run([cli, "query", "-i", index_dirs[("regular-thresholds", False)], "-r", reads, "--pml", "--sa-entries"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"build-SA" in r.stderr


# ---- device-side ingest for the classification flows, and the hand-over to the host parser in the middle of an input
def _fastq_of(n, seed, read_len=120):
    """n FASTQ records cut from the c1 reference (FOUND) or random (NOT_PRESENT), ids with and without descriptions."""
    import numpy as np
    rng = np.random.default_rng(seed)
    ref = "".join(l.strip() for l in lzma.open(os.path.join(GOLD, "ref.fasta.xz"), "rt") if not l.startswith(">")).upper()
    out = []
    for i in range(n):
        L = int(rng.integers(30, read_len * 4))
        if i % 3 == 2:
            s = "".join(rng.choice(list("ACGT"), L))
        else:
            p = int(rng.integers(0, len(ref) - L))
            s = ref[p:p + L]
        out.append("@q%d%s\n%s\n+\n%s\n" % (i, " d" if i % 2 else "", s, "I" * L))
    return out


@pytest.mark.parametrize("chunk", [None, "4096"])
def test_classify_flows_through_the_device_side_ingest(cli, index_dirs, tmp_path, chunk):
    """--classify (report + .bpf), --classify --stdout, --filter, --filter -v: file bytes to the GPU, bins back; same outputs as the host
    parser path (--host-parse), with one chunk and with many (several contexts take turns writing the report)."""
    d = index_dirs[("regular-thresholds", False)]
    recs = _fastq_of(400, seed=3)
    extra = ["--device-ingest"] + (["--chunk-bytes", chunk] if chunk else [])      # (--filter takes the host path unless asked)
    out = {}
    for name, more in (("device", extra), ("host", ["--host-parse"])):
        reads = tmp_path / (name + ".fastq")
        reads.write_text("".join(recs))
        run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--classify", *more)
        rep = open(str(reads) + ".regular-thresholds.pml.report").read()
        bpf = open(str(reads) + ".regular-thresholds.pml.bpf", "rb").read()
        os.remove(str(reads) + ".regular-thresholds.pml.bpf")
        so = run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--classify", "--stdout", *more)
        bpf2 = open(str(reads) + ".regular-thresholds.pml.bpf", "rb").read()
        flt = run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--filter", *more)
        inv = run(cli, "query", "-i", d, "-r", str(reads), "--pml", "--filter", "-v", *more)
        zrep = run(cli, "query", "-i", d, "-r", str(reads), "--zml", "--classify", "--stdout", *more)
        out[name] = (rep, bpf[12:], so, bpf2[12:], flt, inv, zrep)
    assert out["device"] == out["host"]
    rep, _, so, _, flt, inv, _ = out["device"]
    assert rep == so and rep.count(" FOUND ") > 100 and rep.count("NOT_PRESENT") > 100
    assert flt.count(">") + inv.count(">") == 400 and flt.count(">") == rep.count(" FOUND ")


@pytest.mark.parametrize("flow", ["bpf", "classify", "filter"])
def test_device_side_ingest_hands_over_in_the_middle_of_the_input(cli, index_dirs, tmp_path, flow):
    """Several chunks go through the GPU parser, then one holds an empty line where a header is due (the reference stops reading
    there): the host parser continues from that chunk's first byte, appending to the same outputs."""
    d = index_dirs[("regular-thresholds", False)]
    recs = _fastq_of(300, seed=4)
    odd = "".join(recs[:220]) + "\n" + "".join(recs[220:])          # records 0..219 are read, then the input ends
    flags = {"bpf": [], "classify": ["--classify", "--stdout"], "filter": ["--filter"]}[flow]
    out = {}
    for name, more in (("device", ["--device-ingest", "--chunk-bytes", "8192"]), ("host", ["--host-parse"])):
        reads = tmp_path / (name + ".fastq")
        reads.write_text(odd)
        so = run(cli, "query", "-i", d, "-r", str(reads), "--pml", *flags, *more)
        bpf = open(str(reads) + ".regular-thresholds.pml.bpf", "rb").read()[12:] if flow != "filter" else b""
        out[name] = (so, bpf)
    assert out["device"] == out["host"]
    if flow == "bpf":
        from conftest import parse_bpf_records
        assert len(parse_bpf_records(b"\0" * 12 + out["device"][1], 2)) == 220
    if flow == "classify":
        assert out["device"][0].count("\n") == 221                     # header + 220 rows
