import lzma
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

TYPES = ["regular-thresholds", "blocked-thresholds", "sampled-thresholds"]
READ_SETS = ["sample.fastq", "sample.fasta", "long6.fasta", "synth300.fasta", "edge.fasta"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with `-m gpu`; if someone runs everything on a CPU box, skip instead of failing
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def index_dirs(tmp_path_factory):
    """{(type, sep): directory holding index.movi + frozen nulldbs}, unpacked from tests/golden/c1."""
    base = tmp_path_factory.mktemp("c1")
    out = {}
    for typ in TYPES:
        for sep in (False, True):
            tag = typ + (".sep" if sep else "")
            d = base / tag
            d.mkdir()
            with lzma.open(os.path.join(GOLD, "c1", tag, "index.movi.xz")) as f:
                (d / "index.movi").write_bytes(f.read())
            nd = os.path.join(GOLD, "c1", "regular-thresholds" + (".sep" if sep else ""))
            for q in ("pml", "zml"):
                with lzma.open(os.path.join(nd, "movi.%s.nulldb.xz" % q)) as f:
                    (d / ("movi.%s.nulldb" % q)).write_bytes(f.read())
            out[(typ, sep)] = str(d)
    return out


NOTHR_TYPES = ["regular", "blocked", "sampled"]      # MODE 3 / 2 / 5: no thresholds, PML repositions with reposition_randomly


@pytest.fixture(scope="session")
def nothr_index_dirs(tmp_path_factory):
    """{(type, sep): directory} of the non-threshold index types, built by the unmodified reference
    (tests/golden/make_golden_nothresholds.py)."""
    out = {}
    for typ in NOTHR_TYPES:
        for sep in (False, True):
            tag = typ + (".sep" if sep else "")
            d = tmp_path_factory.mktemp("c1_" + tag)
            with lzma.open(os.path.join(GOLD, "c1", tag, "index.movi.xz")) as f:
                (d / "index.movi").write_bytes(f.read())
            out[(typ, sep)] = str(d)
    return out


@pytest.fixture(scope="session")
def regular_index_dir(nothr_index_dirs):
    return nothr_index_dirs[("regular", False)]


def expected_text(read_set, query, sep=False):
    name = "%s%s.%s.txt.xz" % (read_set, ".sep" if sep else "", query)
    with lzma.open(os.path.join(GOLD, "expected", name)) as f:
        return f.read().decode()


def reads_path(read_set):
    return os.path.join(GOLD, "reads", read_set)


@pytest.fixture(scope="session")
def color_index_dir(tmp_path_factory):
    """Movi Color index directory (index.movi + doc_sets_flat.bin + ssa.movi + document info), built, coloured and sampled by the
    unmodified reference (tests/golden/make_golden_color.py)."""
    d = tmp_path_factory.mktemp("color")
    src = os.path.join(GOLD, "color")
    for f in ("index.movi", "doc_sets_flat.bin", "doc_sets.bin", "ssa.movi"):
        with lzma.open(os.path.join(src, f + ".xz")) as g:
            (d / f).write_bytes(g.read())
    for f in ("ref.fa.doc_offsets", "ref.fa.doc_ids"):
        (d / f).write_bytes(open(os.path.join(src, f), "rb").read())
    return str(d)


COLOR_VARIANTS = {"": dict(min_len=1, pvalue=False), ".minlen3": dict(min_len=3, pvalue=False), ".pvalue": dict(min_len=1, pvalue=True),
                  ".minlen0": dict(min_len=0, pvalue=False), ".pvalue.minlen0": dict(min_len=0, pvalue=True)}


def color_expected(tag):
    return open(os.path.join(GOLD, "color", "expected.multi_classify%s.csv" % tag)).read()


def parse_bpf_records(data, itemsize, header=True):
    """[(name bytes, values)] of a .bpf stream (reference src/utils.cpp:202-246); the .sa_entries.bpf file has no header."""
    import numpy as np
    import struct
    p = 12 if header else 0
    out = []
    while p < len(data):
        (nl,) = struct.unpack_from("<H", data, p); p += 2
        name = data[p:p + nl]; p += nl
        (L,) = struct.unpack_from("<Q", data, p); p += 8
        out.append((name, np.frombuffer(data[p:p + itemsize * L], dtype="<u%d" % itemsize))); p += itemsize * L
    return out
