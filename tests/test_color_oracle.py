"""Movi Color: the oracle's restatement of `--multi-classify` / `--sa-entries` against the outputs of the
unmodified reference (tests/golden/color, made by tests/golden/make_golden_color.py).  CPU only."""
import lzma
import os

import numpy as np
import pytest

from conftest import COLOR_VARIANTS, GOLD, color_expected, parse_bpf_records
from movi_b200 import Reads
from oracle.port import OracleIndex


@pytest.fixture(scope="module")
def color_case(color_index_dir):
    rd = Reads(os.path.join(GOLD, "color", "reads.fa"))
    return OracleIndex(color_index_dir), rd


@pytest.mark.parametrize("tag", sorted(COLOR_VARIANTS))
def test_oracle_multi_classify_equals_reference(color_case, tag):
    orc, rd = color_case
    assert orc.load_colors() == 5                       # six documents, two of them share a taxon id
    taxa, raw, score, ml = orc.query_multi_classify(rd.bases, rd.offsets, **COLOR_VARIANTS[tag])
    got = "".join("%s,%d,%d\n" % (name, t[0], t[1]) for name, t in zip(rd.ids(), taxa))
    assert got == color_expected(tag)
    assert np.array_equal(ml, orc.query_pml(rd.bases, rd.offsets))


def test_oracle_sa_entries_equal_reference(color_case):
    orc, rd = color_case
    ml, sa = orc.query_pml_sa(rd.bases, rd.offsets)
    want = parse_bpf_records(lzma.open(os.path.join(GOLD, "color", "expected.sa_entries.bpf.xz")).read(), 8, header=False)
    assert len(want) == rd.count
    for i, (name, v) in enumerate(want):
        s, e = int(rd.offsets[i]), int(rd.offsets[i + 1])
        assert name.decode() == rd.id(i) and np.array_equal(v, sa[s:e]), i
    wml = parse_bpf_records(lzma.open(os.path.join(GOLD, "color", "expected.pml.bpf.xz")).read(), 2)
    assert all(np.array_equal(v, ml[int(rd.offsets[i]):int(rd.offsets[i + 1])]) for i, (_, v) in enumerate(wml))


def test_oracle_build_sa_equals_reference_file(color_case, color_index_dir):
    """find_sampled_SA_entries restated: the oracle's `build-SA` on the colour index == the ssa.movi the reference wrote."""
    import struct
    orc, _ = color_case
    rate, sampled, all_p = OracleIndex(color_index_dir).build_ssa(100)
    want = open(os.path.join(color_index_dir, "ssa.movi"), "rb").read()
    got = struct.pack("<QQ", rate, len(sampled)) + sampled.tobytes() + struct.pack("<Q", len(all_p)) + all_p.tobytes()
    assert got == want
