"""Movi Color on the GPU through the C ABI: movi_gpu_query_multi_classify / movi_gpu_query_pml_sa against the files the
unmodified reference wrote (tests/golden/color) and against the oracle on ragged reads."""
import lzma
import os

import numpy as np
import pytest

from _color import random_reads, result_scores
from conftest import COLOR_VARIANTS, GOLD, color_expected, parse_bpf_records
from movi_b200 import MoveStructure, MoviGpuError, Reads
from oracle.port import OracleIndex, pack_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ms(color_index_dir):
    m = MoveStructure(color_index_dir, device=0)
    assert m.load_colors() == 5
    assert m.load_sampled_sa() == 100
    yield m
    m.close()


@pytest.fixture(scope="module")
def reads():
    return Reads(os.path.join(GOLD, "color", "reads.fa"))


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("tag", sorted(COLOR_VARIANTS))
def test_multi_classify_equals_the_reference_file(ms, reads, tag, wide):
    v = COLOR_VARIANTS[tag]
    ms.set_option("force_wide", int(wide))
    try:
        before = ms.launch_count
        res, ml = ms.query_multi_classify(reads.bases, reads.offsets, want_ml=True, **v)
        assert ms.launch_count == before + 1
    finally:
        ms.set_option("force_wide", 0)
    assert ms.multi_classify_text(reads, res, **v) == color_expected(tag)
    want = parse_bpf_records(lzma.open(os.path.join(GOLD, "color", "expected.pml.bpf.xz")).read(), 2)
    for i, (_, w) in enumerate(want):
        assert np.array_equal(w, ml[int(reads.offsets[i]):int(reads.offsets[i + 1])]), i


@pytest.mark.parametrize("wide", [False, True])
def test_sa_entries_equal_the_reference_file(ms, reads, wide):
    ms.set_option("force_wide", int(wide))
    try:
        ml, sa = ms.query_pml_sa(reads.bases, reads.offsets)
    finally:
        ms.set_option("force_wide", 0)
    want = parse_bpf_records(lzma.open(os.path.join(GOLD, "color", "expected.sa_entries.bpf.xz")).read(), 8, header=False)
    assert len(want) == reads.count
    for i, (_, w) in enumerate(want):
        assert np.array_equal(w, sa[int(reads.offsets[i]):int(reads.offsets[i + 1])]), i
    assert np.array_equal(ml, ms.query_pml(reads.bases, reads.offsets))


def test_ragged_reads_equal_the_oracle(ms, color_index_dir):
    """20 000 ragged reads (empty, 1-base, with N, random): more reads than lanes of a small grid, every counter slot reused."""
    orc = OracleIndex(color_index_dir)
    b, o = pack_reads(random_reads(color_index_dir, 20000, seed=9))
    for v in (dict(min_len=1, pvalue=False), dict(min_len=2, pvalue=True), dict(min_len=0, pvalue=True)):
        res, ml = ms.query_multi_classify(b, o, want_ml=True, **v)
        taxa, raw, score, oml = orc.query_multi_classify(b, o, **v)
        assert np.array_equal(np.stack([res["best_doc"], res["second_best_doc"], res["sum_matching_lengths"]], axis=1), raw), v
        assert np.array_equal(result_scores(res, v["pvalue"]), score), v
        assert np.array_equal(ml, oml), v
    ml, sa = ms.query_pml_sa(b, o)
    oml, osa = orc.query_pml_sa(b, o)
    assert np.array_equal(ml, oml) and np.array_equal(sa, osa)


def test_missing_tables_and_unsupported_layouts_fail_loudly(color_index_dir, index_dirs, reads):
    m = MoveStructure(color_index_dir, device=0)
    try:
        with pytest.raises(MoviGpuError, match="no colour table"):
            m.query_multi_classify(reads.bases, reads.offsets)
        with pytest.raises(MoviGpuError, match="no sampled suffix array"):
            m.query_pml_sa(reads.bases, reads.offsets)
    finally:
        m.close()
    with pytest.raises(RuntimeError, match="doc_sets_flat.bin"):
        m2 = MoveStructure(index_dirs[("regular-thresholds", False)], device=0)
        try:
            m2.load_colors()
        finally:
            m2.close()
