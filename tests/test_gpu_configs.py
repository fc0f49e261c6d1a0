"""Parity at the shapes of BASELINE.json's configs[1..4], scaled so that the CPU oracle finishes in
seconds, plus size-independent properties on the full-size c2 index when its cache is present."""
import os

import numpy as np
import pytest

from conftest import ROOT, TYPES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small():
    """c2-small: 16 haplotypes x 100 kbp pangenome (+reverse complements), all three layouts, built on the fly."""
    from movi_b200.datagen import synth
    d, haps = synth.prepare("c2-small", tuple(TYPES))
    return d, haps


@pytest.fixture(scope="module")
def small_structs(small):
    from movi_b200 import MoveStructure
    d, _ = small
    ms = {t: MoveStructure(os.path.join(d, t)) for t in TYPES}
    yield ms
    for m in ms.values():
        m.close()


def test_c2_c3_shape_short_reads_all_layouts(small, small_structs):
    """configs[1] / configs[2]: 150 bp reads; PML on every layout, count + ZML on regular / blocked; all == oracle."""
    from movi_b200.datagen import synth
    from oracle.port import OracleIndex
    d, haps = small
    bases, offsets = synth.make_reads(haps, 20000, 150, seed=43)
    ref_pml = None
    for t in TYPES:
        orc = OracleIndex(os.path.join(d, t))
        want = orc.query_pml(bases, offsets, threads=0)
        assert np.array_equal(small_structs[t].query_pml(bases, offsets), want)
        ref_pml = want if ref_pml is None else ref_pml
        assert np.array_equal(want, ref_pml)            # layouts agree (reference tests/test_pml.cpp:57-115)
        if t != "sampled-thresholds":
            structs = [small_structs[t]]
            if t == "blocked-thresholds":       # configs[2] names the blocked layout: also with its 6-byte rows kept in HBM
                from movi_b200 import MoveStructure
                from movi_b200.api import LAYOUT_COMPACT
                structs.append(MoveStructure(os.path.join(d, t), layout=LAYOUT_COMPACT))
                assert np.array_equal(structs[-1].query_pml(bases, offsets), want)
            for ms_ in structs:
                assert np.array_equal(ms_.query_zml(bases, offsets), orc.query_zml(bases, offsets, threads=0))
                c, p = ms_.query_backward_search(bases, offsets)
                c2, p2 = orc.query_count(bases, offsets, threads=0)
                assert np.array_equal(c, c2) and np.array_equal(p, p2)


def _long_reads(haps, n, seed):
    from movi_b200.datagen import synth
    return synth.make_reads(haps, n, 10000, seed=seed, sub=0.05, frac_random=0.10, len_jitter=3000)


def test_c4_shape_long_reads_classify_and_wide_indices(small, small_structs):
    """configs[3]: ~10 kbp nanopore-like reads (5 % errors, ragged lengths), PML + classification, also through the
    64-bit-row-index kernel instantiations that tables beyond 2^32 rows use."""
    import torch
    from oracle.port import OracleIndex, classify
    d, haps = small
    bases, offsets = _long_reads(haps, 300, 44)
    ms = small_structs["regular-thresholds"]
    orc = OracleIndex(os.path.join(d, "regular-thresholds"))
    want = orc.query_pml(bases, offsets, threads=0)
    for wide in (0, 1):
        ms.set_option("force_wide", wide)
        try:
            want_z = orc.query_zml(bases, offsets, threads=0)
            c2, p2 = orc.query_count(bases, offsets, threads=0)
            for sector in (-1, 1):                   # 1: sector-cached row fetches (what tables beyond the L2 use)
                ms.set_option("pml_sector", sector)
                try:
                    assert np.array_equal(ms.query_pml(bases, offsets), want)
                    assert np.array_equal(ms.query_zml(bases, offsets), want_z)
                    c, p = ms.query_backward_search(bases, offsets)
                    assert np.array_equal(c, c2) and np.array_equal(p, p2)
                finally:
                    ms.set_option("pml_sector", -1)
        finally:
            ms.set_option("force_wide", 0)
    # device-side classification of the device-resident PMLs
    dev = torch.device("cuda:0")
    n = len(offsets) - 1
    d_ml = torch.from_numpy(want.view(np.int16)).to(dev)
    d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    a, b, s = (torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int64, device=dev))
    ms.classify_device(d_ml.data_ptr(), d_off.data_ptr(), n, 150, 7, a.data_ptr(), b.data_ptr(), s.data_ptr())
    torch.cuda.synchronize()
    a, b, s = a.cpu().numpy(), b.cpu().numpy(), s.cpu().numpy()
    found = 0
    for i in range(n):
        f, ea, eb, es, _ = classify(want[int(offsets[i]):int(offsets[i + 1])], 150, 7)
        assert (a[i], b[i], s[i]) == (ea, eb, es)
        found += f
    assert 0 < found < n        # matching reads FOUND, the 10 % random reads NOT_PRESENT


def test_c5_shape_sampled_index_long_reads_filter(small, small_structs, tmp_path):
    """configs[4]: sampled-thresholds (id checkpoints every 20 rows), long reads, PML + --filter through the CLI."""
    import subprocess
    from movi_b200 import build
    from oracle.port import OracleIndex, classify
    d, haps = small
    bases, offsets = _long_reads(haps, 120, 45)
    sd = os.path.join(d, "sampled-thresholds")
    want = OracleIndex(sd).query_pml(bases, offsets, threads=0)
    assert np.array_equal(small_structs["sampled-thresholds"].query_pml(bases, offsets), want)
    # the memory-constrained form: rows and id checkpoints stay packed in HBM, get_id runs per LF step
    from movi_b200 import MoveStructure
    from movi_b200.api import LAYOUT_COMPACT
    packed = MoveStructure(sd, layout=LAYOUT_COMPACT)
    assert packed.index_device_bytes < 0.65 * small_structs["sampled-thresholds"].index_device_bytes
    assert np.array_equal(packed.query_pml(bases, offsets), want)
    packed.close()
    # --filter needs a null database next to the index: freeze one with percentile 6 (thr = 7)
    import struct
    with open(os.path.join(sd, "movi.pml.nulldb"), "wb") as f:
        f.write(struct.pack("<QdQ", 1, 1.0, 6) + struct.pack("<Q", 1))
    fa = tmp_path / "long.fa"
    with open(fa, "w") as f:
        for i in range(len(offsets) - 1):
            f.write(">read%d\n%s\n" % (i, bases[int(offsets[i]):int(offsets[i + 1])].tobytes().decode()))
    build.build_host()
    r = subprocess.run([build.CLI, "query", "-i", sd, "-r", str(fa), "--pml", "--filter"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-500:]
    got = [l[1:] for l in r.stdout.decode().split("\n") if l.startswith(">")]
    exp = ["read%d" % i for i in range(len(offsets) - 1)
           if classify(want[int(offsets[i]):int(offsets[i + 1])], 150, 7)[0]]
    assert got == exp and 0 < len(exp) < len(offsets) - 1


def test_full_size_c2_properties():
    """On the full bench index (r = 1.55e7, 1 M reads): properties that hold for any correct PML output, and agreement
    of every layout that is present in the cache with the regular-thresholds result."""
    from movi_b200 import MoveStructure
    from movi_b200.datagen import synth
    d = synth.cache_dir("c2")
    for t in TYPES:
        synth._unpack(d, t)             # the cache ships xz-compressed
    if not os.path.exists(os.path.join(d, "regular-thresholds", "index.movi")):
        pytest.skip("full-size c2 cache not present (bench.py builds it on demand: ~4 min)")
    cfg = synth.CONFIGS["c2"]
    haps = synth.pangenome(cfg["genome_len"], cfg["n_hap"])
    bases, offsets = synth.make_reads(haps, 200000, 150, seed=43)
    ms = MoveStructure(os.path.join(d, "regular-thresholds"))
    pml = ms.query_pml(bases, offsets).astype(np.int64).reshape(-1, 150)
    zml = ms.query_zml(bases, offsets).astype(np.int64).reshape(-1, 150)
    cnt, pos = ms.query_backward_search(bases, offsets)
    ms.close()
    # a matching length grows by at most one per base and never exceeds the number of bases consumed
    assert (pml[:, 1:] <= pml[:, :-1] + 1).all() and (pml <= np.arange(1, 151)[None, :]).all()
    assert (zml[:, 1:] <= zml[:, :-1] + 1).all() and (zml <= np.arange(150)[None, :]).all()
    # count: reads matched end to end occur at least once; the matched suffix length is the first ZML restart
    full = pos == 0
    assert full.any() and (cnt[full] >= 1).all() and (cnt >= 1).all()
    first_restart = np.where((zml[:, 1:] < zml[:, :-1] + 1).any(1), (zml[:, 1:] < zml[:, :-1] + 1).argmax(1) + 1, 150)
    assert np.array_equal(150 - pos, first_restart)
    for t in TYPES[1:]:
        if os.path.exists(os.path.join(d, t, "index.movi")):
            m2 = MoveStructure(os.path.join(d, t))
            assert np.array_equal(m2.query_pml(bases, offsets).astype(np.int64).reshape(-1, 150), pml), t
            m2.close()
