"""Parity at the scale of BASELINE.json configs[3]: a table of more than 2^32 rows (35 GB), so that the 64-bit row-index
kernels and the high id bits of the device row are exercised by REAL row indices -- no `force_wide`.  The table is synthesised in
HBM (datagen/synth_gpu.cu), written out as index.movi into RAM-backed storage, mapped by the CPU oracle, and PML / ZML / count
of >= 2000 reads must be bit-equal on both sides."""
import os
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROWS = int(float(os.environ.get("MOVI_SCALE_TEST_ROWS", "4.4e9")))       # 2^32 = 4.295e9


def _storage(need):
    for base in ("/dev/shm", "/tmp"):
        try:
            if shutil.disk_usage(base).free > need:
                return base
        except OSError:
            pass
    return None


def test_table_beyond_2_32_rows_matches_oracle():
    import torch
    from movi_b200.datagen.synth_gpu import SyntheticStructure
    from oracle.port import OracleIndex
    free, total = torch.cuda.mem_get_info(0)
    if free < ROWS * 17 + (8 << 30):          # rows + the generator's position array + slack
        pytest.skip("needs %.0f GB of free HBM" % ((ROWS * 17 + (8 << 30)) / 1e9))
    base = _storage(ROWS * 8 + (2 << 30))
    if base is None:
        pytest.skip("needs %.0f GB of /dev/shm or /tmp for the exported index" % (ROWS * 8 / 1e9))
    assert ROWS > (1 << 32)
    d = os.path.join(base, "movi_scale_test_%d" % os.getpid())
    ms = SyntheticStructure(ROWS, seed=45, profile=1)
    try:
        assert ms.get_option("force_wide") == 0
        dev = torch.device("cuda:0")
        ms.write_index_file(d)
        assert os.path.getsize(os.path.join(d, "index.movi")) > ROWS * 8
        orc = OracleIndex(d)
        assert (orc.r, orc.n) == (ms.r, ms.n)
        # rows beyond 2^32 hold ids beyond 2^32 and decode to the same fields on both sides
        top = orc.row(ROWS - 2)
        assert top["id"] < ROWS and orc.row(ROWS // 2)["n"] >= 1
        assert max(orc.row(k)["id"] for k in range(ROWS - 64, ROWS)) > (1 << 32)

        def both(n_reads, L, resync):
            d_b = torch.zeros(n_reads * L + 64, dtype=torch.uint8, device=dev)
            if resync:
                ms.make_reads_resync(d_b.data_ptr(), n_reads, L, 51, 10, 24, read_seed=7)
            else:
                ms.make_reads(d_b.data_ptr(), n_reads, L, 5, 50)
            return d_b[: n_reads * L].cpu().numpy(), np.arange(n_reads + 1, dtype=np.uint64) * L

        # PML: long reads written down by the query's own walk (they keep matching after an error), default kernel selection
        bases, offsets = both(2048, 3000, True)
        want = orc.query_pml(bases, offsets, threads=0)
        assert want.max() > 50 and np.median(want) >= 4
        for sector in (-1, 0, 1):
            ms.set_option("pml_sector", sector)
            assert np.array_equal(ms.query_pml(bases, offsets), want), "pml_sector=%d" % sector
        ms.set_option("pml_sector", -1)
        # PML / ZML / count on LF-walk reads
        bases, offsets = both(2500, 400, False)
        assert np.array_equal(ms.query_pml(bases, offsets), orc.query_pml(bases, offsets, threads=0))
        for sector in (-1, 1):
            ms.set_option("pml_sector", sector)
            assert np.array_equal(ms.query_zml(bases, offsets), orc.query_zml(bases, offsets, threads=0))
            c, p = ms.query_backward_search(bases, offsets)
            c2, p2 = orc.query_count(bases, offsets, threads=0)
            assert np.array_equal(c, c2) and np.array_equal(p, p2)
        assert c2.max() >= 1 and (p2 < 400).any()
        # the fused classify entry point on the long reads
        bases, offsets = both(512, 3000, True)
        a, b, s, ml = ms.query_classify(bases, offsets, 150, 7, want_ml=True)
        assert np.array_equal(ml, orc.query_pml(bases, offsets, threads=0))
        del orc
    finally:
        ms.close()
        shutil.rmtree(d, ignore_errors=True)
