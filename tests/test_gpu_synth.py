"""The human-scale table generator (datagen/synth_gpu.cu) at a size the CPU oracle can load: the table it builds in HBM,
written out as index.movi, gives the same PML / ZML / count on the CPU oracle as on the GPU -- i.e. it is a valid move
structure and the adopt-device-rows entry point works (also with 64-bit row indices forced)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("profile", [0, 1, 2])
def test_synthetic_table_matches_oracle(tmp_path, profile):
    import torch
    from movi_b200.datagen.synth_gpu import SyntheticStructure
    from oracle.port import OracleIndex
    ms = SyntheticStructure(1 << 20, seed=44, profile=profile)
    n_reads, L = 3000, 400
    dev = torch.device("cuda:0")
    d_bases = torch.zeros(n_reads * L + 64, dtype=torch.uint8, device=dev)
    ms.make_reads(d_bases.data_ptr(), n_reads, L)
    bases = d_bases[: n_reads * L].cpu().numpy()
    assert set(np.unique(bases)) <= set(b"ACGT")
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * L
    orc = OracleIndex(ms.write_index_file(str(tmp_path / "synth")))
    assert (orc.r, orc.n) == (ms.r, ms.n)
    want = orc.query_pml(bases, offsets, threads=0)
    assert want.max() > 30                      # LF-walk reads really match
    for wide in (0, 1):
        ms.set_option("force_wide", wide)
        assert np.array_equal(ms.query_pml(bases, offsets), want)
        assert np.array_equal(ms.query_zml(bases, offsets), orc.query_zml(bases, offsets, threads=0))
        c, p = ms.query_backward_search(bases, offsets)
        c2, p2 = orc.query_count(bases, offsets, threads=0)
        assert np.array_equal(c, c2) and np.array_equal(p, p2)
    ms.close()
