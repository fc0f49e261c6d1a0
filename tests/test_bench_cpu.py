"""bench.py pieces that need no GPU: the FASTA writer the CPU arm and the command-line leg are fed from, the workload table."""
import os
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)


def test_write_fasta_round_trips_through_the_reader(tmp_path):
    import bench
    from movi_b200 import Reads
    rng = np.random.default_rng(5)
    n, L = 300, 77
    bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n * L)]
    offsets = np.arange(n + 1, dtype=np.uint64) * L
    f = str(tmp_path / "r.fa")
    bench.write_fasta(f, bases, offsets, n, repeat=3)
    rd = Reads(f)
    assert rd.count == 3 * n
    assert np.array_equal(rd.bases, np.tile(bases, 3)) and rd.id(0) == "r0000000" and rd.id(n) == "r0000000"
    ragged = np.concatenate([[0], np.cumsum(rng.integers(1, 50, 40))]).astype(np.uint64)
    b2 = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(ragged[-1]))]
    bench.write_fasta(f, b2, ragged, 40)
    rd = Reads(f)
    assert rd.count == 40 and np.array_equal(rd.offsets, ragged) and np.array_equal(rd.bases, b2)


def test_workload_table_names_the_baseline_configs():
    import bench
    assert set(bench.WORKLOADS) == {"c4", "c2", "c3-zml", "c3-count", "g1"}
    assert bench.WORKLOADS["c4"]["algo"] == 32 and bench.WORKLOADS["c3-zml"]["algo"] == 64
    assert bench.C4_ROWS > (1 << 32) and bench.C4_READS * bench.C4_READ_LEN < (1 << 32)      # real 64-bit row ids; one device call per step
    assert bench.C4_READS >= 400_000
