"""The N > 1 path (read sharding + in-order gather) with world_size 2 over gloo on the CPU.
The per-shard compute is stood in for by the CPU oracle: the plumbing is what is under test."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT
from movi_b200.shard import local_shard, shard_bounds


def test_shard_bounds_partition_and_balance():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 3000, 1000)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 3, 4, 8):
        b = shard_bounds(offsets, world)
        assert b[0][0] == 0 and b[-1][1] == 1000
        assert all(b[k][1] == b[k + 1][0] for k in range(world - 1))
        sizes = [int(offsets[i1] - offsets[i0]) for i0, i1 in b]
        assert max(sizes) - min(sizes) <= 2 * 3000
    # degenerate inputs
    assert shard_bounds(np.zeros(1, np.uint64), 4) == [(0, 0)] * 4
    assert shard_bounds(np.array([0, 5], np.uint64), 2)[-1][1] == 1


def test_local_shard_rebases_offsets():
    offsets = np.array([0, 3, 7, 7, 12], dtype=np.uint64)
    bases = np.arange(12, dtype=np.uint8)
    got = [local_shard(bases, offsets, r, 2) for r in range(2)]
    assert np.concatenate([g[0] for g in got]).tolist() == bases.tolist()
    for b, o, (i0, i1) in got:
        assert o[0] == 0 and int(o[-1]) == len(b) and len(o) == i1 - i0 + 1


WORKER = textwrap.dedent("""
    import os, sys, lzma, tempfile
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %(root)r)
    from movi_b200 import Reads
    from movi_b200.shard import local_shard, gather_in_read_order
    from oracle.port import OracleIndex
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    gold = os.path.join(%(root)r, "tests", "golden")
    td = tempfile.mkdtemp()
    open(os.path.join(td, "index.movi"), "wb").write(lzma.open(os.path.join(gold, "c1", "regular-thresholds", "index.movi.xz")).read())
    rd = Reads(os.path.join(gold, "reads", "synth300.fasta"))
    orc = OracleIndex(td)
    b, o, (i0, i1) = local_shard(rd.bases, rd.offsets, rank, world)
    pml = gather_in_read_order(orc.query_pml(b, o), dist)
    cnt = gather_in_read_order(orc.query_count(b, o)[0], dist)
    if rank == 0:
        assert np.array_equal(pml, orc.query_pml(rd.bases, rd.offsets))
        assert np.array_equal(cnt, orc.query_count(rd.bases, rd.offsets)[0])
        print("SHARD_OK", len(pml), len(cnt))
    dist.destroy_process_group()
""")


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300,
                       env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0 and "SHARD_OK" in r.stdout, r.stdout[-2000:]
