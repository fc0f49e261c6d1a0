#!/usr/bin/env python3
"""Timeline of the host-buffer pipeline on the c2 workload (MOVI_GPU_TRACE=1): per-slice copy / kernel / copy-back times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MOVI_GPU_TRACE"] = "1"
import numpy as np, torch
import bench
from movi_b200 import MoveStructure
index_dir, bases, offsets, cfg = bench.prepare_workload(0)
ms = MoveStructure(index_dir, device=0)
n_reads = len(offsets) - 1
h_bases = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).pin_memory()
h_off = torch.from_numpy(offsets.astype(np.int64)).pin_memory()
h_out = torch.empty(len(bases) + 16, dtype=torch.int16).pin_memory()
for bb in [int(x) for x in (sys.argv[1:] or ["16"])]:
    ms.set_option("batch_bases", bb << 20)
    for rep in range(3):
        sys.stderr.write("--- batch_bases %d M, rep %d\n" % (bb, rep))
        t0 = time.perf_counter()
        ms._check(ms._L.movi_gpu_query_pml(ms._ctx, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_out.data_ptr()))
        sys.stderr.write("wall %.3f ms\n" % ((time.perf_counter() - t0) * 1e3))
