#!/usr/bin/env python3
"""Timeline of the host-buffer pipeline on the c2 workload (MOVI_GPU_TRACE=1): per-slice copy / kernel / copy-back times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--trace" in sys.argv:
    sys.argv.remove("--trace")
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
# argv: [--api pml|ml8|cls] batch_bases(M)[:kernel_streams[:pipeline_slots[:size_grid_to_batch]]] ...
api = "pml"
if "--api" in sys.argv:
    k = sys.argv.index("--api"); api = sys.argv[k + 1]; del sys.argv[k:k + 2]
h_out8 = torch.empty(len(bases) + 16, dtype=torch.uint8).pin_memory()
h_above, h_below = torch.zeros(n_reads, dtype=torch.int32).pin_memory(), torch.zeros(n_reads, dtype=torch.int32).pin_memory()
h_sum = torch.zeros(n_reads, dtype=torch.int64).pin_memory()
def call():
    L, c = ms._L, ms._ctx
    if api == "pml": ms._check(L.movi_gpu_query_pml(c, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_out.data_ptr()))
    elif api == "ml8": ms._check(L.movi_gpu_query_ml8(c, 0, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_out8.data_ptr()))
    else: ms._check(L.movi_gpu_query_classify(c, 0, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, 150, 7, h_above.data_ptr(), h_below.data_ptr(), h_sum.data_ptr(), None))
for spec in (sys.argv[1:] or ["8"]):
    f = [int(x) for x in spec.split(":")]
    ms.set_option("batch_bases", f[0] << 20)
    if len(f) > 1: ms.set_option("kernel_streams", f[1])
    if len(f) > 2: ms.set_option("pipeline_slots", f[2])
    if len(f) > 3: ms.set_option("size_grid_to_batch", f[3])
    walls = []
    for rep in range(6):
        sys.stderr.write("--- spec %s, rep %d\n" % (spec, rep))
        t0 = time.perf_counter()
        call()
        walls.append((time.perf_counter() - t0) * 1e3)
        sys.stderr.write("wall %.3f ms\n" % walls[-1])
    print('{"api": "%s", "spec": "%s", "wall_ms_min": %.3f, "wall_ms_median": %.3f}' % (api, spec, min(walls[1:]), sorted(walls[1:])[len(walls[1:]) // 2]), flush=True)
