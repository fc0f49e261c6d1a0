#!/usr/bin/env python3
"""Builder-run timing of the Movi Color entry points (host buffers in, results out) on the reference-built colour index of
tests/golden/color: multi-class classification (counts / p-value scores) and SA entries next to plain PML on the same reads.
Writes gpurun_out/r02_color_timing.jsonl.  The index is tiny (L2-resident): this measures the kernels' instruction cost, not HBM."""
import json
import lzma
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _color import random_reads                      # noqa: E402
from movi_b200 import MoveStructure                  # noqa: E402
from oracle.port import pack_reads                   # noqa: E402

gold = os.path.join(ROOT, "tests", "golden", "color")
out = open(os.path.join(ROOT, "gpurun_out", "r02_color_timing.jsonl"), "w")
with tempfile.TemporaryDirectory() as d:
    for f in ("index.movi", "doc_sets_flat.bin", "ssa.movi"):
        open(os.path.join(d, f), "wb").write(lzma.open(os.path.join(gold, f + ".xz")).read())
    for f in ("ref.fa.doc_offsets", "ref.fa.doc_ids"):
        open(os.path.join(d, f), "wb").write(open(os.path.join(gold, f), "rb").read())
    ms = MoveStructure(d, device=0)
    ms.load_colors(); ms.load_sampled_sa()
    reads = random_reads(d, 20000, seed=3, max_len=2000) * 25          # 500 k reads, ~1 kbp on average
    b, o = pack_reads(reads)
    n_bases = int(o[-1])

    def timed(name, f, reps=3):
        f()
        t = []
        for _ in range(reps):
            t0 = time.perf_counter(); f(); t.append(time.perf_counter() - t0)
        rec = {"what": name, "reads": len(reads), "bases": n_bases, "seconds": min(t), "bases_per_s": n_bases / min(t)}
        out.write(json.dumps(rec) + "\n"); out.flush()
        print(rec)

    timed("movi_gpu_query_pml (host buffers)", lambda: ms.query_pml(b, o))
    timed("movi_gpu_query_multi_classify, counts, no lengths back", lambda: ms.query_multi_classify(b, o))
    timed("movi_gpu_query_multi_classify, counts, lengths back", lambda: ms.query_multi_classify(b, o, want_ml=True))
    timed("movi_gpu_query_multi_classify, p-value scores", lambda: ms.query_multi_classify(b, o, min_len=0, pvalue=True))
    timed("movi_gpu_query_pml_sa (sample rate 100)", lambda: ms.query_pml_sa(b, o))
    ms.close()
