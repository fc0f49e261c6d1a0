#!/usr/bin/env python3
"""Random-gather probe on cudaMalloc'd vs VMM-mapped (large alignment) tables."""
import json, os, sys, lzma, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movi_b200 import MoveStructure
td = tempfile.mkdtemp()
open(os.path.join(td, "index.movi"), "wb").write(lzma.open(os.path.join(ROOT, "tests/golden/c1/regular-thresholds/index.movi.xz")).read())
ms = MoveStructure(td)
for gb in (1, 8, 32):
    for align in (0, 2, 32, 512, 1024):
        try:
            ms.set_option("probe_vmm_align_mb", align)
            sec, loads = ms.gather_probe(gb << 30, 8, True, 512, 8, 256)
            rec = {"kind": "probe_vmm", "table_gb": gb, "align_mb": align, "loads_per_s": loads / sec}
        except Exception as e:
            rec = {"kind": "probe_vmm", "table_gb": gb, "align_mb": align, "error": str(e)}
        print(json.dumps(rec), flush=True)
        open(os.path.join(ROOT, "gpurun_out", "explore_vmm.jsonl"), "a").write(json.dumps(rec) + "\n")
