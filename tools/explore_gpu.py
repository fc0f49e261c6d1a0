#!/usr/bin/env python3
"""Exploratory GPU measurements (not the bench): gather-probe sweep and PML kernel variants on the
cache-resident C1 index.  Writes JSON lines to gpurun_out/explore.jsonl."""
import json
import lzma
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def emit(rec):
    print(json.dumps(rec), flush=True)
    with open(os.path.join(OUT, "explore.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")


def main():
    import torch
    from movi_b200 import MoveStructure
    gold = os.path.join(ROOT, "tests", "golden")
    td = tempfile.mkdtemp()
    with lzma.open(os.path.join(gold, "c1", "regular-thresholds", "index.movi.xz")) as f:
        open(os.path.join(td, "index.movi"), "wb").write(f.read())
    ms = MoveStructure(td)
    what = sys.argv[1:] or ["probe", "pml"]

    if "probe" in what:
        for gb in (0.0625, 1, 8, 32, 96):
            tb = int(gb * (1 << 30))
            for load_bytes in (8, 32):
                for dep in (1, 0):
                    for bps, tpb in ((8, 256), (4, 256), (2, 1024)):
                        if dep == 0 and (bps, tpb) != (8, 256):
                            continue
                        try:
                            sec, loads = ms.gather_probe(tb, load_bytes, bool(dep), 512, bps, tpb)
                        except Exception as e:
                            emit({"probe_error": str(e), "gb": gb})
                            continue
                        emit({"kind": "probe", "table_gb": gb, "load_bytes": load_bytes, "dependent": dep, "blocks_per_sm": bps,
                              "threads": tpb, "loads_per_s": loads / sec, "sector_GBps": loads / sec * 32 / 1e9, "ms": sec * 1e3})

    if "pml" in what:
        txt = lzma.open(os.path.join(gold, "ref.fasta.xz")).read().decode()
        ref = np.frombuffer("".join(l.strip() for l in txt.split("\n") if not l.startswith(">")).upper().encode(), np.uint8)
        rng = np.random.default_rng(1)
        n_reads, L = 1_000_000, 150
        starts = rng.integers(0, len(ref) - L, n_reads)
        bases = ref[(starts[:, None] + np.arange(L)[None, :])].reshape(-1).copy()
        err = rng.random(bases.shape[0]) < 0.005
        bases[err] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(err.sum()))]
        offsets = (np.arange(n_reads + 1, dtype=np.uint64) * L)
        dev = torch.device("cuda:0")
        d_bases = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).to(dev)
        d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev)
        d_out = torch.zeros(len(bases) + 64, dtype=torch.int16, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        ref_out = None
        for variant in (0, 1):
            for bps in (0, 2, 4, 6, 8):
                ms.set_option("pml_variant", variant)
                ms.set_option("blocks_per_sm", bps)
                for _ in range(2):
                    ms.query_pml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_out.data_ptr(), stream)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    ms.query_pml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_out.data_ptr(), stream)
                e1.record()
                torch.cuda.synchronize()
                ms.check_device_status()
                ms_per = e0.elapsed_time(e1) / 5
                o = d_out.cpu().numpy()
                if ref_out is None:
                    ref_out = o
                emit({"kind": "pml_c1", "variant": variant, "blocks_per_sm": bps, "ms": ms_per, "Gbases_per_s": n_reads * L / ms_per / 1e6,
                      "same_as_first": bool(np.array_equal(o, ref_out))})
        # e2e through the host entry point (pageable numpy buffers)
        ms.set_option("blocks_per_sm", 0)
        out = np.empty(len(bases), np.uint16)
        for _ in range(2):
            t0 = time.time(); ms.query_pml(bases, offsets, out); t1 = time.time()
        emit({"kind": "pml_c1_host_pageable", "ms": (t1 - t0) * 1e3, "Gbases_per_s": n_reads * L / (t1 - t0) / 1e9})


if __name__ == "__main__":
    main()
