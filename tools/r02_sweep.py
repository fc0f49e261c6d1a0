#!/usr/bin/env python3
"""Round-2 kernel sweeps (device-resident inputs, CUDA-event timing).  Writes gpurun_out/r02_sweep.jsonl.

usage: r02_sweep.py c2 [pml|zml|count ...]         c2 index, 1 M x 150 bp reads: pml_sector 0/1 x blocks_per_sm
       r02_sweep.py long <n_reads>                 c2 index, n x 10 kbp reads
       r02_sweep.py synth <GB> <n_reads> <len>     synthetic HBM-resident table (datagen/synth_gpu.cu)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def emit(rec):
    print(json.dumps(rec), flush=True)
    open(os.path.join(OUT, "r02_sweep.jsonl"), "a").write(json.dumps(rec) + "\n")


def timed(fn, reps=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def sweep(ms, tag, d_b, d_o, n_reads, n_bases, kinds, settings, extra=None):
    import torch
    dev = d_b.device
    d_out = torch.zeros(n_bases + 64, dtype=torch.int16, device=dev)
    d_cnt = torch.zeros(n_reads, dtype=torch.int64, device=dev)
    d_pos = torch.zeros(n_reads, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    first = {}
    for kind in kinds:
        for setting in settings:
            sector, bps = setting[0], setting[1]
            out_stage = setting[2] if len(setting) > 2 else 32
            ms.set_option("pml_sector", sector)
            ms.set_option("blocks_per_sm", bps)
            ms.set_option("out_stage", out_stage)
            if kind == "pml":
                fn = lambda: ms.query_pml_device(d_b.data_ptr(), d_o.data_ptr(), n_reads, d_out.data_ptr(), st)
            elif kind == "zml":
                fn = lambda: ms.query_zml_device(d_b.data_ptr(), d_o.data_ptr(), n_reads, d_out.data_ptr(), st)
            else:
                fn = lambda: ms.query_count_device(d_b.data_ptr(), d_o.data_ptr(), n_reads, d_cnt.data_ptr(), d_pos.data_ptr(), st)
            t = timed(fn)
            ms.check_device_status()
            res = d_cnt.clone() if kind == "count" else d_out[:n_bases].clone()
            same = True
            if kind in first:
                same = bool(torch.equal(first[kind], res))
            else:
                first[kind] = res
            rec = {"kind": kind, "tag": tag, "pml_sector": sector, "blocks_per_sm": bps, "out_stage": out_stage, "n_reads": n_reads, "n_bases": n_bases,
                   "ms": round(t, 3), "Gbases_per_s": round(n_bases / t / 1e6, 2), "same_as_first": same}
            if extra:
                rec.update(extra)
            emit(rec)
    ms.set_option("pml_sector", -1)
    ms.set_option("blocks_per_sm", 0)


def main():
    import torch
    dev = torch.device("cuda:0")
    what = sys.argv[1]
    settings = [(0, 0), (1, 0), (1, 6), (1, 4), (0, 5)]
    if what in ("c2", "g1", "long"):
        from movi_b200 import MoveStructure
        from movi_b200.datagen import synth
        name = "g1" if what == "g1" else "c2"
        d, haps = synth.prepare(name, ("regular-thresholds",))
        ms = MoveStructure(os.path.join(d, "regular-thresholds"))
        if what in ("c2", "g1"):
            bases, offsets = synth.make_reads(haps, 1_000_000, 150, seed=43)
            kinds = sys.argv[2:] or ["pml"]
            tag = name + " 1Mx150"
        else:
            n = int(sys.argv[2])
            bases, offsets = synth.make_reads(haps, n, 10_000, seed=45, sub=0.05, frac_random=0.0)
            kinds = ["pml"]
            tag = "c2 %dx10k" % n
        d_b = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).to(dev)
        d_o = torch.from_numpy(offsets.astype(np.int64)).to(dev)
        sweep(ms, tag, d_b, d_o, len(offsets) - 1, len(bases), kinds, settings)
        ms.close()
    elif what == "synth":
        from movi_b200.datagen.synth_gpu import SyntheticStructure
        gb, n_reads, L = float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        profile = int(sys.argv[5]) if len(sys.argv) > 5 else 1
        r = int(gb * (1 << 30) / 8)
        ms = SyntheticStructure(r, seed=44, profile=profile)
        n_bases = n_reads * L
        d_b = torch.zeros(n_bases + 64, dtype=torch.uint8, device=dev)
        resync = os.environ.get("MOVI_SWEEP_RESYNC", "1") == "1"
        if resync:
            ms.make_reads_resync(d_b.data_ptr(), n_reads, L, 51, 10, 24, read_seed=3)      # the c4 read mix of bench.py
        else:
            ms.make_reads(d_b.data_ptr(), n_reads, L, 5, 50)
        d_o = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * L
        settings = [(1, 0, 32), (1, 0, 16), (1, 4, 32), (1, 4, 16), (1, 3, 32), (0, 0, 16)]
        if os.environ.get("MOVI_SWEEP_ONE"):
            settings = [(1, 0, 32)]
        sweep(ms, "synth %.0fGB p%d %dx%d %s" % (gb, profile, n_reads, L, "resync" if resync else "lfwalk"), d_b, d_o, n_reads, n_bases,
              sys.argv[6:] or ["pml"], settings, {"n_over_r": round(ms.n / r, 1)})
        ms.close()


if __name__ == "__main__":
    main()
