#!/usr/bin/env python3
"""Sweep PML kernel variants / launch shapes / L2 options on a cached synthetic config (default c2).
usage: explore_pml.py [config] [spec ...]   spec = variant:blocks_per_sm:l2_persist_mb:fetch_granularity[:threads[:row_ld_mode]]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def main():
    import torch
    from movi_b200 import MoveStructure
    from movi_b200.datagen import synth
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    specs = sys.argv[2:] or ["1:0:0:0", "2:0:0:0", "3:0:0:0"]
    cfg = synth.CONFIGS[name]
    d, haps = synth.prepare(name, ("regular-thresholds",))
    bases, offsets = synth.make_reads(haps, cfg["n_reads"], cfg["read_len"], seed=43)
    n_reads, n_bases = cfg["n_reads"], len(bases)
    itype = os.environ.get("MOVI_EXPLORE_TYPE", "regular-thresholds")
    if itype != "regular-thresholds":
        d, _ = synth.prepare(name, (itype,), want_reads=False)
    ms = MoveStructure(os.path.join(d, itype), layout=int(os.environ.get("MOVI_EXPLORE_LAYOUT", "0")))
    print("index", itype, "device layout", ms.device_layout, "bytes", ms.index_device_bytes, flush=True)
    dev = torch.device("cuda:0")
    d_bases = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).to(dev)
    d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    d_out = torch.zeros(n_bases + 64, dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    first = None
    print("l2 persist max MB", ms.get_option("l2_persist_max_mb"), "fetch granularity", ms.get_option("l2_fetch_granularity"), flush=True)
    for spec in specs:
        f = [int(x) for x in spec.split(":")]
        variant, bps, persist, gran = f[:4]
        threads = f[4] if len(f) > 4 else 256
        ld_mode = f[5] if len(f) > 5 else 0
        n_streams = f[6] if len(f) > 6 else 1
        ms.set_option("smem_carveout", f[7] if len(f) > 7 else -1)
        ms.set_option("row_keep_pct", f[8] if len(f) > 8 else 100)
        ms.set_option("row_ld_mode", ld_mode)
        ms.set_option("pml_variant", variant)
        ms.set_option("blocks_per_sm", bps)
        ms.set_option("threads_per_block", threads)
        ms.set_option("l2_persist_mb", persist)
        if gran:
            ms.set_option("l2_fetch_granularity", gran)
        d_out.zero_()
        for _ in range(3):
            ms.query_pml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_out.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 12
        if n_streams == 1:
            e0.record()
            for _ in range(reps):
                ms.query_pml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_out.data_ptr(), stream)
            e1.record()
        else:
            side = [torch.cuda.Stream() for _ in range(n_streams)]
            e0.record()
            for s_ in side:
                s_.wait_stream(torch.cuda.current_stream())
            for i in range(reps):
                ms.query_pml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_out.data_ptr(), side[i % n_streams].cuda_stream)
            for s_ in side:
                torch.cuda.current_stream().wait_stream(s_)
            e1.record()
        torch.cuda.synchronize()
        ms.check_device_status()
        t = e0.elapsed_time(e1) / reps
        o = d_out[:n_bases].cpu().numpy()
        if first is None:
            first = o
        rec = {"kind": "pml_" + name, "spec": spec, "ms": round(t, 4), "Gbases_per_s": round(n_bases / t / 1e6, 2),
               "same_as_first": bool(np.array_equal(o, first)), "gran_now": ms.get_option("l2_fetch_granularity")}
        print(json.dumps(rec), flush=True)
        with open(os.path.join(OUT, "explore_pml.jsonl"), "a") as fh:
            fh.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
