#!/usr/bin/env python3
"""Time the ZML and count kernels on a cached synthetic config (device-resident inputs)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from movi_b200 import MoveStructure
    from movi_b200.datagen import synth
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    cfg = synth.CONFIGS[name]
    d, haps = synth.prepare(name, ("regular-thresholds",))
    bases, offsets = synth.make_reads(haps, cfg["n_reads"], cfg["read_len"], seed=43)
    n_reads, n_bases = cfg["n_reads"], len(bases)
    ms = MoveStructure(os.path.join(d, "regular-thresholds"))
    dev = torch.device("cuda:0")
    d_bases = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).to(dev)
    d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    d_out = torch.zeros(n_bases + 64, dtype=torch.int16, device=dev)
    d_cnt = torch.zeros(n_reads, dtype=torch.int64, device=dev)
    d_pos = torch.zeros(n_reads, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for spec in (sys.argv[2:] or ["1:0"]):
        variant, bps = [int(x) for x in spec.split(":")]
        ms.set_option("bs_variant", variant)
        ms.set_option("blocks_per_sm", bps)
        for kind in ("zml", "count"):
            def go():
                if kind == "zml":
                    ms.query_zml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_out.data_ptr(), stream)
                else:
                    ms.query_count_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_cnt.data_ptr(), d_pos.data_ptr(), stream)
            for _ in range(2):
                go()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                go()
            e1.record(); torch.cuda.synchronize(); ms.check_device_status()
            t = e0.elapsed_time(e1) / 5
            extra = {}
            if kind == "count":
                pos = d_pos.cpu().numpy()
                extra = {"mean_matched_suffix": float((150 - pos).mean())}
            print(json.dumps({"kind": kind + "_" + name, "variant": variant, "blocks_per_sm": bps, "ms": round(t, 3), "Gbases_per_s": round(n_bases / t / 1e6, 2), **extra}), flush=True)


if __name__ == "__main__":
    main()
