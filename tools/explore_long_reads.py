#!/usr/bin/env python3
"""PML / classify on the long-read shape of configs[3] / [4]: 100 k x 10 kbp reads (5 % errors) against the c2 pangenome index."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from movi_b200 import MoveStructure
    from movi_b200.datagen import synth
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    L = 10_000
    d, haps = synth.prepare("c2", ("regular-thresholds",))
    t0 = time.time()
    bases, offsets = synth.make_reads(haps, n_reads, L, seed=45, sub=0.05, frac_random=0.0)
    print("reads made in %.1f s" % (time.time() - t0), flush=True)
    ms = MoveStructure(os.path.join(d, "regular-thresholds"))
    dev = torch.device("cuda:0")
    d_b = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).to(dev)
    d_o = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    d_out = torch.zeros(len(bases) + 64, dtype=torch.int16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for name, fn in (("pml", lambda: ms.query_pml_device(d_b.data_ptr(), d_o.data_ptr(), n_reads, d_out.data_ptr(), st)),
                     ("zml", lambda: ms.query_zml_device(d_b.data_ptr(), d_o.data_ptr(), n_reads, d_out.data_ptr(), st))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize(); ms.check_device_status()
        t = e0.elapsed_time(e1) / 2
        print(json.dumps({"kind": name + "_long_reads_c2", "n_reads": n_reads, "read_len": L, "ms": round(t, 2), "Gbases_per_s": round(len(bases) / t / 1e6, 2),
                          "mean_ml": round(float(d_out[: len(bases)].float().mean().item()), 1)}), flush=True)


if __name__ == "__main__":
    main()
