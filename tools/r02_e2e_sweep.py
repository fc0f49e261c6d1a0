#!/usr/bin/env python3
"""c4 through the host-buffer entry point: slice size (batch_bases) x concurrently walking slices (kernel_streams).
Writes gpurun_out/r02_e2e_sweep.jsonl."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)


def main():
    import torch
    import bench
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    data = bench.C4Data(0, dev, need_device=False)
    ms, n_reads, n_bases = data.ms, data.n_reads, data.n_bases
    h_bases = torch.empty(n_bases + 64, dtype=torch.uint8).pin_memory()
    h_bases[:n_bases].copy_(data.d_bases[0][:n_bases].cpu()); h_bases[n_bases:].zero_()
    h_off = (torch.arange(n_reads + 1, dtype=torch.int64) * data.read_len).pin_memory()
    h_out = torch.empty(n_bases + 16, dtype=torch.int16).pin_memory()
    first = None
    for ks, bb in ((4, 12 << 20), (4, 760_000_000), (4, 1_520_000_000), (2, 760_000_000), (2, 1_520_000_000), (1, 1_520_000_000), (2, 2_100_000_000),
                   (3, 500_000_000), (4, 250_000_000)):
        ms.set_option("kernel_streams", ks)
        ms.set_option("batch_bases", bb)
        fn = lambda: ms._check(ms._L.movi_gpu_query_pml(ms._ctx, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_out.data_ptr()))
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        t = (time.perf_counter() - t0) / 3 * 1e3
        chk = int(h_out[:100000].to(torch.int64).sum()) + int(h_out[n_bases - 100000:n_bases].to(torch.int64).sum())
        first = chk if first is None else first
        rec = {"kernel_streams": ks, "batch_bases": bb, "ms_per_step": round(t, 1), "Gbases_per_s": round(n_bases / t / 1e6, 2), "same": chk == first}
        print(json.dumps(rec), flush=True)
        open(os.path.join(OUT, "r02_e2e_sweep.jsonl"), "a").write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
