#!/usr/bin/env python3
"""PML throughput on synthetic HBM-resident tables of growing size (datagen/synth_gpu.cu).
usage: explore_big.py <table GB> [...]     writes gpurun_out/explore_big.jsonl"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)


def main():
    import torch
    from movi_b200.datagen.synth_gpu import SyntheticStructure
    dev = torch.device("cuda:0")
    for spec in sys.argv[1:]:
        gb, profile = (spec.split(':') + ['0'])[:2]
        gb, profile = float(gb), int(profile)
        r = int(gb * (1 << 30) / 8)
        t0 = time.time()
        ms = SyntheticStructure(r, seed=44, profile=profile)
        torch.cuda.synchronize()
        t_build = time.time() - t0
        for (n_reads, L, sub, rnd) in ((1_000_000, 150, 51, 10), (1_000_000, 150, 5, 50), (40_000, 10_000, 5, 0)):
            n_bases = n_reads * L
            d_bases = torch.zeros(n_bases + 64, dtype=torch.uint8, device=dev)
            ms.make_reads(d_bases.data_ptr(), n_reads, L, sub, rnd)
            d_off = (torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * L)
            d_out = torch.zeros(n_bases + 64, dtype=torch.int16, device=dev)
            stream = torch.cuda.current_stream().cuda_stream
            for sector in (0, 1):
                ms.set_option("pml_sector", sector)
                for _ in range(2):
                    ms.query_pml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_out.data_ptr(), stream)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record()
                for _ in range(reps):
                    ms.query_pml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_out.data_ptr(), stream)
                e1.record(); torch.cuda.synchronize(); ms.check_device_status()
                t = e0.elapsed_time(e1) / reps
                mean_pml = float(d_out[:n_bases].float().mean().item())
                rec = {"kind": "pml_synth", "table_gb": gb, "run_profile": profile, "rows": r, "n_over_r": round(ms.n / r, 1), "wide_idx": r > 0xFFFFFFF0,
                       "pml_sector": sector, "n_reads": n_reads, "read_len": L, "sub_per_1024": sub, "random_every": rnd,
                       "ms": round(t, 3), "Gbases_per_s": round(n_bases / t / 1e6, 2), "mean_pml": round(mean_pml, 1)}
                print(json.dumps(rec), flush=True)
                open(os.path.join(OUT, "explore_big.jsonl"), "a").write(json.dumps(rec) + "\n")
            if os.environ.get("MOVI_EXPLORE_BS"):        # count / ZML on the same table and reads
                d_cnt = torch.zeros(n_reads, dtype=torch.int64, device=dev); d_pos = torch.zeros(n_reads, dtype=torch.int32, device=dev)
                for kind in ("zml", "count"):
                    def go():
                        if kind == "zml": ms.query_zml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_out.data_ptr(), stream)
                        else: ms.query_count_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_cnt.data_ptr(), d_pos.data_ptr(), stream)
                    go(); go(); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); go(); go(); go(); e1.record(); torch.cuda.synchronize(); ms.check_device_status()
                    t = e0.elapsed_time(e1) / 3
                    rec = {"kind": kind + "_synth", "table_gb": gb, "n_reads": n_reads, "read_len": L, "sub_per_1024": sub, "random_every": rnd,
                           "ms": round(t, 3), "Gbases_per_s": round(n_bases / t / 1e6, 2)}
                    if kind == "count": rec["mean_matched"] = float((L - d_pos.float()).mean().item())
                    print(json.dumps(rec), flush=True)
                    open(os.path.join(OUT, "explore_big.jsonl"), "a").write(json.dumps(rec) + "\n")
            del d_bases, d_out, d_off
        sec, loads = ms.gather_probe(0, 8, True, 256, 8, 256)
        rec = {"kind": "probe_on_synth_table", "table_gb": gb, "loads_per_s": loads / sec}
        print(json.dumps(rec), flush=True)
        open(os.path.join(OUT, "explore_big.jsonl"), "a").write(json.dumps(rec) + "\n")
        ms.close(); torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
