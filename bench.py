#!/usr/bin/env python3
"""bench.py -- PML query throughput of the B200 query path.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W
  python bench.py --workload c2 | c3-zml | c3-count       (the L2-sized bacterial pangenome of BASELINE.json configs[1] / [2])

Default workload "c4" = BASELINE.json configs[3], the largest single-GPU configuration: a synthetic human-scale move table
resident in HBM (4.6e9 rows = 36.8 GB per GPU: row indices beyond 2^32, 290x the L2), 10 kbp nanopore-like reads (5 %
substitutions, 10 % random reads), 409 600 reads = 4.1e9 bases per step per GPU -- more reads than the GPU has lanes.
A step = one PML pass over one batch of reads; consecutive steps use different batches.

One JSON line on stdout (rank 0).  `value` = bases/s with inputs resident in HBM (device entry point),
`e2e` = the same through the host-buffer C ABI call with pinned host buffers (H2D + kernel + D2H per step),
`roofline` = 32 algorithmic bytes per base (one 32-byte row sector per LF step) against the measured HBM peak, with the
measured random-sector ceiling of the same table beside it,
`cpu_baseline` = the reference's own multithreaded CPU query (oracle/_ref) on the same index and a sample of the same reads.
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic bytes per base: one 32-byte row sector per LF step (PML), two interval ends for ZML / count.
WORKLOADS = {
    "c4": dict(data="c4", index="regular-thresholds", query="pml", metric="PML query bases/sec", algo=32,
               kernel="movi_pml_kernel<uint64_t,false,5,R8,SECTOR>"),
    "c2": dict(data="c2", index="regular-thresholds", query="pml", metric="PML query bases/sec", algo=32,
               kernel="movi_pml_kernel<uint32_t,false,6,R8>"),
    "c3-zml": dict(data="c2", index="blocked-thresholds", query="zml", metric="ZML query bases/sec", algo=64,
                   kernel="movi_bs_kernel<uint32_t,ZML,R8>"),
    "c3-count": dict(data="c2", index="blocked-thresholds", query="count", metric="count query bases/sec (all bases of the reads)", algo=64,
                     kernel="movi_bs_kernel<uint32_t,COUNT,R8>"),
    # a suffix-sorted index beyond the L2 (1 Gbp of text, 5.3e7 rows = 423 MB, real LCP thresholds); sub-record only
    "g1": dict(data="g1", index="regular-thresholds", query="pml", metric="PML query bases/sec", algo=32,
               kernel="movi_pml_kernel<uint32_t,false,4,R8,SECTOR>"),
}
W = WORKLOADS["c4"]               # replaced in main() from --workload
ARGS_WORKLOAD = "c4"

C4_ROWS = int(float(os.environ.get("MOVI_C4_ROWS", "4.6e9")))       # > 2^32 rows: 36.8 GB of 8-byte rows
C4_READS = int(os.environ.get("MOVI_C4_READS", "409600"))           # per GPU per step; x 10 kbp = 4.096e9 bases (< 2^32: one device call)
C4_READ_LEN = 10_000
C4_SEED = 44
# run-length profile of the synthetic table (datagen/synth_gpu.cu).  2 = run lengths 48..64: the LF image of a run spans at most two
# rows and a walk fast-forwards 0.5 rows per base -- suffix-sorted indexes fast-forward 0.22 - 0.27 (c2, g1).  1 = heavy-tailed
# lengths, the headline of the first half of round 2: a long run's image covers many short ones there, 2.0 fast-forward rows per
# base, an artefact of arranging heavy-tailed runs at random (2.47e10 bases/s against 2.96e10: profiles/r02_bench_c4_profile1.json)
C4_PROFILE = int(os.environ.get("MOVI_C4_PROFILE", "2"))
C4_PROFILE_NOTE = {0: "geometric run lengths, mean 8", 1: "heavy-tailed run lengths: 2.0 fast-forward rows per base",
                   2: "run lengths 48..64: 0.5 fast-forward rows per base (suffix-sorted indexes: 0.22 - 0.27)"}
C4_SUB_PER_1024, C4_RAND_EVERY, C4_LEAD = 51, 10, 24                # 5 % substitutions, 1 read in 10 uniform random, 24 scattering bases
C4_CPU_SAMPLE = int(os.environ.get("MOVI_C4_CPU_SAMPLE", "4096"))   # reads per step of the CPU arm (4.1e7 bases)
N_BATCHES = 2                                                        # distinct read batches the steps cycle over


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9:
                self.rows.append(f)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for f in self.rows:
            try:
                util = float(f[4])
                if util < 20:
                    continue
                sm.append(float(f[1])); smax = float(f[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except ValueError:
                continue
        if not sm:      # nothing caught under load: report whatever was seen
            for f in self.rows:
                try:
                    sm.append(float(f[1])); smax = float(f[2])
                except ValueError:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "sampled": "during the timed region"}


# ------------------------------------------------------------------------------------------------ workloads
class C2Data:
    """configs[1] / [2]: suffix-sorted 64 x 5 Mbp bacterial pangenome (n = 6.4e8, r = 1.55e7: 124 MB of rows), 1 M x 150 bp reads.
    Also the "g1" data set: 20 x 25 Mbp (n = 1e9, r = 5.3e7: 423 MB of rows, beyond the L2)."""

    _haps = {}                       # the pangenomes, kept across the sub-records of one run

    def __init__(self, rank, device, need_device=True):
        from movi_b200.datagen import synth
        self.name = W["data"]
        self.cfg = synth.CONFIGS[self.name]
        t0 = time.time()
        d, haps = synth.prepare(self.name, (W["index"],), want_reads=self.name not in C2Data._haps, log=sys.stderr)
        haps = C2Data._haps[self.name] = haps if haps is not None else C2Data._haps[self.name]
        self.index_dir = os.path.join(d, W["index"])
        self.n_reads, self.read_len = self.cfg["n_reads"], self.cfg["read_len"]
        self.host_batches = [synth.make_reads(haps, self.n_reads, self.read_len, seed=43 + rank + 1000 * b) for b in range(N_BATCHES)]
        self.n_bases = int(self.host_batches[0][1][-1])
        log("[bench] rank %d: %s workload ready in %.1f s (%d batches of %d reads)" % (rank, self.name, time.time() - t0, N_BATCHES, self.n_reads))
        self.ms = None
        if need_device:
            import torch
            from movi_b200 import MoveStructure
            self.ms = MoveStructure(self.index_dir, device=device.index)
            pad = np.zeros((-self.n_bases) % 32 + 32, np.uint8)
            self.d_bases = [torch.from_numpy(np.concatenate([b, pad])).to(device) for b, _ in self.host_batches]
            self.d_off = torch.from_numpy(self.host_batches[0][1].astype(np.int64)).to(device)

    def describe(self):
        if self.name == "g1":
            return ("g1: 20-haplotype x 25 Mbp synthetic pangenome (0.1% divergence, +reverse complements; 1 Gbp of text), suffix-sorted with "
                    "real LCP thresholds, regular-thresholds index (5.3e7 rows = 423 MB: 3.4x the L2), 1M x 150 bp reads PML per GPU per step")
        return ("%s: 64-haplotype x 5 Mbp synthetic bacterial pangenome (0.1%% divergence, +reverse complements), suffix-sorted, %s index "
                "(124 MB of rows: L2-sized), 1M x 150 bp reads %s per GPU per step"
                % (ARGS_WORKLOAD, W["index"], W["query"].upper() if W["query"] != "count" else "count"))

    def l2_note(self):
        return ("each step streams 450 MB of inputs + outputs (> 126 MB L2) and uses a different read batch than the previous one; "
                "the row table stays resident in HBM / L2 as in production")

    def cpu_sample(self):
        bases, offsets = self.host_batches[0]
        return bases, offsets, self.n_reads, "the full per-GPU batch (%d x %d bp)" % (self.n_reads, self.read_len)

    def cpu_index_dir(self):
        return self.index_dir

    def host_batch(self, b):
        return self.host_batches[b]


class C4Data:
    """configs[3]: human-scale table synthesised in HBM (datagen/synth_gpu.cu), 10 kbp reads written down by the query's own walk."""

    def __init__(self, rank, device, need_device=True):
        import torch
        from movi_b200.datagen.synth_gpu import SyntheticStructure
        t0 = time.time()
        self.rank, self.device = rank, device
        self.n_reads, self.read_len = C4_READS, C4_READ_LEN
        self.n_bases = self.n_reads * self.read_len
        self.ms = SyntheticStructure(C4_ROWS, seed=C4_SEED, device=device.index, profile=C4_PROFILE)
        torch.cuda.synchronize()
        t1 = time.time()
        self.d_off = torch.arange(self.n_reads + 1, dtype=torch.int64, device=device) * self.read_len
        self.d_bases = []
        for b in range(N_BATCHES if need_device else 1):
            t = torch.zeros(self.n_bases + 64, dtype=torch.uint8, device=device)
            self.ms.make_reads_resync(t.data_ptr(), self.n_reads, self.read_len, C4_SUB_PER_1024, C4_RAND_EVERY, C4_LEAD, read_seed=1000 * b + rank)
            self.d_bases.append(t)
        torch.cuda.synchronize()
        self.index_dir = None
        log("[bench] rank %d: c4 table (%d rows, %.1f GB, n/r = %.1f) in %.1f s, %d read batches of %d x %d bp in %.1f s"
            % (rank, C4_ROWS, C4_ROWS * 8 / 1e9, self.ms.n / C4_ROWS, t1 - t0, len(self.d_bases), self.n_reads, self.read_len, time.time() - t1))

    def describe(self):
        return ("c4: synthetic human-scale move table resident in HBM (%.2e rows = %.1f GB per GPU, n/r = %.0f, %s, row indices beyond 2^32), "
                "regular-thresholds, %d x %d bp nanopore-like reads per GPU per step (%d %% substitutions, 1 read in %d uniform random; "
                "reads written down by the query's own walk so that they re-synchronise after an error), PML"
                % (C4_ROWS, C4_ROWS * 8 / 1e9, self.ms.n / C4_ROWS, C4_PROFILE_NOTE.get(C4_PROFILE, "profile %d" % C4_PROFILE), self.n_reads, self.read_len,
                   round(C4_SUB_PER_1024 / 10.24), C4_RAND_EVERY))

    def l2_note(self):
        return ("the 36.8 GB row table is 290x the 126 MB L2; each step streams 4.1 GB of reads in and 8.2 GB of lengths out and uses a "
                "different read batch than the previous one")

    def host_batch(self, b):
        bases = self.d_bases[b][: self.n_bases].cpu().numpy()
        return bases, (np.arange(self.n_reads + 1, dtype=np.uint64) * self.read_len)

    def cpu_sample(self):
        ns = min(C4_CPU_SAMPLE, self.n_reads)
        bases = self.d_bases[0][: ns * self.read_len].cpu().numpy()
        offsets = np.arange(ns + 1, dtype=np.uint64) * self.read_len
        return bases, offsets, ns, "the first %d reads (%d bases) of one %d-read batch" % (ns, ns * self.read_len, self.n_reads)

    def cpu_index_dir(self):
        """index.movi of the resident table, for the reference binary and the oracle: exported once per box into RAM-backed
        storage (36.8 GB), reused by the next arm that runs on the same box."""
        if self.index_dir:
            return self.index_dir
        need = C4_ROWS * 8 + (1 << 30)
        name = "movi_bench_c4_r%d_s%d_p%d" % (C4_ROWS, C4_SEED, C4_PROFILE)
        for base in ("/dev/shm", tempfile.gettempdir()):
            d = os.path.join(base, name)
            if os.path.exists(os.path.join(d, "index.movi")) and os.path.getsize(os.path.join(d, "index.movi")) > C4_ROWS * 8:
                self.index_dir = d
                return d
        for base in ("/dev/shm", tempfile.gettempdir()):
            try:
                if shutil.disk_usage(base).free > need * 1.15:
                    t0 = time.time()
                    d = self.ms.write_index_file(os.path.join(base, name))
                    log("[bench] index.movi of the c4 table written to %s in %.1f s" % (d, time.time() - t0))
                    self.index_dir = d
                    return d
            except OSError as e:
                log("[bench] cannot write the c4 index under %s: %s" % (base, e))
        return None


def write_fasta(path, bases, offsets, n_reads, repeat=1):
    L = int(offsets[1] - offsets[0])
    with open(path, "wb") as f:
        fixed = np.all(np.diff(offsets[: n_reads + 1]) == L)
        for rep in range(repeat):
            if fixed:      # fixed length: vectorised
                if rep == 0:
                    hdr = np.array([(">r%07d\n" % i).encode() for i in range(n_reads)], dtype="S10")
                    rec = np.empty((n_reads, 10 + L + 1), dtype=np.uint8)
                    rec[:, :10] = np.frombuffer(hdr.tobytes(), dtype=np.uint8).reshape(n_reads, 10)
                    rec[:, 10:10 + L] = bases[: n_reads * L].reshape(n_reads, L)
                    rec[:, -1] = 10
                    rec = rec.tobytes()
                f.write(rec)
            else:
                for i in range(n_reads):
                    f.write(b">r%07d\n" % i + bases[int(offsets[i]):int(offsets[i + 1])].tobytes() + b"\n")


def cpu_reference_run(index_dir, bases, offsets, n_reads, threads, tmpdir, repeat=1):
    """Time the reference's own CPU query (oracle/_ref, prefetch path = its fastest mode, all host threads) or,
    if that binary is not there, the C oracle with OpenMP, on `repeat` passes over the sample inside ONE process (the
    index is loaded once).  Returns (bases_per_s, kind, cores, seconds_per_pass)."""
    n_bases = int(offsets[n_reads])
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "movi-" + W["index"])
    if os.path.exists(ref_bin):
        fa = os.path.join(tmpdir, "reads_%d_x%d.fa" % (n_reads, repeat))
        if not os.path.exists(fa):
            write_fasta(fa, bases, offsets, n_reads, repeat)
        cmd = [ref_bin, "query", "--index", index_dir, "--read", fa, "--" + W["query"], "--no-output", "-t", str(threads), "-s", "16"]
        # stderr goes to a FILE: the reference prints progress inside a critical section and a pipe would
        # serialise its threads
        errf = os.path.join(tmpdir, "ref_stderr.txt")
        t0 = time.time()
        with open(errf, "w") as ef:
            r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=ef)
        wall = time.time() - t0
        r.stderr = open(errf).read()
        if r.returncode == 0:
            # "Time measured for processing the reads: <x> seconds." (reference src/movi.cpp:387-389, include/commons.hpp:31-44)
            m = re.search(r"processing the reads[^0-9]*([0-9.]+)\s*(ns|us|ms|s)?", r.stderr)
            sec = None
            if m:
                val, unit = float(m.group(1)), (m.group(2) or "s")
                sec = val * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[unit]
            if not sec or sec <= 0 or sec > wall:
                sec = wall
            log("[bench] reference binary: %d x %d reads in %.2f s (process wall %.1f s incl. index load)" % (repeat, n_reads, sec, wall))
            return n_bases * repeat / sec, "reference", threads, sec / repeat
        log("[bench] reference binary failed (%d): %s" % (r.returncode, r.stderr[-300:]))
    from oracle.port import OracleIndex
    orc = OracleIndex(index_dir)
    t0 = time.time()
    for _ in range(repeat):
        getattr(orc, "query_" + W["query"])(bases[:n_bases], offsets[: n_reads + 1], threads=threads)
    sec = time.time() - t0
    return n_bases * repeat / sec, "port", threads, sec / repeat


def sub_records(dev, peak):
    """The L2-sized configurations of BASELINE.json (configs[1] / [2]) as sub-records of the default line: device-resident
    throughput over >= 1 s of back-to-back steps, roofline with the 32 / 64 B-per-base accounting beside the DRAM bytes ncu
    measured for the same kernel (profiles/*_traffic.json), and EVERY read of the batch compared with the C oracle."""
    import torch
    from oracle.port import OracleIndex
    global W, ARGS_WORKLOAD
    keep = (W, ARGS_WORKLOAD)
    out = {}
    for name in ("c2", "c3-zml", "c3-count", "g1"):
        W, ARGS_WORKLOAD = WORKLOADS[name], name
        try:
            if name == "g1":          # its suffix array takes minutes: only when the prepared index travelled with the snapshot
                from movi_b200.datagen import synth
                f = os.path.join(synth.cache_dir("g1"), "regular-thresholds", "index.movi")
                if not (os.path.exists(f) or os.path.exists(f + ".xz")):
                    out["g1"] = {"skipped": "the prepared g1 index (_cache/g1) is not on this box"}
                    continue
            data = C2Data(0, dev)
            ms, n_reads, n_bases = data.ms, data.n_reads, data.n_bases
            is_count = W["query"] == "count"
            d_outs = [torch.zeros(n_bases + 64, dtype=torch.int16, device=dev) for _ in range(2)]
            d_cnt = [torch.zeros(n_reads, dtype=torch.int64, device=dev) for _ in range(2)]
            d_pos = [torch.zeros(n_reads, dtype=torch.int32, device=dev) for _ in range(2)]

            def launch(i, st):
                b = data.d_bases[i % len(data.d_bases)]
                if W["query"] == "pml":
                    ms.query_pml_device(b.data_ptr(), data.d_off.data_ptr(), n_reads, d_outs[i % 2].data_ptr(), st)
                elif W["query"] == "zml":
                    ms.query_zml_device(b.data_ptr(), data.d_off.data_ptr(), n_reads, d_outs[i % 2].data_ptr(), st)
                else:
                    ms.query_count_device(b.data_ptr(), data.d_off.data_ptr(), n_reads, d_cnt[i % 2].data_ptr(), d_pos[i % 2].data_ptr(), st)
            cur = torch.cuda.current_stream()
            for i in range(4):
                launch(i, cur.cuda_stream)
            torch.cuda.synchronize()
            side = [torch.cuda.Stream(device=dev) for _ in range(2)]
            steps = 400 if name != "g1" else 200
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s_ in side:
                s_.wait_stream(cur)
            for i in range(steps):
                launch(i, side[i % 2].cuda_stream)
            for s_ in side:
                cur.wait_stream(s_)
            e1.record()
            torch.cuda.synchronize()
            ms.check_device_status()
            ms_step = e0.elapsed_time(e1) / steps
            # every read of batch 0 against the oracle
            launch(0, cur.cuda_stream)
            torch.cuda.synchronize()
            bases, offsets = data.host_batch(0)
            orc = OracleIndex(data.index_dir)
            work = n_bases
            if is_count:
                wc, wp = orc.query_count(bases, offsets, threads=0)
                ok = np.array_equal(wc, d_cnt[0].cpu().numpy().view(np.uint64)) and np.array_equal(wp, d_pos[0].cpu().numpy())
                L = np.diff(offsets.astype(np.int64))
                work = int(np.minimum(L, L - wp.astype(np.int64) + 1).sum())
            else:
                want = getattr(orc, "query_" + W["query"])(bases, offsets, threads=0)
                ok = np.array_equal(want, d_outs[0][:n_bases].cpu().numpy().view(np.uint16))
            # the host-buffer entry point on this shape: pinned caller buffers against pageable ones (numpy arrays = what a
            # std::vector is to the driver: cudaMemcpyAsync stages them and blocks)
            host = None
            if name == "c2":
                hb = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).pin_memory()
                ho = torch.from_numpy(offsets.astype(np.int64)).pin_memory()
                hout = torch.empty(n_bases + 16, dtype=torch.int16).pin_memory()
                pb, po, pout = np.concatenate([bases, np.zeros(64, np.uint8)]), offsets.astype(np.uint64), np.empty(n_bases + 16, np.uint16)

                def t_host(fn, reps=10):
                    fn(); fn()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        fn()
                    return (time.perf_counter() - t0) * 1e3 / reps
                t_pin = t_host(lambda: ms._check(ms._L.movi_gpu_query_pml(ms._ctx, hb.data_ptr(), ho.data_ptr(), n_reads, 0, hout.data_ptr())))
                t_page = t_host(lambda: ms._check(ms._L.movi_gpu_query_pml(ms._ctx, pb.ctypes.data, po.ctypes.data, n_reads, 0, pout.ctypes.data)))
                ok = ok and np.array_equal(pout[:n_bases], want) and np.array_equal(hout[:n_bases].numpy().view(np.uint16), want)
                host = {"e2e": {"value": n_bases / (t_pin * 1e-3), "unit": "bases/s", "ms_per_step": t_pin, "buffers": "pinned (movi_gpu_host_alloc / cudaHostAlloc)"},
                        "e2e_pageable": {"value": n_bases / (t_page * 1e-3), "unit": "bases/s", "ms_per_step": t_page,
                                         "buffers": "pageable caller memory (numpy arrays; the driver stages every copy)"}}
                del hb, hout
            achieved = W["algo"] * work / (ms_step * 1e-3) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", {"c2": "pml_c2_traffic.json", "c3-zml": "zml_c3_traffic.json", "c3-count": "count_c3_traffic.json",
                                                 "g1": "pml_g1_traffic.json"}[name])
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            rec = {"workload": data.describe(), "metric": W["metric"], "value": n_bases / (ms_step * 1e-3), "unit": "bases/s", "ms_per_step": ms_step,
                   "steps": steps, "timed_region_s": ms_step * steps * 1e-3,
                   "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "kernel": W["kernel"],
                                "algorithmic_bytes_per_base": W["algo"], "bases_walked_per_launch": work, "traffic": traffic,
                                "measured_dram_bytes_per_base": None if not traffic else traffic / work,
                                "frac_by_measured_dram_bytes": None if not traffic else traffic / (ms_step * 1e-3) / 1e9 / peak},
                   "parity_checked": bool(ok), "parity_reads_vs_oracle": int(n_reads)}
            if host:
                rec.update(host)
            out[name.replace("-", "_")] = rec
            log("[bench] sub-record %s: %.3g bases/s, parity %s" % (name, rec["value"], ok))
            ms.close()
            del data, d_outs, d_cnt, d_pos
            torch.cuda.empty_cache()
        except Exception as e:            # a sub-record never takes the headline down with it
            out[name.replace("-", "_")] = {"error": "%s: %s" % (type(e).__name__, e)}
            log("[bench] sub-record %s failed: %s" % (name, e))
    W, ARGS_WORKLOAD = keep
    return out



def cli_file_to_file(index_dir, bases, offsets, n_reads, n_bases):
    """The same boundary as the reference arm -- a FASTA file in, PML of every base, nothing written (`--no-output`), index load not
    timed -- through the drop-in command line (`movi-b200 query`): one full step's reads, parsed, walked on the GPU from the file's
    bytes (device-side ingest).  Timer: the command's own "queried ... in X s" line, like the reference's "processing the reads"."""
    from movi_b200 import build
    out = {"unit": "bases/s", "what": "movi-b200 query --pml --no-output on one step's reads as a FASTA file (device-side ingest), index load excluded "
                                      "like in the reference arm"}
    try:
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            fa = os.path.join(td, "step.fa")
            write_fasta(fa, bases, offsets, n_reads)
            t0 = time.time()
            r = subprocess.run([build.CLI, "query", "-i", index_dir, "-r", fa, "--pml", "--no-output", "--verbose"], stdout=subprocess.DEVNULL,
                               stderr=subprocess.PIPE, text=True)
            wall = time.time() - t0
            m = re.search(r"written in ([0-9.]+) s: (\d+) reads / (\d+) bases", r.stderr)
            if r.returncode != 0 or not m or int(m.group(3)) != n_bases:
                out["error"] = (r.stderr or "")[-300:]
                return out
            sec = float(m.group(1))
            out.update({"value": n_bases / sec, "seconds": sec, "process_wall_s": round(wall, 2), "fasta_bytes": os.path.getsize(fa)})
            log("[bench] movi-b200 query on one step's FASTA: %.3f s (process wall %.1f s incl. CUDA start-up and index upload)" % (sec, wall))
    except Exception as e:
        out["error"] = "%s: %s" % (type(e).__name__, e)
    return out


def make_data(rank, device, need_device=True):
    return C4Data(rank, device, need_device) if W["data"] == "c4" else C2Data(rank, device, need_device)


def run_reference(args, rank, world):
    """The reference's own CPU implementation on the same index and a bounded sample of the same reads: W warm-up passes in
    one process, then K timed passes in another (the index is loaded once per process; its load is not timed)."""
    if rank != 0:
        return 0
    device = None
    if W["data"] == "c4":
        import torch
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: the c4 table is synthesised on the GPU; no CUDA device")
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(device)
    data = make_data(0, device, need_device=False)
    bases, offsets, n_reads, what = data.cpu_sample()
    index_dir = data.cpu_index_dir()
    if index_dir is None:
        emit_line({"impl": "reference", "unavailable": "no room to export the c4 index for the CPU arm (needs 37 GB of /dev/shm or /tmp)"})
        return 0
    if data.ms is not None:
        data.ms.close()                  # the GPU was only the data generator
        data.d_bases = None
    threads = os.cpu_count() or 1
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        if args.warmup:
            cpu_reference_run(index_dir, bases, offsets, n_reads, threads, td, repeat=min(args.warmup, 3))
        value, kind, cores, sec = cpu_reference_run(index_dir, bases, offsets, n_reads, threads, td, repeat=args.steps)
    sample = "%s per step; %d steps back to back inside one reference process (index loaded once, load not timed)" % (what, args.steps)
    line = {
        "impl": "reference", "metric": W["metric"], "value": value, "unit": "bases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(sec * 1e3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": workload_config(data, None),
        "cpu_baseline": {"value": value, "unit": "bases/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


def workload_config(data, index_bytes):
    c = {"workload": data.describe(), "index": W["index"], "reads_per_gpu": data.n_reads, "read_len": data.read_len,
         "l2": data.l2_note(),
         "launch": "steps alternate over 2 CUDA streams (batch i+1 ramps up while batch i drains); one kernel launch per step",
         "parallelism": "reads sharded per GPU, index replicated, no collective on the query path"}
    return c


def emit_line(line):
    """The one JSON line of this run, on the process's ORIGINAL stdout (see main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # stdout carries exactly one line, the JSON record: libraries that write to fd 1 on their own (NCCL prints its version there
    # on the first communicator) are sent to stderr for the duration of the run
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary records (classify, ml8, c2 / c3 sub-records)")
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global W, ARGS_WORKLOAD
    W, ARGS_WORKLOAD = WORKLOADS[args.workload], args.workload
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the query path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    # rank 0 prepares first so that the ranks do not race on a cached index
    if rank == 0:
        data = make_data(rank, dev)
    barrier()
    if rank != 0:
        data = make_data(rank, dev)
    ms = data.ms
    n_reads, n_bases = data.n_reads, data.n_bases
    is_count = W["query"] == "count"
    long_reads = data.read_len > 255

    d_off = data.d_off
    d_outs = [torch.zeros(n_bases + 64, dtype=torch.int16, device=dev) for _ in range(2)] if not is_count else [None, None]
    d_cnts = [torch.zeros(n_reads, dtype=torch.int64, device=dev) for _ in range(2)]
    d_poss = [torch.zeros(n_reads, dtype=torch.int32, device=dev) for _ in range(2)]

    def launch(i, cuda_stream):
        b = data.d_bases[i % len(data.d_bases)]
        if W["query"] == "pml":
            ms.query_pml_device(b.data_ptr(), d_off.data_ptr(), n_reads, d_outs[i % 2].data_ptr(), cuda_stream)
        elif W["query"] == "zml":
            ms.query_zml_device(b.data_ptr(), d_off.data_ptr(), n_reads, d_outs[i % 2].data_ptr(), cuda_stream)
        else:
            ms.query_count_device(b.data_ptr(), d_off.data_ptr(), n_reads, d_cnts[i % 2].data_ptr(), d_poss[i % 2].data_ptr(), cuda_stream)

    stream = torch.cuda.current_stream().cuda_stream
    for i in range(args.warmup):
        launch(i, stream)
    torch.cuda.synchronize()
    ms.check_device_status()

    # ---- timed region: inputs resident in HBM.  Steps alternate between two streams so that the drain of one batch (lanes
    # finishing their last reads) overlaps the ramp-up of the next one, as in a production stream of batches; every step is a
    # complete pass over its batch.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier(); torch.cuda.synchronize()
    l0 = ms.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    side = [torch.cuda.Stream(device=dev) for _ in range(2)]
    e0.record()
    for s_ in side:
        s_.wait_stream(torch.cuda.current_stream())
    for i in range(args.steps):
        launch(i, side[i % 2].cuda_stream)
    for s_ in side:
        torch.cuda.current_stream().wait_stream(s_)
    e1.record()
    torch.cuda.synchronize(); barrier()
    launches = ms.launch_count - l0
    elapsed_ms = e0.elapsed_time(e1)
    ms.check_device_status()
    if sampler is not None and elapsed_ms < 1000.0:
        # a timed region shorter than a few nvidia-smi samples (the L2-sized workloads): keep the same load on until one second
        # has been sampled -- the clocks then describe this load, the value above is from the K timed steps only
        t_end = time.time() + 1.0
        while time.time() < t_end:
            for i in range(10):
                launch(i, stream)
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None

    # last device-resident result of batch 0 (for the checks below)
    launch(0, stream)
    torch.cuda.synchronize()

    # ---- e2e: host pinned buffers through the public host entry point (H2D of the reads + walk + D2H of the results per step)
    h_bases_np, h_off_np = data.host_batch(0)
    pad = (-n_bases) % 32 + 32
    h_bases = torch.empty(n_bases + pad, dtype=torch.uint8).pin_memory()
    h_bases[:n_bases].copy_(torch.from_numpy(h_bases_np)); h_bases[n_bases:].zero_()
    h_off = torch.from_numpy(h_off_np.astype(np.int64)).pin_memory()
    h_out = torch.empty(n_bases + 16, dtype=torch.int16).pin_memory() if not is_count else None
    h_cnt, h_pos = torch.zeros(n_reads, dtype=torch.int64).pin_memory(), torch.zeros(n_reads, dtype=torch.int32).pin_memory()
    e2e_steps = args.steps if n_bases < (1 << 30) else max(3, min(args.steps, 6 if world == 1 else 3))

    def step_e2e():
        L, c = ms._L, ms._ctx
        if W["query"] == "pml":
            ms._check(L.movi_gpu_query_pml(c, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_out.data_ptr()))
        elif W["query"] == "zml":
            ms._check(L.movi_gpu_query_zml(c, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_out.data_ptr()))
        else:
            ms._check(L.movi_gpu_query_count(c, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_cnt.data_ptr(), h_pos.data_ptr()))

    def time_host(fn, steps):
        for _ in range(2 if n_bases < (1 << 30) else 1):
            fn()
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        ms_ = (time.perf_counter() - t0) * 1e3 / steps
        barrier()
        return ms_

    e2e_ms = time_host(step_e2e, e2e_steps)

    # consistency inside the bench: the device-resident result equals the result of the host-buffer entry point
    if is_count:
        same = bool(torch.equal(d_cnts[0].cpu(), h_cnt) and torch.equal(d_poss[0].cpu(), h_pos))
    else:
        same = bool(torch.equal(d_outs[0][: n_bases].cpu(), h_out[: n_bases]))

    # ---- extra: PML + binary classification through the fused host entry point -- only 16 bytes per read come back
    e2e_cls_ms = None
    if W["query"] == "pml" and not args.no_extras:
        h_above, h_below = torch.zeros(n_reads, dtype=torch.int32).pin_memory(), torch.zeros(n_reads, dtype=torch.int32).pin_memory()
        h_sum = torch.zeros(n_reads, dtype=torch.int64).pin_memory()

        def step_cls():
            ms._check(ms._L.movi_gpu_query_classify(ms._ctx, 0, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, 150, 7,
                                                    h_above.data_ptr(), h_below.data_ptr(), h_sum.data_ptr(), None))
        e2e_cls_ms = time_host(step_cls, e2e_steps)
        found = int(((h_above.numpy().astype(np.int64) * 2) > (h_above.numpy().astype(np.int64) + h_below.numpy())).sum())

    # ---- extra (150 bp reads): byte-wide results (movi_gpu_query_ml8) -- half the copy-back of the headline e2e
    e2e_ml8_ms = None
    if not is_count and not long_reads and not args.no_extras:
        h_out8 = torch.empty(n_bases + 16, dtype=torch.uint8).pin_memory()

        def step_ml8():
            ms._check(ms._L.movi_gpu_query_ml8(ms._ctx, 1 if W["query"] == "zml" else 0, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0,
                                               h_out8.data_ptr()))
        e2e_ml8_ms = time_host(step_ml8, e2e_steps)
        same = same and bool(torch.equal(h_out8[: n_bases].to(torch.int16), h_out[: n_bases]))

    # ---- the random-sector ceiling of THIS table: dependent random 32-byte loads over the resident rows (movi_gpu_gather_probe)
    probe = None
    if rank == 0:
        try:
            sec, loads = ms.gather_probe(0, 32, True, 256, 8, 256)
            probe = loads / sec
        except Exception as e:         # compact layouts have no 8-byte table to probe
            log("[bench] probe skipped: %s" % e)

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms, e2e_cls_ms or 0.0, e2e_ml8_ms or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = float(t[0]), float(t[1])
        e2e_cls_ms = float(t[2]) if e2e_cls_ms is not None else None
        e2e_ml8_ms = float(t[3]) if e2e_ml8_ms is not None else None
        ok = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())

    if rank == 0:
        value = n_bases * world * args.steps / (elapsed_ms * 1e-3)
        e2e_value = n_bases * world / (e2e_ms * 1e-3)
        peak, peak_src = measured_peaks()
        per_launch_s = elapsed_ms * 1e-3 / args.steps
        work_bases = n_bases
        if is_count:       # the backward search of a read stops at its first failed extension: count the steps it really took
            L = np.diff(h_off_np.astype(np.int64))
            work_bases = int(np.minimum(L, L - h_pos.numpy().astype(np.int64) + 1).sum())
        achieved = W["algo"] * work_bases / per_launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", {"c4": "pml_c4_p%d_traffic.json" % C4_PROFILE, "c2": "pml_c2_traffic.json", "c3-zml": "zml_c3_traffic.json",
                                            "c3-count": "count_c3_traffic.json", "g1": "pml_g1_traffic.json"}[args.workload])
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": W["kernel"], "algorithmic_bytes_per_base": W["algo"], "bases_walked_per_launch": int(work_bases)}
        if probe:
            roof["random_sector_ceiling"] = {
                "sectors_per_s": probe, "GBps": probe * 32 / 1e9, "frac_of_hbm_peak": probe * 32 / 1e9 / peak,
                "achieved_frac_of_ceiling": achieved / (probe * 32 / 1e9) / (W["algo"] / 32),
                "what": "dependent random 32-byte loads over the same resident table, all lanes busy (movi_gpu_gather_probe): what the memory "
                        "system delivers for this access pattern; the cost is per sector request, whatever its DRAM page or TLB locality "
                        "(profiles/r02_tlb_probe.jsonl)"}
        line = {
            "metric": W["metric"], "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": workload_config(data, None),            # (identical to the reference arm's: the driver compares them)
            "index_device_bytes": int(ms.index_device_bytes),
            "roofline": roof,
            "e2e": {"value": e2e_value, "unit": "bases/s", "h2d_bytes_per_step": int(n_bases + 8 * (n_reads + 1)),
                    "d2h_bytes_per_step": int(12 * n_reads if is_count else 2 * n_bases), "ms_per_step": e2e_ms, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "e2e_classify": None if e2e_cls_ms is None else {
                "value": n_bases * world / (e2e_cls_ms * 1e-3), "unit": "bases/s", "ms_per_step": e2e_cls_ms,
                "h2d_bytes_per_step": int(n_bases + 8 * (n_reads + 1)), "d2h_bytes_per_step": int(16 * n_reads), "reads_found_rank0": found,
                "what": "PML + Classifier::classify fused on the device (movi_gpu_query_classify): --classify / --filter workflows"},
            "e2e_ml8": None if e2e_ml8_ms is None else {
                "value": n_bases * world / (e2e_ml8_ms * 1e-3), "unit": "bases/s", "ms_per_step": e2e_ml8_ms,
                "h2d_bytes_per_step": int(n_bases + 8 * (n_reads + 1)), "d2h_bytes_per_step": int(n_bases),
                "what": "same query through movi_gpu_query_ml8: reads <= 255 bases, lengths returned as uint8 (identical values)"},
            "clocks": clocks,
            "parity_checked": same,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            index_dir = data.cpu_index_dir()
            if index_dir is None:
                line["cpu_baseline"] = {"value": None, "unit": "bases/s", "cores": threads, "kind": "reference",
                                        "sample": "skipped: no room to export the index for the CPU arm"}
            else:
                bases, offsets, ns, what = data.cpu_sample()
                with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
                    # (several passes inside one process: a single pass of the c4 sample lasts 0.3 s, the index load 20 s)
                    v, kind, cores, sec = cpu_reference_run(index_dir, bases, offsets, ns, threads, td, repeat=8 if W["data"] == "c4" else 1)
                # the CPU leg doubles as the checker: the GPU result of the whole sample against the C oracle
                from oracle.port import OracleIndex
                orc = OracleIndex(index_dir)
                nb = int(offsets[ns])
                t0 = time.time()
                if is_count:
                    wc, wp = orc.query_count(bases[:nb], offsets[: ns + 1], threads=0)
                    ok = np.array_equal(wc, h_cnt[:ns].numpy().view(np.uint64)) and np.array_equal(wp, h_pos[:ns].numpy())
                else:
                    want = getattr(orc, "query_" + W["query"])(bases[:nb], offsets[: ns + 1], threads=0)
                    ok = np.array_equal(want, h_out[:nb].numpy().view(np.uint16))
                log("[bench] oracle check of %d reads (%d bases): %s in %.1f s" % (ns, nb, "equal" if ok else "DIFFERENT", time.time() - t0))
                line["parity_checked"] = bool(same and ok)
                line["parity_reads_vs_oracle"] = int(ns)
                line["cpu_baseline"] = {"value": v, "unit": "bases/s", "cores": cores, "kind": kind,
                                        "sample": "%s, %.2f s per pass on %d threads%s, index load excluded"
                                                  % (what, sec, cores, ", 8 passes" if W["data"] == "c4" else "")}
        if world == 1 and args.workload == "c4" and not args.no_extras:
            ms.close()
            data.d_bases = None
            del d_outs, h_bases, h_out
            torch.cuda.empty_cache()
            if not args.no_cpu_baseline and data.cpu_index_dir():
                line["e2e_cli"] = cli_file_to_file(data.cpu_index_dir(), h_bases_np, h_off_np, n_reads, n_bases)
            line.update(sub_records(dev, peak))
        emit_line(line)
    ms.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
