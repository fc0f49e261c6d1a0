#!/usr/bin/env python3
"""bench.py -- PML query throughput of the B200 query path on BASELINE.json's configs[1].

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload ("c2"): synthetic 64-haplotype bacterial pangenome (5 Mbp base genome, 0.1 % SNP/indel divergence, every
haplotype + its reverse complement; n = 6.4e8), regular-thresholds index, 1 000 000 x 150 bp reads per GPU
(0.5 % substitutions, 2 % random reads).  A step = one PML pass over the rank's 1 M reads.

One JSON line on stdout (rank 0).  `value` = bases/s with inputs resident in HBM (device entry point),
`e2e` = the same through the host-buffer C ABI call with pinned host buffers (H2D + kernel + D2H per step),
`roofline` = 32 algorithmic bytes per base (one 32-byte row sector per LF step) against the measured HBM peak,
`cpu_baseline` = the reference's own multithreaded CPU query (oracle/_ref) or the C oracle, same inputs, same box.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "c2"                   # synthetic data set (movi_b200/datagen/synth.py); every workload below queries it
# --workload: c2 is the bench line of BASELINE.json (configs[1]); the c3-* lines are configs[2] (same pangenome, blocked-thresholds
# index, ZML / count).  Algorithmic bytes per base: one 32-byte row sector per LF step (PML), two interval ends for ZML / count.
WORKLOADS = {
    "c2": dict(index="regular-thresholds", query="pml", metric="PML query bases/sec", algo=32, kernel="movi_pml_kernel4<uint32_t,false,6,R8>"),
    "c3-zml": dict(index="blocked-thresholds", query="zml", metric="ZML query bases/sec", algo=64, kernel="movi_bs_kernel2<uint32_t,ZML,R8>"),
    "c3-count": dict(index="blocked-thresholds", query="count", metric="count query bases/sec (all bases of the reads)", algo=64,
                     kernel="movi_bs_kernel2<uint32_t,COUNT,R8>"),
}
W = WORKLOADS["c2"]               # replaced in main() from --workload
ARGS_WORKLOAD = "c2"


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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def prepare_workload(rank, need_haps=True):
    from movi_b200.datagen import synth
    cfg = synth.CONFIGS[WORKLOAD]
    t0 = time.time()
    d, haps = synth.prepare(WORKLOAD, (W["index"],), want_reads=need_haps, log=sys.stderr)
    bases, offsets = synth.make_reads(haps, cfg["n_reads"], cfg["read_len"], seed=43 + rank)
    log("[bench] rank %d: workload ready in %.1f s (%d reads, %d bases)" % (rank, time.time() - t0, cfg["n_reads"], len(bases)))
    return os.path.join(d, W["index"]), bases, offsets, cfg


def write_fasta(path, bases, offsets, n_reads):
    L = int(offsets[1] - offsets[0])
    with open(path, "wb") as f:
        if np.all(np.diff(offsets[: n_reads + 1]) == L):      # fixed length: vectorised
            hdr = np.array([(">r%07d\n" % i).encode() for i in range(n_reads)], dtype="S10")
            rec = np.empty((n_reads, 10 + L + 1), dtype=np.uint8)
            rec[:, :10] = np.frombuffer(hdr.tobytes(), dtype=np.uint8).reshape(n_reads, 10)
            rec[:, 10:10 + L] = bases[: n_reads * L].reshape(n_reads, L)
            rec[:, -1] = 10
            f.write(rec.tobytes())
        else:
            for i in range(n_reads):
                f.write(b">r%07d\n" % i + bases[int(offsets[i]):int(offsets[i + 1])].tobytes() + b"\n")


def cpu_reference_run(index_dir, bases, offsets, n_reads, threads, tmpdir):
    """Time the reference's own CPU query (oracle/_ref, prefetch path = its fastest mode, all host threads) or,
    if that binary is not there, the C oracle with OpenMP.  Returns (bases_per_s, kind, cores, seconds)."""
    n_bases = int(offsets[n_reads])
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "movi-" + W["index"])
    if os.path.exists(ref_bin):
        fa = os.path.join(tmpdir, "reads_%d.fa" % n_reads)
        if not os.path.exists(fa):
            write_fasta(fa, bases, offsets, n_reads)
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
            return n_bases / sec, "reference", threads, sec
        log("[bench] reference binary failed (%d): %s" % (r.returncode, r.stderr[-300:]))
    from oracle.port import OracleIndex
    orc = OracleIndex(index_dir)
    t0 = time.time()
    getattr(orc, "query_" + W["query"])(bases[:n_bases], offsets[: n_reads + 1], threads=threads)
    sec = time.time() - t0
    return n_bases / sec, "port", threads, sec


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    index_dir, bases, offsets, cfg = prepare_workload(0)
    threads = os.cpu_count() or 1
    n_reads = cfg["n_reads"]
    vals, kind, secs = [], "port", []
    with tempfile.TemporaryDirectory() as td:
        for i in range(args.warmup + args.steps):
            v, kind, cores, sec = cpu_reference_run(index_dir, bases, offsets, n_reads, threads, td)
            if i >= args.warmup:
                vals.append(v); secs.append(sec)
    value = float(np.sum([int(offsets[n_reads])] * len(secs)) / np.sum(secs))
    sample = "%d x %d bp reads of the c2 workload per step (the full per-GPU batch), index load excluded" % (n_reads, cfg["read_len"])
    line = {
        "impl": "reference", "metric": W["metric"], "value": value, "unit": "bases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs) * 1e3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": workload_config(cfg, None),
        "cpu_baseline": {"value": value, "unit": "bases/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


def workload_config(cfg, index_bytes):
    c = {"workload": "%s: 64-haplotype x 5 Mbp synthetic bacterial pangenome (0.1%% divergence, +reverse complements), "
                     "%s index, 1M x 150 bp reads %s per GPU" % (ARGS_WORKLOAD, W["index"], W["query"].upper() if W["query"] != "count" else "count"),
         "index": W["index"], "reads_per_gpu": cfg["n_reads"], "read_len": cfg["read_len"],
         "l2": "per-step streamed inputs+outputs (450 MB) exceed the 126 MB L2; the row table stays resident in HBM/L2 as in production",
         "launch": "steps alternate over 2 CUDA streams (batch i+1 ramps up while batch i drains); one kernel launch per step",
         "parallelism": "reads sharded per GPU, index replicated, no collective on the query path"}
    if index_bytes is not None:
        c["index_device_bytes"] = int(index_bytes)
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
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
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
    from movi_b200 import MoveStructure

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

    # rank 0 builds the cached index first so that the ranks do not race on it
    if rank == 0:
        index_dir, bases, offsets, cfg = prepare_workload(rank)
    barrier()
    if rank != 0:
        index_dir, bases, offsets, cfg = prepare_workload(rank)
    n_reads, n_bases = cfg["n_reads"], int(offsets[-1])

    ms = MoveStructure(index_dir, device=local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    pad = (-n_bases) % 32 + 32
    h_bases = torch.from_numpy(np.concatenate([bases, np.zeros(pad, np.uint8)])).pin_memory()
    h_off = torch.from_numpy(offsets.astype(np.int64)).pin_memory()
    h_out = torch.empty(n_bases + 16, dtype=torch.int16).pin_memory()
    d_bases = h_bases.to(dev)
    d_off = h_off.to(dev)
    d_out = torch.zeros(n_bases + 16, dtype=torch.int16, device=dev)
    is_count = W["query"] == "count"
    h_cnt, h_pos = torch.zeros(n_reads, dtype=torch.int64).pin_memory(), torch.zeros(n_reads, dtype=torch.int32).pin_memory()
    d_cnts = [torch.zeros(n_reads, dtype=torch.int64, device=dev) for _ in range(2)]
    d_poss = [torch.zeros(n_reads, dtype=torch.int32, device=dev) for _ in range(2)]
    d_outs = [d_out, torch.zeros_like(d_out)]

    def launch(i, cuda_stream):
        if W["query"] == "pml":
            ms.query_pml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_outs[i % 2].data_ptr(), cuda_stream)
        elif W["query"] == "zml":
            ms.query_zml_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_outs[i % 2].data_ptr(), cuda_stream)
        else:
            ms.query_count_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, d_cnts[i % 2].data_ptr(), d_poss[i % 2].data_ptr(), cuda_stream)

    def step_device():
        launch(0, stream)

    def step_e2e():
        L, c = ms._L, ms._ctx
        if W["query"] == "pml":
            ms._check(L.movi_gpu_query_pml(c, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_out.data_ptr()))
        elif W["query"] == "zml":
            ms._check(L.movi_gpu_query_zml(c, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_out.data_ptr()))
        else:
            ms._check(L.movi_gpu_query_count(c, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, h_cnt.data_ptr(), h_pos.data_ptr()))

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    ms.check_device_status()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- timed region: inputs resident in HBM
    barrier(); torch.cuda.synchronize()
    l0 = ms.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # steps alternate between two streams so that the drain of one batch (lanes finishing their last reads) overlaps
    # the ramp-up of the next one, as in a production stream of batches; every step is a complete pass over the batch
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

    # ---- e2e: host pinned buffers through the public host entry point
    for _ in range(2):
        step_e2e()
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    # ---- extra (c2 only): PML + binary classification through the fused host entry point -- only 16 bytes per read come back
    e2e_cls_ms = None
    if W["query"] == "pml":
        h_above, h_below = torch.zeros(n_reads, dtype=torch.int32).pin_memory(), torch.zeros(n_reads, dtype=torch.int32).pin_memory()
        h_sum = torch.zeros(n_reads, dtype=torch.int64).pin_memory()

        def step_cls():
            ms._check(ms._L.movi_gpu_query_classify(ms._ctx, 0, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0, 150, 7,
                                                    h_above.data_ptr(), h_below.data_ptr(), h_sum.data_ptr(), None))
        for _ in range(2):
            step_cls()
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_cls()
        torch.cuda.synchronize()
        e2e_cls_ms = (time.perf_counter() - t0) * 1e3
        barrier()

    # ---- extra (150 bp reads): byte-wide results (movi_gpu_query_ml8) -- half the copy-back of the headline e2e
    e2e_ml8_ms, ml8_same = None, True
    if not is_count:
        h_out8 = torch.empty(n_bases + 16, dtype=torch.uint8).pin_memory()

        def step_ml8():
            ms._check(ms._L.movi_gpu_query_ml8(ms._ctx, 1 if W["query"] == "zml" else 0, h_bases.data_ptr(), h_off.data_ptr(), n_reads, 0,
                                               h_out8.data_ptr()))
        for _ in range(2):
            step_ml8()
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_ml8()
        torch.cuda.synchronize()
        e2e_ml8_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        ml8_same = bool(torch.equal(h_out8[: n_bases].to(torch.int16), h_out[: n_bases]))

    # keep the GPU under the same load a little longer so that the clock sampler sees it (not part of any number)
    if sampler is not None:
        t_end = time.time() + 1.2
        while time.time() < t_end:
            for _ in range(10):
                step_device()
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None

    # consistency inside the bench: the device-resident result equals the result of the host-buffer entry point
    if is_count:
        same = bool(torch.equal(d_cnts[0].cpu(), h_cnt) and torch.equal(d_poss[0].cpu(), h_pos))
    else:
        same = bool(torch.equal(d_out[: n_bases].cpu(), h_out[: n_bases])) and ml8_same

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
        total_bases = n_bases * world * args.steps
        value = total_bases / (elapsed_ms * 1e-3)
        e2e_value = total_bases / (e2e_ms * 1e-3)
        peak, peak_src = measured_peaks()
        per_launch_s = elapsed_ms * 1e-3 / args.steps
        work_bases = n_bases
        if is_count:       # the backward search of a read stops at its first failed extension: count the steps it really took
            L = np.diff(offsets.astype(np.int64))
            work_bases = int(np.minimum(L, L - h_pos.numpy().astype(np.int64) + 1).sum())
        achieved = W["algo"] * work_bases / per_launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", {"c2": "pml_c2_traffic.json", "c3-zml": "zml_c3_traffic.json", "c3-count": "count_c3_traffic.json"}[args.workload])
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": W["metric"], "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": workload_config(cfg, ms.index_device_bytes),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": W["kernel"], "algorithmic_bytes_per_base": W["algo"], "bases_walked_per_launch": int(work_bases)},
            "e2e": {"value": e2e_value, "unit": "bases/s", "h2d_bytes_per_step": int(n_bases + 8 * (n_reads + 1)),
                    "d2h_bytes_per_step": int(12 * n_reads if is_count else 2 * n_bases), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "e2e_classify": None if e2e_cls_ms is None else {
                "value": total_bases / (e2e_cls_ms * 1e-3), "unit": "bases/s", "ms_per_step": e2e_cls_ms / args.steps,
                "h2d_bytes_per_step": int(n_bases + 8 * (n_reads + 1)), "d2h_bytes_per_step": int(16 * n_reads),
                "what": "PML + Classifier::classify fused on the device (movi_gpu_query_classify): --classify / --filter workflows"},
            "e2e_ml8": None if e2e_ml8_ms is None else {
                "value": total_bases / (e2e_ml8_ms * 1e-3), "unit": "bases/s", "ms_per_step": e2e_ml8_ms / args.steps,
                "h2d_bytes_per_step": int(n_bases + 8 * (n_reads + 1)), "d2h_bytes_per_step": int(n_bases),
                "what": "same query through movi_gpu_query_ml8: reads <= 255 bases, lengths returned as uint8 (identical values)"},
            "clocks": clocks,
            "parity_checked": same,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            with tempfile.TemporaryDirectory() as td:
                v, kind, cores, sec = cpu_reference_run(index_dir, bases, offsets, n_reads, threads, td)
            # the CPU leg doubles as the checker: the first 2000 reads of the GPU result against the C oracle
            from oracle.port import OracleIndex
            ns = min(2000, n_reads)
            orc = OracleIndex(index_dir)
            if is_count:
                wc, wp = orc.query_count(bases[: int(offsets[ns])], offsets[: ns + 1], threads=0)
                ok = np.array_equal(wc, h_cnt[:ns].numpy().view(np.uint64)) and np.array_equal(wp, h_pos[:ns].numpy())
            else:
                want = getattr(orc, "query_" + W["query"])(bases[: int(offsets[ns])], offsets[: ns + 1], threads=0)
                ok = np.array_equal(want, h_out[: int(offsets[ns])].numpy().view(np.uint16))
            line["parity_checked"] = bool(same and ok)
            line["cpu_baseline"] = {"value": v, "unit": "bases/s", "cores": cores, "kind": kind,
                                    "sample": "the full per-GPU batch (%d x %d bp), %.2f s of wall time on %d threads, index load excluded"
                                              % (n_reads, cfg["read_len"], sec, cores)}
        emit_line(line)
    ms.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
