#!/usr/bin/env python3
"""BASELINE.json configs[4]: sampled-thresholds index (id checkpoints every 20 rows; the c2 pangenome), 100 k x 10 kbp reads,
PML + --filter through the command line -- timed, next to the reference, and checked against the oracle's classification.
Also: PML kernel time of the device layouts of the c2 index (expanded 8-byte rows / 6-byte blocked / 4-byte sampled).
Writes gpurun_out/r02_c5.jsonl."""
import json
import os
import struct
import subprocess
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
    open(os.path.join(OUT, "r02_c5.jsonl"), "a").write(json.dumps(rec) + "\n")


def main():
    import torch
    import bench
    from movi_b200 import MoveStructure, build
    from movi_b200.api import LAYOUT_COMPACT, LAYOUT_EXPANDED8
    from movi_b200.datagen import synth
    from oracle.port import OracleIndex, classify
    d, haps = synth.prepare("c2", ("regular-thresholds", "blocked-thresholds", "sampled-thresholds"))
    dev = torch.device("cuda:0")
    # ---- device layouts on the c2 read set (1 M x 150 bp)
    bases, offsets = synth.make_reads(haps, 1_000_000, 150, seed=43)
    d_b = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).to(dev)
    d_o = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    d_out = torch.zeros(len(bases) + 64, dtype=torch.int16, device=dev)
    first = None
    for typ, layout, name in (("regular-thresholds", LAYOUT_EXPANDED8, "R8 (8 B/row)"), ("blocked-thresholds", LAYOUT_COMPACT, "B6 (6 B/row + id blocks)"),
                              ("sampled-thresholds", LAYOUT_EXPANDED8, "sampled expanded to R8"), ("sampled-thresholds", LAYOUT_COMPACT, "T4 (4 B/row + checkpoints)")):
        ms = MoveStructure(os.path.join(d, typ), layout=layout)
        st = torch.cuda.current_stream().cuda_stream
        fn = lambda: ms.query_pml_device(d_b.data_ptr(), d_o.data_ptr(), len(offsets) - 1, d_out.data_ptr(), st)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize(); ms.check_device_status()
        t = e0.elapsed_time(e1) / 10
        res = d_out[: len(bases)].clone()
        same = True if first is None else bool(torch.equal(first, res))
        first = res if first is None else first
        emit({"what": "c2 PML, device layout " + name, "index": typ, "ms": round(t, 3), "Gbases_per_s": round(len(bases) / t / 1e6, 2),
              "index_device_MB": round(ms.index_device_bytes / 1e6, 1), "same_as_first": same})
        ms.close()
    del d_b, d_out
    # ---- configs[4] through the command line
    n_reads, L = 100_000, 10_000
    sd = os.path.join(d, "sampled-thresholds")
    if not os.path.exists(os.path.join(sd, "movi.pml.nulldb")):
        with open(os.path.join(sd, "movi.pml.nulldb"), "wb") as f:
            f.write(struct.pack("<QdQ", 1, 1.0, 6) + struct.pack("<Q", 1))
    t0 = time.time()
    bases, offsets = synth.make_reads(haps, n_reads, L, seed=45, sub=0.05, frac_random=0.10)
    cli, ref = build.CLI, os.path.join(ROOT, "oracle", "_ref", "movi-sampled-thresholds")
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        fa = os.path.join(td, "reads.fasta")
        bench.write_fasta(fa, bases, offsets, n_reads)
        emit({"what": "c5 input", "reads": n_reads, "read_len": L, "made_in_s": round(time.time() - t0, 1)})
        # expected FOUND set of the first 3000 reads from the oracle
        ns = 3000
        orc = OracleIndex(sd)
        want = orc.query_pml(bases[: ns * L], offsets[: ns + 1], threads=0)
        exp = ["r%07d" % i for i in range(ns) if classify(want[i * L:(i + 1) * L], 150, 7)[0]]
        for name, extra in (("expanded", ["--layout", "expanded"]), ("compact T4", ["--layout", "compact"])):
            best, got = None, None
            for rep in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([cli, "query", "-i", sd, "-r", fa, "--pml", "--filter", "--verbose"] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                t = time.perf_counter() - t0
                if r.returncode:
                    emit({"what": "movi-b200 c5 " + name, "failed": r.stderr.decode()[-300:]}); break
                best = t if best is None else min(best, t)
                ids = [l[1:].decode() for l in r.stdout.split(b"\n") if l.startswith(b">")]
                got = [i for i in ids if int(i[1:]) < ns]
                info = [l for l in r.stderr.decode().split("\n") if "[INFO]" in l]
            if best is not None:
                emit({"what": "movi-b200 query --pml --filter, sampled-thresholds index, " + name, "wall_s": round(best, 3), "bases_per_s_wall": round(n_reads * L / best),
                      "found_total": len(ids), "first_%d_reads_equal_oracle" % ns: got == exp, "found_in_first": len(exp), "info": info[-1].strip() if info else ""})
        if os.path.exists(ref):
            t0 = time.perf_counter()
            r = subprocess.run([ref, "query", "--index", sd, "--read", fa, "--pml", "--filter", "-t", str(os.cpu_count() or 1)], stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, cwd=td)
            t = time.perf_counter() - t0
            n_found = sum(1 for l in r.stdout.split(b"\n") if l.startswith(b">"))
            emit({"what": "reference query --pml --filter -t %d, sampled-thresholds" % (os.cpu_count() or 1), "wall_s": round(t, 3), "bases_per_s_wall": round(n_reads * L / t),
                  "found_total": n_found, "rc": r.returncode})


if __name__ == "__main__":
    main()
