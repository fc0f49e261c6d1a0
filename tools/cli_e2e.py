#!/usr/bin/env python3
"""File-to-file timing of `movi-b200 query` next to the reference binary on the host cores (same box, outputs on tmpfs).

usage: cli_e2e.py [n_reads_millions=10] [long]
  default: the c2 index, n x 150 bp reads in one FASTA file (10 M reads = 1.6 GB in, 3.2 GB of .bpf out);
  long   : the c2 index, 400 k x 10 kbp reads (configs[3] / [4] shape): the long-read occupancy check.
Writes gpurun_out/r02_cli_e2e.jsonl."""
import json
import os
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
    open(os.path.join(OUT, "r02_cli_e2e.jsonl"), "a").write(json.dumps(rec) + "\n")


def timed(cmd, **kw):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, **kw)
    return time.perf_counter() - t0, r


def main():
    import bench
    from movi_b200 import build
    from movi_b200.datagen import synth
    millions = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    long_reads = "long" in sys.argv[2:]
    gpus = [a for a in sys.argv[2:] if a.startswith("gpus=")]
    gpus = int(gpus[0][5:]) if gpus else 1
    d, haps = synth.prepare("c2", ("regular-thresholds",))
    index_dir = os.path.join(d, "regular-thresholds")
    n_reads, L = (int(millions * 1e6), 150) if not long_reads else (int(millions * 1e6), 10_000)
    cli = build.CLI
    ref = os.path.join(ROOT, "oracle", "_ref", "movi-regular-thresholds")
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        fa = os.path.join(td, "reads.fasta")
        t0 = time.time()
        per = min(n_reads, 1_000_000 if not long_reads else 50_000)
        with open(fa, "wb") as f:
            pass
        done = 0
        while done < n_reads:                       # the file is written a million reads at a time (bounded host memory)
            m = min(per, n_reads - done)
            bases, offsets = synth.make_reads(haps, m, L, seed=43 + done, **({"sub": 0.05, "frac_random": 0.0} if long_reads else {}))
            tmp = fa + ".part"
            bench.write_fasta(tmp, bases, offsets, m)
            with open(fa, "ab") as f, open(tmp, "rb") as g:
                while True:
                    buf = g.read(1 << 26)
                    if not buf:
                        break
                    f.write(buf)
            os.remove(tmp)
            done += m
        rep = [a for a in sys.argv[2:] if a.startswith("x")]
        if rep:                                    # "x10": the file concatenated 10 times (a big input without the generator's time)
            k = int(rep[0][1:])
            one = fa + ".one"
            os.rename(fa, one)
            with open(fa, "wb") as f:
                for _ in range(k):
                    with open(one, "rb") as g:
                        while True:
                            buf = g.read(1 << 26)
                            if not buf:
                                break
                            f.write(buf)
            os.remove(one)
            n_reads *= k
        n_bases = n_reads * L
        size = os.path.getsize(fa)
        emit({"what": "input", "reads": n_reads, "read_len": L, "fasta_bytes": size, "made_in_s": round(time.time() - t0, 1), "gpus": gpus})
        threads = os.cpu_count() or 1
        bpf = fa + ".regular-thresholds.pml.bpf"
        sums = {}
        runs = [("movi-b200 --pml -> .bpf (device-side ingest + .bpf emission)", [cli, "query", "-i", index_dir, "-r", fa, "--pml", "--verbose", "--gpus", str(gpus)], bpf),
                ("movi-b200 --pml --no-output (device-side ingest)", [cli, "query", "-i", index_dir, "-r", fa, "--pml", "--no-output", "--verbose", "--gpus", str(gpus)], None),
                ("movi-b200 --pml -> .bpf --host-parse (host parser, device-side .bpf emission)", [cli, "query", "-i", index_dir, "-r", fa, "--pml", "--host-parse", "--verbose", "--gpus", str(gpus)], bpf),
                ("movi-b200 --pml --filter > /dev/null (device-side ingest, classification on the device)", [cli, "query", "-i", index_dir, "-r", fa, "--pml", "--filter", "--verbose", "--gpus", str(gpus)], None)]
        if "classify" in sys.argv[2:]:             # the classification flows only: device-side ingest against the host parser (and the reference)
            rpt = fa + ".regular-thresholds.pml.report"
            runs = [runs[-1],
                    ("movi-b200 --pml --filter --host-parse > /dev/null (host parser, fused classify)", [cli, "query", "-i", index_dir, "-r", fa, "--pml", "--filter", "--host-parse", "--verbose", "--gpus", str(gpus)], None),
                    ("movi-b200 --pml --classify -> report + .bpf (device-side ingest)", [cli, "query", "-i", index_dir, "-r", fa, "--pml", "--classify", "--verbose", "--gpus", str(gpus)], bpf),
                    ("movi-b200 --pml --classify --host-parse -> report + .bpf (host parser)", [cli, "query", "-i", index_dir, "-r", fa, "--pml", "--classify", "--host-parse", "--verbose", "--gpus", str(gpus)], bpf)]
            if os.path.exists(ref) and "noref" not in sys.argv[2:]:
                runs += [("reference --pml --filter -t %d > /dev/null" % threads, [ref, "query", "--index", index_dir, "--read", fa, "--pml", "--filter", "-t", str(threads)], None)]
        if os.path.exists(ref) and "noref" not in sys.argv[2:] and "classify" not in sys.argv[2:]:
            runs += [("reference --pml -t %d -> .bpf" % threads, [ref, "query", "--index", index_dir, "--read", fa, "--pml", "-t", str(threads)], bpf),
                     ("reference --pml -t %d --no-output" % threads, [ref, "query", "--index", index_dir, "--read", fa, "--pml", "--no-output", "-t", str(threads)], None)]
        if rep:                                    # the big input: the device-ingest runs and the reference only
            runs = [r_ for r_ in runs if "host" not in r_[0]]
        for name, cmd, outfile in runs:
            if "--filter" in cmd and not os.path.exists(os.path.join(index_dir, "movi.pml.nulldb")):
                import struct
                with open(os.path.join(index_dir, "movi.pml.nulldb"), "wb") as f:
                    f.write(struct.pack("<QdQ", 1, 1.0, 6) + struct.pack("<Q", 1))
            best, info, startup = None, "", []
            for rep_i in range((1 if rep else 2) if "reference" in name else (2 if rep else 3)):
                if outfile and os.path.exists(outfile):
                    os.remove(outfile)
                t, r = timed(cmd, cwd=td)
                if r.returncode != 0:
                    emit({"what": name, "failed": r.stderr.decode()[-400:]})
                    best = None
                    break
                best = t if best is None else min(best, t)
                lines = [l for l in r.stderr.decode().split("\n") if "[INFO]" in l or "processing the reads" in l]
                info = lines[-1] if lines else ""
                startup = [l for l in lines if "start-up" in l]
            if best is None:
                continue
            rec = {"what": name, "wall_s": round(best, 3), "bases_per_s_wall": round(n_bases / best), "info": info.strip()}
            if "reference" not in name and startup:
                rec["startup"] = startup[-1].strip()
            m = __import__("re").search(r"written in ([0-9.]+) s", info)
            if m:
                rec["steady_state_s"] = float(m.group(1)); rec["bases_per_s_steady"] = round(n_bases / float(m.group(1)))
            if outfile and os.path.exists(outfile):
                rec["out_bytes"] = os.path.getsize(outfile)
                if "reference" not in name:              # (the reference writes its records in completion order)
                    import hashlib
                    h = hashlib.md5()
                    with open(outfile, "rb") as f:
                        f.seek(12)
                        while True:
                            buf = f.read(1 << 26)
                            if not buf:
                                break
                            h.update(buf)
                    rec["md5_after_header"] = h.hexdigest(); sums[name] = rec["md5_after_header"]
            emit(rec)
        if len(sums) > 1:
            emit({"what": "device-side ingest and host parser write the same .bpf", "same": len(set(sums.values())) == 1})


if __name__ == "__main__":
    main()
