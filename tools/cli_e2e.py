#!/usr/bin/env python3
"""File-to-file timing of `movi-b200 query` on the c2 workload (1 M x 150 bp FASTA) next to the reference binary on the host cores."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from movi_b200 import build


def timed(cmd, **kw):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)
    return time.perf_counter() - t0, r


def main():
    index_dir, bases, offsets, cfg = bench.prepare_workload(0)
    n_reads = len(offsets) - 1
    cli = build.CLI
    ref = os.path.join(ROOT, "oracle", "_ref", "movi-regular-thresholds")
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        fa = os.path.join(td, "reads.fasta")
        bench.write_fasta(fa, bases, offsets, n_reads)
        size = os.path.getsize(fa)
        out = []
        for name, args in (("pml -> .bpf", ["--pml"]), ("pml --no-output", ["--pml", "--no-output"]), ("pml --classify (report + .bpf)", ["--pml", "--classify"]),
                           ("pml --filter > /dev/null", ["--pml", "--filter"]), ("count -> .matches", ["--count"])):
            if "--classify" in args or "--filter" in args:
                if not os.path.exists(os.path.join(index_dir, "movi.pml.nulldb")):
                    continue
            best = None
            for rep in range(3):
                t, r = timed([cli, "query", "-i", index_dir, "-r", fa, "--verbose"] + args)
                if r.returncode != 0:
                    print("FAILED", name, r.stderr.decode()[-300:]); break
                best = t if best is None else min(best, t)
                info = [l for l in r.stderr.decode().split("\n") if "[INFO]" in l]
            if best is not None:
                out.append({"cmd": "movi-b200 query " + " ".join(args), "what": name, "wall_s": round(best, 3), "bases_per_s_wall": round(len(bases) / best), "info": info[-1] if info else ""})
                print(json.dumps(out[-1]), flush=True)
        if os.path.exists(ref):
            threads = os.cpu_count() or 1
            for name, args in (("reference pml -> .bpf", ["--pml"]), ("reference pml --no-output", ["--pml", "--no-output"])):
                t, r = timed([ref, "query", "--index", index_dir, "--read", fa, "-t", str(threads)] + args, cwd=td)
                out.append({"cmd": "reference query -t %d %s" % (threads, " ".join(args)), "what": name, "wall_s": round(t, 3), "bases_per_s_wall": round(len(bases) / t)})
                print(json.dumps(out[-1]), flush=True)
        print(json.dumps({"fasta_bytes": size, "n_reads": n_reads}))


if __name__ == "__main__":
    main()
