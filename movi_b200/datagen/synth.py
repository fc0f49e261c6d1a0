"""Seeded synthetic inputs for the BASELINE.json configurations.

C2 / C3 / C5 data: base genome of `genome_len` uniform ACGT (seed 42); `n_hap` haplotypes, each the base with
per-position SNP 0.08 %, 1-bp deletion 0.01 %, 1-bp insertion 0.01 %; text = every haplotype followed by its
reverse complement (what movi-prepare-ref writes, reference src/prepare_ref.cpp:61-71).
Reads: substrings of a random haplotype / strand with 0.5 % substitutions, 2 % of reads fully random, ACGT only.

Everything is regenerated from the seeds; indexes are cached under _cache/ (git-ignored) because the suffix
array of a 6.4e8-character text takes minutes.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(HERE, "_build")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[ACGT] = np.frombuffer(b"TGCA", dtype=np.uint8)


def tools():
    subprocess.check_call(["make", "-s", "-C", HERE])
    return os.path.join(BUILD, "bwt_thresholds"), os.path.join(BUILD, "build_index")


def revcomp(a):
    return _COMP[a[::-1]]


def base_genome(genome_len, seed=42):
    return ACGT[np.random.default_rng(seed).integers(0, 4, genome_len)]


def haplotype(base, h, seed=42, snp=0.0008, dele=0.0001, ins=0.0001):
    """Haplotype h of the pangenome (deterministic in (seed, h))."""
    rng = np.random.default_rng([seed, 1000 + h])
    L = len(base)
    hap = base.copy()
    s = np.nonzero(rng.random(L) < snp)[0]
    hap[s] = ACGT[(np.searchsorted(ACGT, hap[s]) + rng.integers(1, 4, len(s))) % 4]     # always a different base
    keep = rng.random(L) >= dele
    insert = rng.random(L) < ins
    reps = keep.astype(np.int64) + insert.astype(np.int64)
    out = np.repeat(hap, reps)
    # inserted copies: the second element of every (keep & insert) pair / the only one of (!keep & insert) -> random base
    ends = np.cumsum(reps)
    ins_pos = ends[insert] - 1
    out[ins_pos] = ACGT[rng.integers(0, 4, len(ins_pos))]
    return out


def pangenome(genome_len, n_hap, seed=42):
    base = base_genome(genome_len, seed)
    return [haplotype(base, h, seed) for h in range(n_hap)]


def make_reads(haps, n_reads, read_len, seed=43, sub=0.005, frac_random=0.02, len_jitter=0):
    """(bases uint8[total], offsets uint64[n+1]); ACGT only."""
    rng = np.random.default_rng(seed)
    if len_jitter:
        lens = rng.integers(max(1, read_len - len_jitter), read_len + len_jitter + 1, n_reads)
    else:
        lens = np.full(n_reads, read_len, dtype=np.int64)
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    bases = np.empty(int(offsets[-1]), dtype=np.uint8)
    which = rng.integers(0, len(haps), n_reads)
    strand = rng.integers(0, 2, n_reads)
    is_random = rng.random(n_reads) < frac_random
    u = rng.random(n_reads)
    for h in range(len(haps)):
        hap = haps[h]
        sel = np.nonzero((which == h) & ~is_random)[0]
        if len(sel) == 0:
            continue
        if len_jitter == 0:
            L = read_len
            starts = (u[sel] * (len(hap) - L)).astype(np.int64)
            block = hap[starts[:, None] + np.arange(L)[None, :]]
            rc = strand[sel] == 1
            block[rc] = _COMP[block[rc][:, ::-1]]
            dst = (offsets[sel].astype(np.int64)[:, None] + np.arange(L)[None, :])
            bases[dst] = block
        else:
            for i in sel:
                L = int(lens[i])
                st = int(u[i] * (len(hap) - L))
                seg = hap[st:st + L]
                bases[int(offsets[i]):int(offsets[i + 1])] = revcomp(seg) if strand[i] else seg
    for i in np.nonzero(is_random)[0]:
        bases[int(offsets[i]):int(offsets[i + 1])] = ACGT[rng.integers(0, 4, int(lens[i]))]
    err = np.nonzero(rng.random(len(bases)) < sub)[0]
    bases[err] = ACGT[(np.searchsorted(ACGT, bases[err]) + rng.integers(1, 4, len(err))) % 4]
    return bases, offsets


def write_text(haps, path):
    """The BWT input: hap0 rc(hap0) hap1 rc(hap1) ... as one raw byte string (no separators)."""
    with open(path, "wb") as f:
        for hap in haps:
            f.write(hap.tobytes())
            f.write(revcomp(hap).tobytes())


def build_indexes(haps, out_dir, types=("regular-thresholds",), checkpoint=20, keep_bwt=False, log=sys.stderr):
    """Text -> BWT + thresholds (run-length form) -> index.movi per requested type under out_dir/<type>/."""
    bwtgen, builder = tools()
    os.makedirs(out_dir, exist_ok=True)
    prefix = os.path.join(out_dir, "ref")
    need = [t for t in types if not os.path.exists(os.path.join(out_dir, t, "index.movi"))]
    if not need:
        return
    if not os.path.exists(prefix + ".thr_pos"):
        t0 = time.time()
        write_text(haps, prefix + ".txt")
        subprocess.check_call([bwtgen, "--raw", prefix + ".txt", prefix, "--rle"], stderr=log)
        os.remove(prefix + ".txt")
        print("[datagen] BWT + thresholds in %.0f s" % (time.time() - t0), file=log)
    for t in need:
        t0 = time.time()
        subprocess.check_call([builder, prefix, os.path.join(out_dir, t), t, "--rle", "--checkpoint", str(checkpoint)], stderr=log)
        print("[datagen] %s index in %.0f s" % (t, time.time() - t0), file=log)
    if not keep_bwt:
        for ext in (".bwt.heads", ".bwt.len", ".thr_pos"):
            if os.path.exists(prefix + ext):
                os.remove(prefix + ext)


CONFIGS = {
    # name: genome_len, n_hap, n_reads, read_len
    "c2": dict(genome_len=5_000_000, n_hap=64, n_reads=1_000_000, read_len=150),
    "c2-small": dict(genome_len=100_000, n_hap=16, n_reads=20_000, read_len=150),   # CPU-sized twin used by tests
    # a suffix-sorted index BEYOND the L2: 20 haplotypes x 25 Mbp (+ reverse complements) = 1 Gbp of text, real LCP thresholds
    "g1": dict(genome_len=25_000_000, n_hap=20, n_reads=1_000_000, read_len=150),
}


def cache_dir(name):
    return os.environ.get("MOVI_CACHE", os.path.join(ROOT, "_cache", name))


def _unpack(d, t):
    """The cache ships xz-compressed (48 MB instead of 124 MB for c2): unpack on first use."""
    f = os.path.join(d, t, "index.movi")
    if not os.path.exists(f) and os.path.exists(f + ".xz"):
        tmp = f + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as out:
            subprocess.check_call(["xz", "-dc", "-T0", f + ".xz"], stdout=out)
        os.replace(tmp, f)


def _pack(d, t):
    f = os.path.join(d, t, "index.movi")
    if os.path.exists(f) and not os.path.exists(f + ".xz") and os.path.getsize(f) > (8 << 20):
        with open(f + ".xz", "wb") as out:
            subprocess.check_call(["xz", "-1", "-T0", "-c", f], stdout=out)


def prepare(name, types=("regular-thresholds",), want_reads=True, log=sys.stderr):
    """Make sure the indexes of configuration `name` exist; return (dir, haps or None)."""
    cfg = CONFIGS[name]
    d = cache_dir(name)
    for t in types:
        _unpack(d, t)
    missing = [t for t in types if not os.path.exists(os.path.join(d, t, "index.movi"))]
    haps = None
    if missing or want_reads:
        haps = pangenome(cfg["genome_len"], cfg["n_hap"])
    if missing:
        build_indexes(haps, d, types, log=log)
        json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
    for t in types:
        _pack(d, t)
    return d, haps


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "c2-small"
    types = sys.argv[2].split(",") if len(sys.argv) > 2 else ["regular-thresholds"]
    t0 = time.time()
    d, _ = prepare(name, types, want_reads=False)
    print("prepared", d, "in %.0f s" % (time.time() - t0))
