#!/usr/bin/env python3
"""Movi Color goldens from the UNMODIFIED reference -- run in the build container only.

A six-document reference (two clades of three close strains; documents 2 and 3 share a taxon id) is indexed by the
reference's own `build`, coloured by its `color` action (flat color table, the default) and sampled by `build-SA`;
the reference then answers `query --pml --multi-classify --no-prefetch -t1` (default, `--min-len 3`, `--pvalue-scoring`,
`--min-len 0`) and `query --pml --sa-entries --no-prefetch -t1`.  What is committed under tests/golden/color/:

  index.movi.xz doc_sets_flat.bin.xz doc_sets.bin.xz ssa.movi.xz ref.fa.doc_offsets ref.fa.doc_ids    the index directory
                                                                                       (doc_sets.bin: `color --color-vectors`)
  reads.fa                                                                             220 reads (10 % random, some with N, one empty-ish)
  expected.multi_classify[.minlen3|.pvalue|.minlen0|.vectors].csv                      the reference's -o files (.vectors: --color-vectors)
  expected.pml.bpf.xz expected.sa_entries.bpf.xz                                       the reference's binary outputs
"""
import lzma
import os
import random
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("MOVI_REFERENCE", "/root/reference")
BIN = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "color")
WORK = os.environ.get("MOVI_GOLDEN_WORK", "/tmp/movi_golden_color")
MOVI = os.path.join(BIN, "movi-regular-thresholds")


def sh(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, **kw).stdout


def xz(src, dst):
    with lzma.open(dst, "wb", preset=9 | lzma.PRESET_EXTREME) as f:
        f.write(open(src, "rb").read())


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not mounted at " + REF)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "movi_b200", "datagen")])
    bwtgen = os.path.join(ROOT, "movi_b200", "datagen", "_build", "bwt_thresholds")
    shutil.rmtree(WORK, ignore_errors=True)
    idx = os.path.join(WORK, "idx")
    os.makedirs(idx)
    os.makedirs(OUT, exist_ok=True)

    rng = random.Random(11)

    def mut(s, rate):
        s = list(s)
        for i in range(len(s)):
            if rng.random() < rate:
                s[i] = rng.choice("ACGT")
        return "".join(s)

    def rc(s):
        return s[::-1].translate(str.maketrans("ACGT", "TGCA"))

    base = "".join(rng.choice("ACGT") for _ in range(12000))
    clades = [mut(base, 0.06), mut(base, 0.06)]
    docs = [mut(clades[i // 3], 0.012) for i in range(6)]
    ref_fasta = os.path.join(WORK, "ref.fasta")
    with open(ref_fasta, "w") as f:
        for i, d in enumerate(docs):
            f.write(">doc%d\n" % i)
            for k in range(0, len(d), 80):
                f.write(d[k:k + 80] + "\n")
    reads = os.path.join(OUT, "reads.fa")
    with open(reads, "w") as f:
        for i in range(220):
            L = rng.randint(40, 1200)
            if i % 10 == 9:
                s = "".join(rng.choice("ACGT") for _ in range(L))
            else:
                d = docs[rng.randrange(6)]
                p = rng.randrange(0, len(d) - L)
                s = mut(d[p:p + L], 0.01)
                if rng.random() < 0.5:
                    s = rc(s)
            if i % 17 == 3:
                s = s[:L // 2] + "N" + s[L // 2 + 1:]
            if i == 5:
                s = "A"
            if i == 6:
                s = "NNNN"
            f.write(">r%d%s\n%s\n" % (i, " extra words" if i % 4 == 0 else "", s))

    fa = os.path.join(idx, "ref.fa")
    sh([os.path.join(BIN, "movi-prepare-ref"), ref_fasta, fa])
    open(os.path.join(idx, "ref.fa.doc_ids"), "w").write("".join("%d\n" % t for t in (562, 9606, 1280, 1280, 28901, 10090)))
    sh([bwtgen, fa, fa, "-q"])
    sh([MOVI, "build", "--index", idx, "--fasta", fa, "--verify"])
    sh([MOVI, "color", "--index", idx])
    sh([MOVI, "build-SA", "--index", idx])
    sh([MOVI, "color", "--index", idx, "--color-vectors"])                    # the same colours as vectors of vectors: doc_sets.bin
    for f in ("index.movi", "doc_sets_flat.bin", "doc_sets.bin", "ssa.movi"):
        xz(os.path.join(idx, f), os.path.join(OUT, f + ".xz"))
    for f in ("ref.fa.doc_offsets", "ref.fa.doc_ids"):
        shutil.copy(os.path.join(idx, f), os.path.join(OUT, f))

    scratch = os.path.join(WORK, "reads.fa")
    shutil.copy(reads, scratch)
    variants = {"": [], ".minlen3": ["--min-len", "3"], ".pvalue": ["--pvalue-scoring"], ".minlen0": ["--min-len", "0"],
                ".pvalue.minlen0": ["--pvalue-scoring", "--min-len", "0"]}
    for tag, extra in variants.items():
        o = os.path.join(WORK, "mc%s.csv" % tag)
        sh([MOVI, "query", "--index", idx, "--read", scratch, "--pml", "--multi-classify", "--no-prefetch", "-t1", "-o", o] + extra)
        shutil.copy(o, os.path.join(OUT, "expected.multi_classify%s.csv" % tag))
        if tag == "":
            xz(o + ".pml.bpf", os.path.join(OUT, "expected.pml.bpf.xz"))
    o = os.path.join(WORK, "mc.vectors.csv")
    sh([MOVI, "query", "--index", idx, "--read", scratch, "--pml", "--multi-classify", "--color-vectors", "--no-prefetch", "-t1", "-o", o])
    shutil.copy(o, os.path.join(OUT, "expected.multi_classify.vectors.csv"))
    sh([MOVI, "query", "--index", idx, "--read", scratch, "--pml", "--sa-entries", "--no-prefetch", "-t1"])
    xz(scratch + ".regular-thresholds.pml.sa_entries.bpf", os.path.join(OUT, "expected.sa_entries.bpf.xz"))
    assert open(scratch + ".regular-thresholds.pml.bpf", "rb").read() == lzma.open(os.path.join(OUT, "expected.pml.bpf.xz")).read()
    print("wrote", sorted(os.listdir(OUT)))
    print({f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})


if __name__ == "__main__":
    main()
