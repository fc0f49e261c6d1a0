#!/usr/bin/env python3
"""Regenerate tests/golden/ from the UNMODIFIED reference (run in the build container only).

Needs /root/reference (read-only mount) and the binaries `make -C oracle ref` compiles from it
into oracle/_ref/.  Nothing here runs on the GPU box: the outputs are committed.

What it writes (all small, xz-compressed where bulky):
  c1/<type>[.sep]/index.movi.xz     indexes of the reference's tests_data/ref.fasta built by the
                                    reference's own `build` (sizes must equal tests/test_build.cpp:31-97)
  c1/regular-thresholds[.sep]/movi.{pml,zml}.nulldb.xz   frozen null databases (the reference seeds them with time(0))
  reads/*.fa|fq                     read sets: the reference's sample.fastq / sample.fasta, the six long
                                    reads recovered from reads.fasta.pmls.filtered.sorted, a seeded
                                    synthetic set and an edge-case set
  ref_goldens/*                     the reference's own golden files for these inputs (verbatim)
  expected/<set>[.sep].<q>.txt.xz   stdout of `movi-<type> query --<q> --no-prefetch -t1 --stdout`
                                    (identical for every layout; asserted here)
  expected/*.report / *.filtered    --classify / --filter outputs
  MANIFEST.json                     sizes and checks performed
"""
import json
import lzma
import os
import random
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("MOVI_REFERENCE", "/root/reference")
BIN = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")
WORK = os.environ.get("MOVI_GOLDEN_WORK", "/tmp/movi_golden_work")
TYPES = {"regular-thresholds": 6, "blocked-thresholds": 8, "sampled-thresholds": 7}
GOLDEN_SIZES = {  # reference tests/test_build.cpp:37,53,45 and :79,95,87
    ("regular-thresholds", False): 948119, ("blocked-thresholds", False): 711733, ("sampled-thresholds", False): 475326,
    ("regular-thresholds", True): 948232, ("blocked-thresholds", True): 711854, ("sampled-thresholds", True): 505009,
}


def sh(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, **kw).stdout


def xz_write(path, data):
    with lzma.open(path, "wb", preset=9 | lzma.PRESET_EXTREME) as f:
        f.write(data)


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def make_read_sets(ref_seq):
    os.makedirs(os.path.join(GOLD, "reads"), exist_ok=True)
    os.makedirs(os.path.join(GOLD, "ref_goldens"), exist_ok=True)
    td = os.path.join(REF, "tests_data")
    for f in ("sample.fastq", "sample.fasta"):
        shutil.copy(os.path.join(td, f), os.path.join(GOLD, "reads", f))
    for f in ("sample.fastq.pmls.sorted", "sample.fasta.pmls.filtered_notfound.sorted"):
        shutil.copy(os.path.join(td, f), os.path.join(GOLD, "ref_goldens", f))

    # six FOUND nanopore reads: sequences survive in the reference's filter golden; the sort decoupled
    # names from sequences, floor(len/150) == above+below of exactly one FOUND report line re-pairs them.
    lines = open(os.path.join(td, "reads.fasta.pmls.filtered.sorted")).read().split("\n")
    seqs = [l for l in lines if l and not l.startswith(">")]
    report = open(os.path.join(td, "reads.fasta.pmls.report.sorted")).read().split("\n")
    found = [l for l in report if " FOUND " in l]
    header = [l for l in report if l.startswith("read id:")][0]
    pairs = []
    for s in seqs:
        nb = len(s) // 150
        match = [l for l in found if int(l.split()[3]) + int(l.split()[4]) == nb]
        assert len(match) == 1, (len(s), match)
        pairs.append((match[0].split()[0], s, match[0]))
    pairs.sort()
    with open(os.path.join(GOLD, "reads", "long6.fasta"), "w") as f:
        for name, s, _ in pairs:
            f.write(">%s nanopore read recovered from the reference's filter golden\n%s\n" % (name, s))
    with open(os.path.join(GOLD, "ref_goldens", "long6.report.sorted"), "w") as f:
        for _, _, l in sorted(pairs, key=lambda p: p[2]):
            f.write(l + "\n")
        f.write(header + "\n")

    # seeded synthetic set (SURVEY 8d, C1 stand-in for the missing reads.fasta)
    rng = random.Random(1)
    both = ref_seq + revcomp(ref_seq)
    with open(os.path.join(GOLD, "reads", "synth300.fasta"), "w") as f:
        for i in range(300):
            L = rng.randint(150, 2000)
            if rng.random() < 0.10:
                s = "".join(rng.choice("ACGT") for _ in range(L))
            else:
                p = rng.randint(0, len(both) - L)
                s = list(both[p:p + L])
                for j in range(L):
                    if rng.random() < 0.01:
                        s[j] = rng.choice("ACGT")
                s = "".join(s)
            desc = " len=%d" % L if i % 3 == 0 else ("\tx" if i % 3 == 1 else "")
            f.write(">syn.%d%s\n" % (i, desc))
            for k in range(0, L, 70):       # multi-line FASTA
                f.write(s[k:k + 70] + "\n")

    # edge cases: illegal characters, lower case, tiny reads, long exact matches
    p = 1234
    exact = ref_seq[p:p + 2500]
    edge = [
        ("one_A", "A"), ("one_N", "N"), ("two", "GT"), ("acgt", "ACGT"), ("all_N", "NNNNNNNN"),
        ("trailing_N", ref_seq[100:180] + "N"), ("trailing_NN", ref_seq[100:180] + "NN"),
        ("leading_N", "N" + ref_seq[200:300]), ("leading_NNN", "NNN" + ref_seq[200:300]),
        ("middle_N", ref_seq[400:460] + "N" + ref_seq[461:520]),
        ("many_N", "".join(c if (i % 7) else "N" for i, c in enumerate(ref_seq[600:760]))),
        ("lower", ref_seq[800:900].lower()), ("mixed_case", ref_seq[900:950] + ref_seq[950:1000].lower()),
        ("percent", ref_seq[1000:1050] + "%" + ref_seq[1051:1100]), ("dollar", ref_seq[1100:1150] + "$" + ref_seq[1151:1200]),
        ("iupac", "ACGTRYKMSWACGTBDHVACGT"), ("exact2500", exact), ("exact_rc", revcomp(exact)),
        ("homopolymer_A", "A" * 400), ("homopolymer_T", "T" * 400), ("dinuc", "AC" * 200),
        ("N_then_exact", "N" + exact[:300]), ("exact_then_N", exact[:300] + "N"),
        ("second_last_N", ref_seq[1500:1600] + "N" + "G"), ("first_illegal_only", "N" + "C"),
    ]
    with open(os.path.join(GOLD, "reads", "edge.fasta"), "w") as f:
        for name, s in edge:
            f.write(">edge.%s\n%s\n" % (name, s))
    return ["sample.fastq", "sample.fasta", "long6.fasta", "synth300.fasta", "edge.fasta"]


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not mounted at " + REF)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "movi_b200", "datagen")])
    bwtgen = os.path.join(ROOT, "movi_b200", "datagen", "_build", "bwt_thresholds")
    shutil.rmtree(WORK, ignore_errors=True)
    os.makedirs(WORK)
    manifest = {"index_sizes": {}, "checks": []}

    ref_fasta = os.path.join(REF, "tests_data", "ref.fasta")
    ref_seq = "".join(l.strip() for l in open(ref_fasta) if not l.startswith(">")).upper()
    xz_write(os.path.join(GOLD, "ref.fasta.xz"), open(ref_fasta, "rb").read())
    sets = make_read_sets(ref_seq)

    os.makedirs(os.path.join(GOLD, "expected"), exist_ok=True)
    for sep in (False, True):
        tag = ".sep" if sep else ""
        outs = {}
        for typ in TYPES:
            d = os.path.join(WORK, typ + tag)
            os.makedirs(d)
            fa = os.path.join(d, "ref.fa")
            sh([os.path.join(BIN, "movi-prepare-ref"), ref_fasta, fa] + (["separators"] if sep else []))
            sh([bwtgen, fa, fa, "-q"])
            sh([os.path.join(BIN, "movi-" + typ), "build", "--index", d, "--fasta", fa, "--verify"] + (["--separators"] if sep else []))
            size = os.path.getsize(os.path.join(d, "index.movi"))
            assert size == GOLDEN_SIZES[(typ, sep)], (typ, sep, size)
            manifest["index_sizes"][typ + tag] = size
            gd = os.path.join(GOLD, "c1", typ + tag)
            os.makedirs(gd, exist_ok=True)
            xz_write(os.path.join(gd, "index.movi.xz"), open(os.path.join(d, "index.movi"), "rb").read())
            if TYPES[typ] == 6:   # the classifier only needs one frozen copy (it is layout independent)
                for q in ("pml", "zml"):
                    xz_write(os.path.join(gd, "movi.%s.nulldb.xz" % q), open(os.path.join(d, "movi.%s.nulldb" % q), "rb").read())
            for s in sets:
                rf = os.path.join(GOLD, "reads", s)
                qs = ["pml"] if TYPES[typ] == 7 else ["pml", "zml", "count"]
                for q in qs:
                    extra = ["--separators"] if sep else []
                    o = sh([os.path.join(BIN, "movi-" + typ), "query", "--index", d, "--read", rf, "--" + q,
                            "--no-prefetch", "-t1", "--stdout"] + extra)
                    outs.setdefault((s, q), {})[typ] = o
        # every layout must give identical text (reference tests/test_pml.cpp:57-115)
        for (s, q), per in outs.items():
            vals = list(per.values())
            assert all(v == vals[0] for v in vals), ("layouts disagree", s, q, sep)
            xz_write(os.path.join(GOLD, "expected", "%s%s.%s.txt.xz" % (s, tag, q)), vals[0])
        manifest["checks"].append("all layouts identical for %d (set, query) pairs, separators=%s" % (len(outs), sep))

    # the reference's own goldens, re-checked against what we just produced
    got = lzma.open(os.path.join(GOLD, "expected", "sample.fastq.pml.txt.xz")).read().decode().split("\n")
    want = open(os.path.join(GOLD, "ref_goldens", "sample.fastq.pmls.sorted")).read().split("\n")
    assert sorted(l for l in got if l) == sorted(l for l in want if l)
    manifest["checks"].append("sample.fastq.pmls.sorted reproduced")

    # classification / filter outputs (PML), regular-thresholds index, frozen nulldb
    d = os.path.join(WORK, "regular-thresholds")
    for s in ("long6.fasta", "sample.fasta", "synth300.fasta"):
        rf = os.path.join(GOLD, "reads", s)
        rep = sh([os.path.join(BIN, "movi-regular-thresholds"), "query", "--index", d, "--read", rf, "--pml",
                  "--classify", "--no-prefetch", "-t1", "--stdout", "--no-output"])
        # --no-output suppresses the report too; run without it but in a scratch copy of the reads
        scratch = os.path.join(WORK, s)
        shutil.copy(rf, scratch)
        rep = sh([os.path.join(BIN, "movi-regular-thresholds"), "query", "--index", d, "--read", scratch, "--pml",
                  "--classify", "--no-prefetch", "-t1", "--stdout"])
        open(os.path.join(GOLD, "expected", s + ".pml.report"), "wb").write(rep)
        bpf = scratch + ".regular-thresholds.pml.bpf"
        if s == "sample.fasta":
            xz_write(os.path.join(GOLD, "expected", s + ".pml.bpf.xz"), open(bpf, "rb").read())
        for inv in (False, True):
            flt = sh([os.path.join(BIN, "movi-regular-thresholds"), "query", "--index", d, "--read", scratch, "--pml",
                      "--filter", "--no-prefetch", "-t1"] + (["--invert"] if inv else []))
            xz_write(os.path.join(GOLD, "expected", s + (".pml.filtered_notfound.xz" if inv else ".pml.filtered.xz")), flt)
    rep = open(os.path.join(GOLD, "expected", "long6.fasta.pml.report")).read().split("\n")
    want = open(os.path.join(GOLD, "ref_goldens", "long6.report.sorted")).read().split("\n")
    norm = lambda ls: sorted(" ".join(l.split()) for l in ls if l)
    assert norm(rep) == norm(want), (norm(rep), norm(want))
    manifest["checks"].append("reads.fasta.pmls.report.sorted FOUND lines reproduced for the six recovered reads (thr=7)")
    flt = lzma.open(os.path.join(GOLD, "expected", "sample.fasta.pml.filtered_notfound.xz")).read().decode().split("\n")
    want = open(os.path.join(GOLD, "ref_goldens", "sample.fasta.pmls.filtered_notfound.sorted")).read().split("\n")
    assert sorted(l for l in flt if l) == sorted(l for l in want if l)
    manifest["checks"].append("sample.fasta.pmls.filtered_notfound.sorted reproduced")

    json.dump(manifest, open(os.path.join(GOLD, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
