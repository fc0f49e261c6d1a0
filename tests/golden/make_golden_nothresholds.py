#!/usr/bin/env python3
"""Supplement to make_golden.py for the non-threshold index types "regular", "blocked" and "sampled" (MODE 3 / 2 / 5,
src/utils.cpp:20-36): indexes built by the UNMODIFIED reference builder (their sizes are the reference's own known answers,
tests/test_build.cpp:33,41,49) and the reference's PML output on the five read sets (ZML / count, where the type has them, are
checked to equal the thresholds indexes' output and not stored again).  PML on such an index repositions with
reposition_randomly (src/move_structure_query.cpp:308-311), and every type splits its runs differently, so each has its own
expected PML.  Runs in the build container only (needs oracle/_ref and the datagen tools)."""
import json, lzma, os, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref")
BWTGEN = os.path.join(ROOT, "movi_b200", "datagen", "_build", "bwt_thresholds")
SETS = ("sample.fastq", "sample.fasta", "long6.fasta", "synth300.fasta", "edge.fasta")


def sh(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)
    assert r.returncode == 0, (cmd, r.stderr.decode()[-800:])
    return r.stdout


SIZES = {"regular": 871479, "blocked": 654253, "sampled": 437006}      # reference tests/test_build.cpp:33,49,41


def main():
    for typ, size_want in SIZES.items():
        with tempfile.TemporaryDirectory() as td:
            ref = os.path.join(td, "ref.fasta")
            open(ref, "wb").write(lzma.open(os.path.join(HERE, "ref.fasta.xz")).read())
            fa = os.path.join(td, "ref.fa")
            sh([os.path.join(BIN, "movi-prepare-ref"), ref, fa])
            sh([BWTGEN, fa, fa, "-q"])
            sh([os.path.join(BIN, "movi-" + typ), "build", "--index", td, "--fasta", fa, "--verify"])
            size = os.path.getsize(os.path.join(td, "index.movi"))
            assert size == size_want, (typ, size)
            gd = os.path.join(HERE, "c1", typ)
            os.makedirs(gd, exist_ok=True)
            with lzma.open(os.path.join(gd, "index.movi.xz"), "wb", preset=9) as f:
                f.write(open(os.path.join(td, "index.movi"), "rb").read())
            for s in SETS:
                for q in (("pml",) if typ == "sampled" else ("pml", "zml", "count")):
                    o = sh([os.path.join(BIN, "movi-" + typ), "query", "--index", td, "--read", os.path.join(HERE, "reads", s), "--" + q, "--no-prefetch", "-t1", "--stdout"])
                    if q == "pml":
                        with lzma.open(os.path.join(HERE, "expected", "%s.%s.pml.txt.xz" % (s, typ)), "wb", preset=9) as f:
                            f.write(o)
                    else:
                        assert o == lzma.open(os.path.join(HERE, "expected", "%s.%s.txt.xz" % (s, q))).read(), (typ, s, q)
                    print(typ, s, q, len(o))
    m = json.load(open(os.path.join(HERE, "MANIFEST.json")))
    m.setdefault("index_sizes", {}).update(SIZES)
    m.setdefault("checks", []).append("regular / blocked / sampled (MODE 3 / 2 / 5): index sizes = tests/test_build.cpp:33,49,41; PML outputs of 5 read sets stored, ZML / count equal the thresholds indexes' (make_golden_nothresholds.py)")
    json.dump(m, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
