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


SIZES = {("regular", False): 871479, ("blocked", False): 654253, ("sampled", False): 437006,       # reference tests/test_build.cpp:33,49,41
         ("regular", True): 871496, ("blocked", True): 654280, ("sampled", True): 464203}          # --separators: :75,91,83


def main():
    for (typ, sep), size_want in SIZES.items():
        tag = ".sep" if sep else ""
        with tempfile.TemporaryDirectory() as td:
            ref = os.path.join(td, "ref.fasta")
            open(ref, "wb").write(lzma.open(os.path.join(HERE, "ref.fasta.xz")).read())
            fa = os.path.join(td, "ref.fa")
            sh([os.path.join(BIN, "movi-prepare-ref"), ref, fa] + (["separators"] if sep else []))
            sh([BWTGEN, fa, fa, "-q"])
            sh([os.path.join(BIN, "movi-" + typ), "build", "--index", td, "--fasta", fa, "--verify"] + (["--separators"] if sep else []))
            size = os.path.getsize(os.path.join(td, "index.movi"))
            assert size == size_want, (typ, sep, size)
            gd = os.path.join(HERE, "c1", typ + tag)
            os.makedirs(gd, exist_ok=True)
            with lzma.open(os.path.join(gd, "index.movi.xz"), "wb", preset=9) as f:
                f.write(open(os.path.join(td, "index.movi"), "rb").read())
            for s in SETS:
                for q in (("pml",) if typ == "sampled" else ("pml", "zml", "count")):
                    o = sh([os.path.join(BIN, "movi-" + typ), "query", "--index", td, "--read", os.path.join(HERE, "reads", s), "--" + q, "--no-prefetch", "-t1", "--stdout"]
                           + (["--separators"] if sep else []))
                    if q == "pml":
                        with lzma.open(os.path.join(HERE, "expected", "%s.%s%s.pml.txt.xz" % (s, typ, tag)), "wb", preset=9) as f:
                            f.write(o)
                    else:
                        assert o == lzma.open(os.path.join(HERE, "expected", "%s%s.%s.txt.xz" % (s, tag, q))).read(), (typ, sep, s, q)
                    print(typ + tag, s, q, len(o))
    m = json.load(open(os.path.join(HERE, "MANIFEST.json")))
    m.setdefault("index_sizes", {}).update({t + (".sep" if sp else ""): v for (t, sp), v in SIZES.items()})
    m.setdefault("checks", []).append("regular / blocked / sampled (MODE 3 / 2 / 5), with and without separators: index sizes = tests/test_build.cpp:33-95; PML outputs of 5 read sets stored, ZML / count equal the thresholds indexes' (make_golden_nothresholds.py)")
    json.dump(m, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
