#!/usr/bin/env python3
"""Supplement to make_golden.py for the non-threshold "regular" index type (MODE 3, src/utils.cpp:20-24): index built by the
UNMODIFIED reference builder (its size is the reference's own known answer, tests/test_build.cpp:33: 871479 bytes) and the
reference's PML / ZML / count output on the five read sets.  PML on such an index repositions with reposition_randomly
(src/move_structure_query.cpp:308-311).  Runs in the build container only (needs oracle/_ref and the datagen tools)."""
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


def main():
    with tempfile.TemporaryDirectory() as td:
        ref = os.path.join(td, "ref.fasta")
        open(ref, "wb").write(lzma.open(os.path.join(HERE, "ref.fasta.xz")).read())
        fa = os.path.join(td, "ref.fa")
        sh([os.path.join(BIN, "movi-prepare-ref"), ref, fa])
        sh([BWTGEN, fa, fa, "-q"])
        sh([os.path.join(BIN, "movi-regular"), "build", "--index", td, "--fasta", fa, "--verify"])
        size = os.path.getsize(os.path.join(td, "index.movi"))
        assert size == 871479, size            # reference tests/test_build.cpp:33
        gd = os.path.join(HERE, "c1", "regular")
        os.makedirs(gd, exist_ok=True)
        with lzma.open(os.path.join(gd, "index.movi.xz"), "wb", preset=9) as f:
            f.write(open(os.path.join(td, "index.movi"), "rb").read())
        for s in SETS:
            for q in ("pml", "zml", "count"):
                o = sh([os.path.join(BIN, "movi-regular"), "query", "--index", td, "--read", os.path.join(HERE, "reads", s), "--" + q, "--no-prefetch", "-t1", "--stdout"])
                with lzma.open(os.path.join(HERE, "expected", "%s.regular.%s.txt.xz" % (s, q)), "wb", preset=9) as f:
                    f.write(o)
                same = None
                if q != "pml":
                    same = o == lzma.open(os.path.join(HERE, "expected", "%s.%s.txt.xz" % (s, q))).read()
                print(s, q, len(o), "same as the thresholds index:" if same is not None else "", same if same is not None else "")
    m = json.load(open(os.path.join(HERE, "MANIFEST.json")))
    m.setdefault("index_sizes", {})["regular"] = 871479
    m.setdefault("checks", []).append("regular (MODE 3): index size 871479 = tests/test_build.cpp:33; pml / zml / count outputs of 5 read sets (make_golden_regular.py)")
    json.dump(m, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
