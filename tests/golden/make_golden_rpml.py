#!/usr/bin/env python3
"""Supplement to make_golden.py: expected `--pml --rpml` output of the UNMODIFIED reference (oracle/_ref binaries) on the committed
index fixtures -> tests/golden/expected/<set>[.sep].rpml.txt.xz.  Runs in the build container only (needs oracle/_ref)."""
import json, lzma, os, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref")
TYPES = ("regular-thresholds", "blocked-thresholds", "sampled-thresholds")
SETS = ("sample.fastq", "sample.fasta", "long6.fasta", "synth300.fasta", "edge.fasta")


def main():
    n = 0
    with tempfile.TemporaryDirectory() as td:
        for sep in (False, True):
            tag = ".sep" if sep else ""
            outs = {}
            for typ in TYPES:
                d = os.path.join(td, typ + tag)
                os.makedirs(d)
                with open(os.path.join(d, "index.movi"), "wb") as f:
                    f.write(lzma.open(os.path.join(HERE, "c1", typ + tag, "index.movi.xz")).read())
                for s in SETS:
                    r = subprocess.run([os.path.join(BIN, "movi-" + typ), "query", "--index", d, "--read", os.path.join(HERE, "reads", s), "--pml", "--rpml",
                                        "--no-prefetch", "-t1", "--stdout"] + (["--separators"] if sep else []), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                    assert r.returncode == 0, r.stderr.decode()[-500:]
                    outs.setdefault(s, {})[typ] = r.stdout
            for s, per in outs.items():
                vals = list(per.values())
                assert all(v == vals[0] for v in vals), ("layouts disagree", s, sep)
                pml = lzma.open(os.path.join(HERE, "expected", "%s%s.pml.txt.xz" % (s, tag))).read()
                with lzma.open(os.path.join(HERE, "expected", "%s%s.rpml.txt.xz" % (s, tag)), "wb", preset=9) as f:
                    f.write(vals[0])
                n += 1
                print(s + tag, "rpml differs from pml:", vals[0] != pml, len(vals[0]))
    m = json.load(open(os.path.join(HERE, "MANIFEST.json")))
    m.setdefault("checks", []).append("--rpml: all layouts identical for %d (set, separators) pairs (make_golden_rpml.py)" % n)
    json.dump(m, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
