"""Classification of a file chunk on the device (movi_gpu_file_chunk_classify / _fetch / _filter) through the C ABI, against the
host parser + movi_gpu_query_classify + the host writers (which restate the reference's report / filter output)."""
import os
import threading

import numpy as np
import pytest

from conftest import GOLD, reads_path
from movi_b200 import MoveStructure, Reads
from movi_b200 import _native as N

pytestmark = pytest.mark.gpu


def _host_text(rd, mode, thr, above, below, total):
    H = N.host_lib()
    rfd, wfd = os.pipe()
    buf = []
    t = threading.Thread(target=lambda: buf.append(os.fdopen(rfd, "rb").read()))
    t.start()
    try:
        assert H.movi_host_classify_bins(wfd, mode, rd._h, thr, above.ctypes.data, below.ctypes.data, total.ctypes.data, None) == 0
    finally:
        os.close(wfd)
        t.join()
    return buf[0]


@pytest.mark.parametrize("read_set,fastq", [("synth300.fasta", False), ("sample.fastq", True), ("edge.fasta", False), ("long6.fasta", False)])
def test_chunk_classify_fetch_and_filter(index_dirs, read_set, fastq):
    ms = MoveStructure(index_dirs[("regular-thresholds", False)], device=0)
    try:
        raw = open(reads_path(read_set), "rb").read()
        rd = Reads(reads_path(read_set))
        for thr, bw in ((7, 150), (3, 40)):
            a0, b0, s0 = ms.query_classify(rd.bases, rd.offsets, bw, thr)
            a, b, s, res, (bases, off, ids, id_off) = ms.classify_file_chunk(raw, fastq, bw, thr, want_reads=True)
            assert res.parse_flags == 0 and res.n_reads == rd.count
            assert np.array_equal(a, a0) and np.array_equal(b, b0) and np.array_equal(s, s0)
            hid, hoff = rd.id_blob
            assert np.array_equal(bases, rd.bases) and np.array_equal(off, rd.offsets) and np.array_equal(ids, hid) and np.array_equal(id_off, hoff)
            for invert in (False, True):
                text, _ = ms.filter_file_chunk(raw, fastq, bw, thr, invert=invert)
                assert text == _host_text(rd, 2 if invert else 1, thr, a0, b0, s0), (read_set, thr, invert)
            zt, _ = ms.filter_file_chunk(raw, fastq, bw, thr, zml=True)
            za, zb, zs = ms.query_classify(rd.bases, rd.offsets, bw, thr, zml=True)
            assert zt == _host_text(rd, 1, thr, za, zb, zs)
    finally:
        ms.close()
