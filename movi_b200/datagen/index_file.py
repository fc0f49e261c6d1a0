"""Writer of a regular-thresholds index.movi (MODE 6) from a table held elsewhere (bench / test data, not product).

The byte layout is the reference's serialize / deserialize pair (src/move_structure_io.cpp:66-109 header,
:145-184 basic data, :361-397 main table, :219-254 overflow tables, :269-287 counts and the backward-search seeds);
SURVEY Appendix A.2.  `rows_chunks` yields the 8-byte rows in the reference's MoveRow bit layout, in order."""
import os
import struct

import numpy as np

MOVI_MAGIC = 0x4D4F5649


def write_regular_thresholds(directory, n, r, end_bwt_idx, end_thresholds, rows_chunks, last_runs, last_offsets, first_runs,
                             first_offsets, counts=(0, 0, 0, 0), original_r=None, alphabet=b"ACGT"):
    os.makedirs(directory, exist_ok=True)
    sigma = len(alphabet)
    alphamap = np.full(256, 256, dtype=np.uint64)
    for k, ch in enumerate(alphabet):
        alphamap[ch] = k
    path = os.path.join(directory, "index.movi")
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        h = bytearray(48)
        struct.pack_into("<IBBBb", h, 0, MOVI_MAGIC, 2, 0, 0, 6)
        struct.pack_into("<QQQQ", h, 16, n, r, r if original_r is None else original_r, end_bwt_idx)
        f.write(h)
        f.write(struct.pack("<4Q", *end_thresholds))
        f.write(b"\0" * 64)
        f.write(struct.pack("<Q", 256))
        f.write(alphamap.tobytes())
        f.write(struct.pack("<Q", sigma))
        f.write(bytes(alphabet))
        f.write(struct.pack("<HB", 0, 0))
        written = 0
        for chunk in rows_chunks:
            a = np.ascontiguousarray(chunk)
            f.write(memoryview(a).cast("B"))
            written += a.nbytes
        if written != 8 * r:
            raise ValueError("row table: %d bytes written, %d expected" % (written, 8 * r))
        f.write(struct.pack("<QQQ", 0, 0, 0))
        f.write(struct.pack("<Q", sigma) + struct.pack("<%dQ" % sigma, *counts[:sigma]))
        f.write(struct.pack("<Q", sigma + 1))
        for arr in (last_runs, last_offsets, first_runs, first_offsets):
            f.write(struct.pack("<%dQ" % (sigma + 1), *list(arr)[: sigma + 1]))
    os.replace(tmp, path)
    return directory
