"""Render query results the way the reference prints them (for comparison with tests/golden/expected)."""


def ml_text(ids, offsets, ml):
    """`movi query --pml|--zml --stdout` (reference src/utils.cpp:204-209): values in read order, blank after each."""
    out = []
    for i, name in enumerate(ids):
        s, e = int(offsets[i]), int(offsets[i + 1])
        out.append(">" + name + "\n" + "".join("%d " % v for v in ml[s:e][::-1]) + "\n")
    return "".join(out)


def count_text(ids, offsets, counts, pos):
    """`movi query --count --stdout` (reference src/utils.cpp:248-256)."""
    out = []
    for i, name in enumerate(ids):
        L = int(offsets[i + 1]) - int(offsets[i])
        out.append("%s\t%d/%d\t%d\n" % (name, L - int(pos[i]), L, int(counts[i])))
    return "".join(out)
