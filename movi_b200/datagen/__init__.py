"""Synthetic bench / test data: pangenomes, reads, BWT + thresholds, index files (not on the query path)."""
