"""Synthetic workloads for tests/, bench.py, smoke() and tools/ (scaffolding, not part of the product package):
`synth` builds seeded images and raw head maps with planted faces, `standin` is the stand-in for the (unchanged,
out-of-scope) backbone that delivers them to the detector."""
