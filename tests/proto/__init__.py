"""Test infrastructure: CPU prototypes of the encode kernels, fuzzed against the oracle (they import oracle/, so they live under tests/)."""
