"""CPU oracle (test infrastructure only) -- see oracle/spectral.py."""
