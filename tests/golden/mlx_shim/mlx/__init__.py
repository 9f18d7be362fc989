"""NumPy stand-in for the `mlx` package (golden-vector generation only).

MLX is not installable in the build container.  This shim implements just the
`mlx.core` surface the reference's *Python* hot path touches, with float32 /
complex64 NumPy arithmetic, so that the reference's own unmodified Python code
(padding, framing, trims, dB order of operations, Griffin-Lim update ...) can
be executed to produce tests/golden/*.npz.  It is never imported by the
product package or at GPU run time.
"""
