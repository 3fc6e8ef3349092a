"""shoulder_b200 — B200-native multiplane slicing / unrolling backend for the hot path of
gregspangenberg/shoulder (reference ``src/shoulder/humerus/slice.py``).

Only the slice provider is replaced; everything above it (landmarks, metrics, plotting) is the
reference's own Python and is not rebuilt here.  See DESIGN.md for the scope table.
"""
__version__ = "0.1.0"
