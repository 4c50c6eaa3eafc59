"""Serving-side contracts of the path: the reference's Triton ensemble I/O, served by this engine."""
