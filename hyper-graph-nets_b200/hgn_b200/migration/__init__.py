"""B200 mirror of the reference's ``src/migration`` package (same module and class names)."""
