"""Empty stand-in for matplotlib (display/export only; not on the solve path). Test infrastructure."""
