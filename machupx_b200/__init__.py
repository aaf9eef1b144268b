"""machupx_b200 — B200-native lifting-line solve path with the MachUpX Scene API."""
