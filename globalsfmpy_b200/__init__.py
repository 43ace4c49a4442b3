"""B200-native robust rotation averaging (GlobalSfMpy-compatible hot path)."""
