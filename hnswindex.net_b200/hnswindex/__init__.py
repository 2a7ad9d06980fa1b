"""hnswindex -- Python face of the B200-native HNSW engine (same public API as the reference's `hnswindex` package)."""
from .bindings import Index, lib, last_error, LIB_PATH

__all__ = ["Index", "lib", "last_error", "LIB_PATH"]
