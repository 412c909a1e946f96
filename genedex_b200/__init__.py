"""genedex_b200 -- B200-native batched FM-index search behind genedex's API.

The product is the CUDA library genedex_b200/csrc/libgenedex_b200.so (C ABI: include/genedex_b200.h);
this package mirrors the crate's public interface on top of it.  Importing fails if the library
has not been built -- there is no CPU fallback.
"""
from . import _lib, alphabet
from .alphabet import Alphabet
from .index import (Cursor, FmIndex, FmIndexConfig, GenedexError, Hit, InvalidSymbolError,
                    PerformancePriority, pack_queries)

_lib.load()  # fail loudly when the CUDA extension is missing

__all__ = ["Alphabet", "Cursor", "FmIndex", "FmIndexConfig", "GenedexError", "Hit", "InvalidSymbolError",
           "PerformancePriority", "alphabet", "pack_queries"]
