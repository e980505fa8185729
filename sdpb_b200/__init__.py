"""sdpb_b200 — B200-native Schur-complement step of SDPB.

Python is only the test/bench harness here: this module is a ctypes binding of
the C-ABI in ``include/sdpb_b200.h`` (``sdpb_b200/libsdpb_b200.so``, built by
``make -C sdpb_b200/csrc`` or ``__graft_entry__.build()``).  The product is the
CUDA library and the C++ host solver under ``sdpb_b200/csrc``; there is no
CPU fallback anywhere in this package — creating a context without a CUDA
device raises.
"""
from .capi import (  # noqa: F401
    SchurContext,
    SdpbB200Error,
    BlockShape,
    elem_words,
    stored_limbs,
    load_library,
    LIB_PATH,
)
