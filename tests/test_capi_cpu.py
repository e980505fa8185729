"""CPU-side checks of the C-ABI library: it loads, exports every symbol the
header declares, and refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import sdpb_b200
from sdpb_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdpb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdpb_b200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdpb_b200.h but not exported"


def test_element_format_sizes():
    lib = capi.load_library()
    for prec, nl in ((128, 4), (256, 6), (448, 9), (664, 13), (768, 14), (960, 17), (1536, 26)):
        assert lib.sdpb_b200_stored_limbs(prec) == nl
        assert lib.sdpb_b200_elem_words(prec) == capi.elem_words(prec) == ((nl + 2) & ~1)


def test_pack_unpack_roundtrip():
    import ctypes
    lib = capi.load_library()
    lib.sdpb_b200_pack_mpf.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, capi.u64p, capi.u64p]
    lib.sdpb_b200_pack_mpf.restype = None
    lib.sdpb_b200_unpack_mpf.argtypes = [ctypes.c_int, capi.u64p, capi.u64p, ctypes.POINTER(ctypes.c_long)]
    lib.sdpb_b200_unpack_mpf.restype = ctypes.c_int
    prec = 768
    limbs = np.array([3, 0, 7, 9], dtype=np.uint64)  # size 4, low limb nonzero
    out = np.zeros(capi.elem_words(prec), dtype=np.uint64)
    lib.sdpb_b200_pack_mpf(prec, -4, -2, capi._ptr(limbs), capi._ptr(out))
    assert int(out[0]) == ((0xFFFFFFFF << 32) | 0xFFFFFFFE)  # sign -1, exp -2
    assert list(out[1 + 14 - 4:1 + 14]) == [3, 0, 7, 9]
    back = np.zeros(14, dtype=np.uint64)
    e = ctypes.c_long()
    size = lib.sdpb_b200_unpack_mpf(prec, capi._ptr(out), capi._ptr(back), ctypes.byref(e))
    assert size == -4 and e.value == -2 and list(back[:4]) == [3, 0, 7, 9]


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(sdpb_b200.SdpbB200Error) as ei:
        sdpb_b200.SchurContext(768, [(1, 4)], 3)
    assert "no CPU fallback" in str(ei.value)


def test_unsupported_precision_is_rejected():
    with pytest.raises(sdpb_b200.SdpbB200Error):
        sdpb_b200.SchurContext(64 * 40, [(1, 4)], 3)


def test_makefile_tracks_every_kernel_header():
    """A header missing from HDRS means `make` keeps stale objects after an edit (the GPU box then
    runs old kernels): every .h / .cuh next to the kernels must be a dependency."""
    import os
    import re
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sdpb_b200", "csrc")
    text = open(os.path.join(csrc, "Makefile")).read()
    hdrs = re.search(r"^HDRS := (.*)$", text, re.M).group(1).split()
    for name in os.listdir(csrc):
        if name.endswith((".h", ".cuh")):
            assert name in hdrs, f"{name} is not listed in HDRS of sdpb_b200/csrc/Makefile"


def test_ptr_array_passes_prebuilt_tables_through():
    import ctypes
    import numpy as np
    from sdpb_b200.capi import ptr_array
    blocks = [np.zeros((2, 2, 4), dtype=np.uint64), np.zeros((0, 0, 4), dtype=np.uint64), None]
    table = ptr_array(blocks)
    assert isinstance(table, ctypes.Array) and len(table) == 3
    assert ctypes.addressof(table[0].contents) == blocks[0].ctypes.data
    assert not table[1] and not table[2]          # empty / missing block -> NULL
    assert ptr_array(table) is table              # built once, reused by every later call
