"""ctypes binding of include/sdpb_b200.h.

Matrices are numpy ``uint64`` arrays of shape ``(width, height, elem_words)``:
column-major packed elements exactly as the C-ABI wants them.  Method names and
argument meaning mirror the reference functions they replace
(src/sdp_solve/SDP_Solver/run/run.cxx:14-17,37-45 and run/step/step.cxx:12-24).
"""
import ctypes
import os
from dataclasses import dataclass

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsdpb_b200.so")
_lib = None

u64p = ctypes.POINTER(ctypes.c_uint64)
u64pp = ctypes.POINTER(u64p)


class SdpbB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"sdpb_b200 error {code}: {message}")
        self.code = code
        self.message = message


def load_library():
    """Load libsdpb_b200.so; fails loudly if the CUDA extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C sdpb_b200/csrc` "
            "(there is no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.sdpb_b200_elem_words.restype = ctypes.c_int
    lib.sdpb_b200_elem_words.argtypes = [ctypes.c_int]
    lib.sdpb_b200_stored_limbs.restype = ctypes.c_int
    lib.sdpb_b200_stored_limbs.argtypes = [ctypes.c_int]
    lib.sdpb_b200_create.restype = ctypes.c_int
    lib.sdpb_b200_create.argtypes = [
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_int,
        ctypes.c_char_p, ctypes.c_size_t]
    lib.sdpb_b200_destroy.restype = None
    lib.sdpb_b200_destroy.argtypes = [ctypes.c_void_p]
    lib.sdpb_b200_last_error.restype = ctypes.c_char_p
    lib.sdpb_b200_last_error.argtypes = [ctypes.c_void_p]
    lib.sdpb_b200_set_block.restype = ctypes.c_int
    lib.sdpb_b200_set_block.argtypes = [ctypes.c_void_p, ctypes.c_int, u64p, u64p, u64p]
    lib.sdpb_b200_cholesky_decomposition.restype = ctypes.c_int
    lib.sdpb_b200_cholesky_decomposition.argtypes = [ctypes.c_void_p, ctypes.c_int, u64pp, u64pp]
    lib.sdpb_b200_compute_bilinear_pairings.restype = ctypes.c_int
    lib.sdpb_b200_compute_bilinear_pairings.argtypes = [ctypes.c_void_p, u64pp, u64pp, u64pp]
    lib.sdpb_b200_initialize_schur_complement_solver.restype = ctypes.c_int
    lib.sdpb_b200_initialize_schur_complement_solver.argtypes = [
        ctypes.c_void_p, u64pp, u64pp, u64p, ctypes.POINTER(ctypes.c_int32)]
    lib.sdpb_b200_solve_schur_complement_equation.restype = ctypes.c_int
    lib.sdpb_b200_solve_schur_complement_equation.argtypes = [ctypes.c_void_p, u64pp, u64p]
    lib.sdpb_b200_scale_multiply_add.restype = ctypes.c_int
    lib.sdpb_b200_scale_multiply_add.argtypes = [ctypes.c_void_p, ctypes.c_int, u64pp, u64pp, ctypes.c_int, u64pp]
    lib.sdpb_b200_last_solve_ms.restype = ctypes.c_float
    lib.sdpb_b200_last_solve_ms.argtypes = [ctypes.c_void_p]
    lib.sdpb_b200_schur_step.restype = ctypes.c_int
    lib.sdpb_b200_schur_step.argtypes = [ctypes.c_void_p] + [u64pp] * 8 + [
        u64p, ctypes.POINTER(ctypes.c_int32)]
    lib.sdpb_b200_upload_XY.restype = ctypes.c_int
    lib.sdpb_b200_upload_XY.argtypes = [ctypes.c_void_p, u64pp, u64pp]
    lib.sdpb_b200_schur_step_resident.restype = ctypes.c_int
    lib.sdpb_b200_schur_step_resident.argtypes = [ctypes.c_void_p]
    lib.sdpb_b200_download.restype = ctypes.c_int
    lib.sdpb_b200_download.argtypes = [ctypes.c_void_p] + [u64pp] * 6 + [u64p]
    lib.sdpb_b200_set_concurrency.restype = ctypes.c_int
    lib.sdpb_b200_set_concurrency.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.sdpb_b200_last_timings_ms.restype = ctypes.c_int
    lib.sdpb_b200_last_timings_ms.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    lib.sdpb_b200_kernel_launches.restype = ctypes.c_long
    lib.sdpb_b200_kernel_launches.argtypes = [ctypes.c_void_p]
    lib.sdpb_b200_kernel_timings.restype = ctypes.c_int
    lib.sdpb_b200_kernel_timings.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                             ctypes.POINTER(ctypes.c_float)]
    lib.sdpb_b200_host_alloc.restype = ctypes.c_int
    lib.sdpb_b200_host_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    lib.sdpb_b200_host_free.restype = None
    lib.sdpb_b200_host_free.argtypes = [ctypes.c_void_p]
    lib.sdpb_b200_comm_get_unique_id.restype = ctypes.c_int
    lib.sdpb_b200_comm_get_unique_id.argtypes = [ctypes.c_void_p]
    lib.sdpb_b200_comm_init.restype = ctypes.c_int
    lib.sdpb_b200_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.sdpb_b200_comm_init_local.restype = ctypes.c_int
    lib.sdpb_b200_comm_init_local.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(ctypes.POINTER(ctypes.c_int))]
    lib.sdpb_b200_cholesky_diagonals.restype = ctypes.c_int
    lib.sdpb_b200_cholesky_diagonals.argtypes = [ctypes.c_void_p, u64p, u64p, u64p, u64p]
    lib.sdpb_b200_direction_begin.restype = ctypes.c_int
    lib.sdpb_b200_direction_begin.argtypes = [ctypes.c_void_p, u64p]
    lib.sdpb_b200_direction_R_errors.restype = ctypes.c_int
    lib.sdpb_b200_direction_R_errors.argtypes = [ctypes.c_void_p, u64p, u64p]
    lib.sdpb_b200_direction_set_residues.restype = ctypes.c_int
    lib.sdpb_b200_direction_set_residues.argtypes = [ctypes.c_void_p, u64pp, u64pp, u64p]
    lib.sdpb_b200_compute_search_direction.restype = ctypes.c_int
    lib.sdpb_b200_compute_search_direction.argtypes = [ctypes.c_void_p, u64p, ctypes.c_int]
    lib.sdpb_b200_direction_frobenius.restype = ctypes.c_int
    lib.sdpb_b200_direction_frobenius.argtypes = [ctypes.c_void_p, u64p]
    lib.sdpb_b200_direction_get.restype = ctypes.c_int
    lib.sdpb_b200_direction_get.argtypes = [ctypes.c_void_p, u64pp, u64pp, u64p, u64pp]
    lib.sdpb_b200_direction_put.restype = ctypes.c_int
    lib.sdpb_b200_direction_put.argtypes = [ctypes.c_void_p, u64pp, u64pp]
    lib.sdpb_b200_step_length_iterations.restype = ctypes.c_int
    lib.sdpb_b200_step_length_iterations.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    lib.sdpb_b200_step_length.restype = ctypes.c_int
    lib.sdpb_b200_step_length.argtypes = [ctypes.c_void_p, ctypes.c_int, u64p]
    lib.sdpb_b200_last_step_length_ms.restype = ctypes.c_float
    lib.sdpb_b200_last_step_length_ms.argtypes = [ctypes.c_void_p]
    lib.sdpb_b200_last_direction_ms.restype = ctypes.c_float
    lib.sdpb_b200_last_direction_ms.argtypes = [ctypes.c_void_p]
    lib.sdpb_b200_scalar_op.restype = ctypes.c_int
    lib.sdpb_b200_scalar_op.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_long, u64p, u64p, u64p]
    _lib = lib
    return lib


def stored_limbs(prec_bits):
    return (prec_bits + 63) // 64 + 2


def elem_words(prec_bits):
    return (stored_limbs(prec_bits) + 2) & ~1


@dataclass(frozen=True)
class BlockShape:
    """Shape helpers of one SDP block (reference Block_Info.hxx:54-115)."""
    m: int  # dimensions[j]
    n: int  # num_points[j]

    @property
    def schur_size(self):
        return self.n * self.m * (self.m + 1) // 2

    def psd_size(self, parity):
        even = self.m * ((self.n + 1) // 2)
        return even if parity == 0 else self.m * self.n - even

    @property
    def pairing_size(self):
        return self.m * self.n

    def basis_height(self, parity):
        degree = self.n - 1
        return (degree + parity) // 2 + 1 - parity


def _ptr(a):
    return a.ctypes.data_as(u64p) if a is not None and a.size else None


def ptr_array(arrays):
    """C array of uint64_t* from a list of numpy arrays (None / empty -> NULL).  A ctypes array
    built by an earlier call passes through: marshalling 1200 block pointers costs milliseconds of
    Python, which a caller that reuses its buffers (every solver iteration does) pays once."""
    if isinstance(arrays, ctypes.Array):
        return arrays
    arr = (u64p * max(1, len(arrays)))()
    for i, a in enumerate(arrays):
        arr[i] = _ptr(a) if a is not None else None
    return arr


class PinnedPool:
    """Page-locked host arrays from sdpb_b200_host_alloc (freed on close)."""

    def __init__(self):
        self.lib = load_library()
        self.ptrs = []

    def empty(self, shape):
        nbytes = int(np.prod(shape)) * 8
        if nbytes == 0:
            return np.zeros(shape, dtype=np.uint64)
        p = ctypes.c_void_p()
        rc = self.lib.sdpb_b200_host_alloc(ctypes.byref(p), nbytes)
        if rc != 0:
            raise SdpbB200Error(rc, "sdpb_b200_host_alloc failed")
        self.ptrs.append(p)
        buf = (ctypes.c_uint64 * (nbytes // 8)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint64).reshape(shape)

    def slab(self, shapes):
        """Arrays of the given shapes carved back to back out of ONE page-locked allocation
        (what a reference-side shim would pack El::BigFloat blocks into): the library then
        moves adjacent blocks with a single DMA."""
        sizes = [int(np.prod(s)) for s in shapes]
        total = sum(sizes)
        if total == 0:
            return [np.zeros(s, dtype=np.uint64) for s in shapes]
        flat = self.empty((total,))
        out, pos = [], 0
        for s, n in zip(shapes, sizes):
            out.append(flat[pos:pos + n].reshape(s))
            pos += n
        return out

    def like(self, arrays):
        out = []
        for a in arrays:
            b = self.empty(a.shape)
            b[...] = a
            out.append(b)
        return out

    def close(self):
        for p in self.ptrs:
            self.lib.sdpb_b200_host_free(p)
        self.ptrs = []


class StepContextBase:
    """Shared driver logic over a C library exposing the sdpb_b200 call surface."""

    def __init__(self, prec_bits, shapes, N):
        self.prec = prec_bits
        self.shapes = [s if isinstance(s, BlockShape) else BlockShape(*s) for s in shapes]
        self.N = N
        self.ew = elem_words(prec_bits)
        self.J = len(self.shapes)

    # allocation helpers -------------------------------------------------
    def empty(self, h, w):
        return np.zeros((w, h, self.ew), dtype=np.uint64)

    def alloc_psd_blocks(self):
        return [self.empty(s.psd_size(p), s.psd_size(p)) for s in self.shapes for p in (0, 1)]

    def alloc_pairing_blocks(self):
        return [self.empty(s.pairing_size, s.pairing_size) for s in self.shapes for p in (0, 1)]

    def alloc_solve_vectors(self):
        return [self.empty(s.schur_size, 1) for s in self.shapes], self.empty(self.N, 1)

    def alloc_schur_outputs(self):
        L = [self.empty(s.schur_size, s.schur_size) for s in self.shapes]
        P = [self.empty(s.schur_size, self.N) for s in self.shapes]
        Q = self.empty(self.N, self.N)
        return L, P, Q

    # ---- the search direction on the resident state (compute_search_direction.cxx:44-90) ----
    def _dir(self, name):
        return getattr(self.lib, self.PREFIX + name)

    def direction_begin(self):
        """minus_XY = -X Y on the X, Y of the last step; returns the per-block traces (2J, ew)."""
        out = np.zeros((2 * self.J, self.ew), dtype=np.uint64)
        self._check(self._dir("direction_begin")(self.handle, _ptr(out)))
        return out

    def direction_R_errors(self, mu):
        out = np.zeros((2 * self.J, self.ew), dtype=np.uint64)
        self._check(self._dir("direction_R_errors")(self.handle, _ptr(np.ascontiguousarray(mu)), _ptr(out)))
        return out

    def direction_set_residues(self, primal_residues, dual_residues, primal_residue_p):
        self._check(self._dir("direction_set_residues")(self.handle, ptr_array(primal_residues),
                                                         ptr_array(dual_residues), _ptr(primal_residue_p)))

    def compute_search_direction(self, beta_mu, is_corrector):
        self._check(self._dir("compute_search_direction")(self.handle, _ptr(np.ascontiguousarray(beta_mu)),
                                                           int(bool(is_corrector))))

    def direction_frobenius(self):
        out = np.zeros((2 * self.J, self.ew), dtype=np.uint64)
        self._check(self._dir("direction_frobenius")(self.handle, _ptr(out)))
        return out

    def direction_get(self):
        """(dx [J vectors], dX [2J blocks], dy, dY [2J blocks]) of the last compute_search_direction."""
        dx, dy = self.alloc_solve_vectors()
        dX, dY = self.alloc_psd_blocks(), self.alloc_psd_blocks()
        self._check(self._dir("direction_get")(self.handle, ptr_array(dx), ptr_array(dX), _ptr(dy), ptr_array(dY)))
        return dx, dX, dy, dY


    def direction_put(self, dX, dY):
        """Replace the resident dX, dY (2J blocks each) by the caller's."""
        self._check(self._dir("direction_put")(self.handle, ptr_array(dX), ptr_array(dY)))

    def step_length(self, which):
        """Per block-parity smallest eigenvalue of L^-1 dM L^-T (step_length.cxx:27-46), M = X / dX
        (which 0) or Y / dY (which 1) of the last compute_search_direction; (2J, ew), empty blocks 0."""
        out = np.zeros((2 * self.J, self.ew), dtype=np.uint64)
        self._check(self._dir("step_length")(self.handle, int(which), _ptr(out)))
        return out


class SchurContext(StepContextBase):
    """Device-resident state of the Schur-complement step on one B200."""
    PREFIX = "sdpb_b200_"

    def __init__(self, prec_bits, shapes, N, device=0):
        super().__init__(prec_bits, shapes, N)
        self.lib = load_library()
        self.handle = ctypes.c_void_p()
        dims = (ctypes.c_int * max(1, self.J))(*[s.m for s in self.shapes])
        npts = (ctypes.c_int * max(1, self.J))(*[s.n for s in self.shapes])
        err = ctypes.create_string_buffer(512)
        rc = self.lib.sdpb_b200_create(ctypes.byref(self.handle), prec_bits, device, self.J,
                                       dims, npts, N, err, len(err))
        if rc != 0:
            self.handle = None
            raise SdpbB200Error(rc, err.value.decode())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sdpb_b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SdpbB200Error(rc, self.lib.sdpb_b200_last_error(self.handle).decode())

    def set_block(self, j, B, bases_even, bases_odd):
        self._check(self.lib.sdpb_b200_set_block(self.handle, j, _ptr(B), _ptr(bases_even), _ptr(bases_odd)))

    def cholesky_decomposition(self, which, A, L=None):
        self._check(self.lib.sdpb_b200_cholesky_decomposition(
            self.handle, which, ptr_array(A), ptr_array(L) if L is not None else None))

    def compute_bilinear_pairings(self, Y, A_X_inv=None, A_Y=None):
        self._check(self.lib.sdpb_b200_compute_bilinear_pairings(
            self.handle, ptr_array(Y),
            ptr_array(A_X_inv) if A_X_inv is not None else None,
            ptr_array(A_Y) if A_Y is not None else None))

    def initialize_schur_complement_solver(self, L=None, P=None, Q=None, block_timings_ms=None):
        bt = block_timings_ms.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if block_timings_ms is not None else None
        self._check(self.lib.sdpb_b200_initialize_schur_complement_solver(
            self.handle,
            ptr_array(L) if L is not None else None,
            ptr_array(P) if P is not None else None,
            _ptr(Q) if Q is not None else None, bt))

    def solve_schur_complement_equation(self, dx, dy):
        """In place: dx[j] (1, P_j, ew) holds r_x -> dx, dy (1, N, ew) holds r_y -> dy
        (solve_schur_complement_equation.cxx:16-79) on the device-resident factors."""
        self._check(self.lib.sdpb_b200_solve_schur_complement_equation(self.handle, ptr_array(dx), _ptr(dy)))

    def scale_multiply_add(self, alpha, A, B, beta, C):
        """C[b] = alpha A[b] B[b] + beta C[b] on the PSD-shaped blocks (scale_multiply_add.cxx:4-16);
        alpha in (1, -1), beta in (0, 1); C is overwritten."""
        self._check(self.lib.sdpb_b200_scale_multiply_add(self.handle, int(alpha), ptr_array(A), ptr_array(B),
                                                          int(beta), ptr_array(C)))

    def last_solve_ms(self):
        return float(self.lib.sdpb_b200_last_solve_ms(self.handle))

    def last_step_length_ms(self):
        return float(self.lib.sdpb_b200_last_step_length_ms(self.handle))

    def step_length_iterations(self):
        """Laguerre steps per block-parity of the last step_length call."""
        out = (ctypes.c_int * max(1, 2 * self.J))()
        self._check(self.lib.sdpb_b200_step_length_iterations(self.handle, out))
        return list(out)[:2 * self.J]

    def last_direction_ms(self):
        return float(self.lib.sdpb_b200_last_direction_ms(self.handle))

    def schur_step(self, X, Y, X_chol=None, Y_chol=None, A_X_inv=None, A_Y=None, L=None, P=None, Q=None,
                   block_timings_ms=None):
        bt = block_timings_ms.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if block_timings_ms is not None else None
        opt = lambda v: ptr_array(v) if v is not None else None  # noqa: E731
        self._check(self.lib.sdpb_b200_schur_step(
            self.handle, ptr_array(X), ptr_array(Y), opt(X_chol), opt(Y_chol), opt(A_X_inv), opt(A_Y),
            opt(L), opt(P), _ptr(Q) if Q is not None else None, bt))

    def upload_XY(self, X, Y):
        self._check(self.lib.sdpb_b200_upload_XY(self.handle, ptr_array(X), ptr_array(Y)))

    def schur_step_resident(self):
        self._check(self.lib.sdpb_b200_schur_step_resident(self.handle))

    def download(self, X_chol=None, Y_chol=None, A_X_inv=None, A_Y=None, L=None, P=None, Q=None):
        opt = lambda v: ptr_array(v) if v is not None else None  # noqa: E731
        self._check(self.lib.sdpb_b200_download(
            self.handle, opt(X_chol), opt(Y_chol), opt(A_X_inv), opt(A_Y), opt(L), opt(P),
            _ptr(Q) if Q is not None else None))

    # multi-GPU ----------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id (call on rank 0, then broadcast it to the other ranks)."""
        lib = load_library()
        buf = ctypes.create_string_buffer(128)
        rc = lib.sdpb_b200_comm_get_unique_id(buf)
        if rc != 0:
            raise SdpbB200Error(rc, "sdpb_b200_comm_get_unique_id failed (is libnccl.so.2 loadable?)")
        return buf.raw

    def comm_init(self, rank, world, unique_id, num_blocks_global, global_block_index):
        """Join the communicator; this context holds the blocks `global_block_index` of
        `num_blocks_global` (Block_Info::block_indices in the reference)."""
        assert len(global_block_index) == self.J and len(unique_id) == 128
        idx = (ctypes.c_int * max(1, self.J))(*global_block_index)
        self._check(self.lib.sdpb_b200_comm_init(self.handle, rank, world, unique_id, num_blocks_global, idx))

    def comm_init_from_torch(self, dist, rank, world, num_blocks_global, global_block_index):
        """Rendezvous through an initialised torch.distributed group (any backend)."""
        import torch
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.frombuffer(bytearray(self.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(t, 0)
        self.comm_init(rank, world, bytes(t.cpu().numpy().tobytes()), num_blocks_global, global_block_index)

    @staticmethod
    def comm_init_local(contexts, num_blocks_global, global_block_index):
        """All ranks inside this process on one device (sdpb_b200_comm_init_local): contexts[r]
        holds the blocks global_block_index[r].  Collective calls must then be made from one host
        thread per context at the same time."""
        lib = load_library()
        world = len(contexts)
        handles = (ctypes.c_void_p * world)(*[c.handle for c in contexts])
        keep = [(ctypes.c_int * max(1, len(ix)))(*ix) for ix in global_block_index]
        idx = (ctypes.POINTER(ctypes.c_int) * world)(*[ctypes.cast(k, ctypes.POINTER(ctypes.c_int)) for k in keep])
        rc = lib.sdpb_b200_comm_init_local(handles, world, num_blocks_global, idx)
        if rc != 0:
            raise SdpbB200Error(rc, lib.sdpb_b200_last_error(contexts[0].handle).decode())

    def cholesky_diagonals(self, X_diag=None, Y_diag=None, S_diag=None, Q_diag=None):
        """Diagonals of the resident Cholesky factors (all update_cond_numbers reads of them)."""
        self._check(self.lib.sdpb_b200_cholesky_diagonals(self.handle, _ptr(X_diag), _ptr(Y_diag), _ptr(S_diag),
                                                          _ptr(Q_diag)))

    def set_concurrency(self, level):
        """0: one stream, program order (per-kernel timing mode); 1: concurrent chains (default)."""
        self._check(self.lib.sdpb_b200_set_concurrency(self.handle, int(level)))

    def kernel_launches(self):
        return int(self.lib.sdpb_b200_kernel_launches(self.handle))

    def kernel_timings(self):
        """[(kernel name, ms)] of the last resident step, in launch order."""
        cap = 4096
        names = (ctypes.c_char_p * cap)()
        ms = (ctypes.c_float * cap)()
        n = self.lib.sdpb_b200_kernel_timings(self.handle, cap, names, ms)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def last_timings_ms(self):
        ms = (ctypes.c_float * 9)()
        self.lib.sdpb_b200_last_timings_ms(self.handle, ms, 9)
        return list(ms)

    def scalar_op(self, op, a, b, k=0):
        r = np.zeros_like(a)
        count = a.size // self.ew
        self._check(self.lib.sdpb_b200_scalar_op(self.handle, op, k, count, _ptr(a), _ptr(b), _ptr(r)))
        return r
