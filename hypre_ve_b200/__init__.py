"""hypre_ve_b200 -- Python binding (ctypes) of libhypre_b200.so, the B200-native implementation of
BoomerAMG's data-parallel hot path.  The product is the shared library and its C-ABI
(include/hypre_b200.h); this module only loads it for tests and bench.py.

There is NO CPU fallback: loading fails loudly when the CUDA library has not been built, and
every call raises B200Error when the device path reports an error.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhypre_b200.so")


class B200Error(RuntimeError):
    pass


_lib = None

_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)

# name -> (restype, argtypes): mirrors include/hypre_b200.h one to one
SIGNATURES = {
    "b200_init": (_i, [_i, C.POINTER(_vp)]),
    "b200_finalize": (_i, [_vp]),
    "b200_last_error": (C.c_char_p, []),
    "b200_stream": (_vp, [_vp]),
    "b200_sync": (_i, [_vp]),
    "b200_malloc": (_i, [_vp, C.POINTER(_vp), _sz]),
    "b200_free": (_i, [_vp, _vp]),
    "b200_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz]),
    "b200_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "b200_memcpy_d2d": (_i, [_vp, _vp, _vp, _sz]),
    "b200_memset": (_i, [_vp, _vp, _i, _sz]),
    "b200_launch_count": (C.c_longlong, []),
    "b200_pool_trim": (_i, [_vp, C.POINTER(_sz)]),
    "b200_pool_stats": (_i, [_vp, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "b200_timer_start": (_i, [_vp]),
    "b200_timer_stop_ms": (_i, [_vp, _dp]),
    "b200_csr_create": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, C.POINTER(_vp)]),
    "b200_csr_create_from_host": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_csr_destroy": (_i, [_vp, _vp]),
    "b200_csr_dims": (_i, [_vp, _ip, _ip, _ip]),
    "b200_csr_stream_bytes_per_entry": (_i, [_vp]),
    "b200_dist_matrix_stream_bytes_per_entry": (_i, [_vp]),
    "b200_csr_download": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "b200_csr_matvec": (_i, [_vp, _d, _vp, _vp, _d, _vp, _vp]),
    "b200_csr_matvecT": (_i, [_vp, _d, _vp, _vp, _d, _vp, _vp]),
    "b200_csr_row_stats": (_i, [_vp, _vp, _ip, _ip, _dp, _dp, _dp, _dp]),
    "b200_csr_sorted_copy": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "b200_csr_transpose": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "b200_csr_multiply": (_i, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_vec_fill": (_i, [_vp, _i, _d, _vp]),
    "b200_vec_copy": (_i, [_vp, _i, _vp, _vp]),
    "b200_vec_scale": (_i, [_vp, _i, _d, _vp]),
    "b200_vec_axpy": (_i, [_vp, _i, _d, _vp, _vp]),
    "b200_vec_dot": (_i, [_vp, _i, _vp, _vp, _dp]),
    "b200_generate_laplacian": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _dp, C.POINTER(_vp)]),
    "b200_generate_laplacian27": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _dp, C.POINTER(_vp)]),
    "b200_generate_difconv": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _dp, C.POINTER(_vp)]),
    "b200_generate_rotate7pt": (_i, [_vp, _i, _i, _i, _i, _i, _i, _d, _d, C.POINTER(_vp)]),
    "b200_ij_create": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp)]),
    "b200_ij_create_rows": (_i, [_vp, _i, _i, _i, C.POINTER(_vp)]),
    "b200_ij_destroy": (_i, [_vp, _vp]),
    "b200_ij_set_values": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _ip]),
    "b200_ij_assemble": (_i, [_vp, _vp, C.POINTER(_vp), _ip]),
    "b200_ij_num_rejected": (C.c_longlong, [_vp]),
    "b200_parcsr_create_from_host": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_parcsr_destroy": (_i, [_vp, _vp]),
    "b200_parcsr_local_rows": (_i, [_vp, _ip, _ip, _ip, _ip]),
    "b200_parcsr_diag": (_vp, [_vp]),
    "b200_parcsr_offd": (_vp, [_vp]),
    "b200_parcsr_matvec": (_i, [_vp, _d, _vp, _vp, _d, _vp, _vp]),
    "b200_amg_create": (_i, [C.POINTER(_vp)]),
    "b200_amg_destroy": (_i, [_vp, _vp]),
    "b200_amg_set_int": (_i, [_vp, C.c_char_p, _i]),
    "b200_amg_set_real": (_i, [_vp, C.c_char_p, _d]),
    "b200_amg_setup": (_i, [_vp, _vp, _vp]),
    "b200_amg_solve": (_i, [_vp, _vp, _vp, _vp]),
    "b200_amg_solve_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _ip, _dp]),
    "b200_amg_num_levels": (_i, [_vp]),
    "b200_amg_level_A": (_vp, [_vp, _i]),
    "b200_amg_level_P": (_vp, [_vp, _i]),
    "b200_amg_level_S": (_vp, [_vp, _i]),
    "b200_amg_level_CF": (_vp, [_vp, _i]),
    "b200_amg_level_l1": (_vp, [_vp, _i]),
    "b200_amg_setup_times": (_i, [_vp, _dp]),
    "b200_strength": (_i, [_vp, _vp, _d, _d, C.POINTER(_vp)]),
    "b200_pmis": (_i, [_vp, _vp, _i, _vp]),
    "b200_hmis": (_i, [_vp, _vp, _i, _vp]),
    "b200_extpi_interp": (_i, [_vp, _vp, _vp, _vp, _d, _i, C.POINTER(_vp)]),
    "b200_create_2nd_s": (_i, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_agg_coarsen": (_i, [_vp, _vp, _i, _vp]),
    "b200_multipass_interp": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_l1_norms": (_i, [_vp, _vp, _i, _vp]),
    "b200_l1_norms_blocks": (_i, [_vp, _vp, _i, _i, _vp]),
    "b200_l1_norms_cf": (_i, [_vp, _vp, _vp, _vp]),
    "b200_relax_gs": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "b200_pcg_solve": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _ip, _dp, _vp]),
    "b200_pcg_solve_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ip, _dp, _vp]),
    "b200_parcsr_diag_scale": (_i, [_vp, _vp, _vp, _vp]),
    "b200_gmres_solve": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ip, _dp, _vp, _ip]),
    "b200_bicgstab_solve": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ip, _dp, _vp, _ip]),
    "b200_comm_create_single": (_i, [C.POINTER(_vp)]),
    "b200_comm_group_create": (_i, [_i, C.POINTER(_vp)]),
    "b200_comm_group_destroy": (_i, [_vp]),
    "b200_comm_group_abort": (_i, [_vp]),
    "b200_comm_threads_enable_p2p": (_i, [_vp, _vp]),
    "b200_comm_abort": (_i, [_vp]),
    "b200_comm_create_threads": (_i, [_vp, _i, C.POINTER(_vp)]),
    "b200_comm_nccl_unique_id": (_i, [C.c_char_p]),
    "b200_comm_create_nccl": (_i, [_vp, _i, _i, C.c_char_p, C.POINTER(_vp)]),
    "b200_comm_destroy": (_i, [_vp, _vp]),
    "b200_comm_rank": (_i, [_vp]),
    "b200_comm_size": (_i, [_vp]),
    "b200_dist_generate_laplacian": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _dp, C.POINTER(_vp)]),
    "b200_dist_generate_difconv": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _dp, C.POINTER(_vp)]),
    "b200_dist_generate_rotate7pt": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _d, C.POINTER(_vp)]),
    "b200_dist_matrix_create_from_host": (_i, [_vp, _vp, _i, _vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_dist_matrix_create_from_ij": (_i, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_dist_matrix_destroy": (_i, [_vp, _vp]),
    "b200_dist_matrix_info": (_i, [_vp, _ip, _ip, _ip, _ip, _ip, _ip, _ip]),
    "b200_dist_matrix_download": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "b200_dist_matvec": (_i, [_vp, _vp, _d, _vp, _vp, _d, _vp, _vp]),
    "b200_dist_relax_gs": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b200_dist_amg_setup": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_dist_amg_destroy": (_i, [_vp, _vp]),
    "b200_dist_amg_num_levels": (_i, [_vp]),
    "b200_dist_amg_level_A": (_vp, [_vp, _i]),
    "b200_dist_amg_level_P": (_vp, [_vp, _i]),
    "b200_dist_amg_level_view": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(_vp)]),
    "b200_dist_amg_level_cf": (_i, [_vp, _vp, _i, _vp]),
    "b200_dist_amg_setup_ms": (_i, [_vp, _dp]),
    "b200_dist_pcg_solve": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _ip, _dp, _vp]),
    "b200_dist_gmres_solve": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ip, _dp, _vp, _ip]),
    "b200_dist_bicgstab_solve": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ip, _dp, _vp, _ip]),
}


def load_library(path=None):
    """dlopen libhypre_b200.so and attach the C-ABI signatures.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise B200Error(
            "libhypre_b200.so not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback." % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _chk(rc):
    if rc != 0:
        raise B200Error(_lib.b200_last_error().decode())


def _np_ptr(a):
    return a.ctypes.data_as(_vp)


class DeviceArray:
    """A device buffer owned by a Handle (stream-ordered pool allocation)."""

    def __init__(self, handle, n, dtype):
        self.h = handle
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        p = _vp()
        _chk(_lib.b200_malloc(handle.p, C.byref(p), max(1, self.n) * self.dtype.itemsize))
        self.ptr = p

    @classmethod
    def from_numpy(cls, handle, a):
        a = np.ascontiguousarray(a)
        d = cls(handle, a.size, a.dtype)
        if a.size:
            _chk(_lib.b200_memcpy_h2d(handle.p, d.ptr, _np_ptr(a), a.nbytes))
        return d

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.size == self.n
        if a.size:
            _chk(_lib.b200_memcpy_h2d(self.h.p, self.ptr, _np_ptr(a), a.nbytes))

    def numpy(self):
        out = np.empty(self.n, dtype=self.dtype)
        if self.n:
            _chk(_lib.b200_memcpy_d2h(self.h.p, _np_ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            _chk(_lib.b200_free(self.h.p, self.ptr))
            self.ptr = None


def _download_raw(handle, ptr, n, dtype):
    out = np.empty(int(n), dtype=dtype)
    if n:
        _chk(_lib.b200_memcpy_d2h(handle.p, _np_ptr(out), _vp(ptr) if not isinstance(ptr, _vp) else ptr, out.nbytes))
    return out


class Csr:
    def __init__(self, handle, p, owned=True):
        self.h, self.p, self.owned = handle, p, owned

    @classmethod
    def from_host(cls, handle, indptr, indices, data, ncols=None):
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        data = None if data is None else np.ascontiguousarray(data, dtype=np.float64)
        p = _vp()
        nrows = indptr.size - 1
        if ncols is None:
            ncols = int(indices.max()) + 1 if indices.size else 0
        _chk(_lib.b200_csr_create_from_host(handle.p, nrows, ncols, indices.size, _np_ptr(indptr), _np_ptr(indices),
                                            _np_ptr(data) if data is not None else None, C.byref(p)))
        return cls(handle, p)

    def set_ncols(self, n):
        pass

    @property
    def dims(self):
        a, b, c = _i(), _i(), _i()
        _chk(_lib.b200_csr_dims(self.p, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    @property
    def stream_bytes_per_entry(self):
        """12 (int32 column + FP64 value) or 9 / 5 / 2 with the dictionary-compressed solve copy"""
        return _lib.b200_csr_stream_bytes_per_entry(self.p)

    def download(self, with_data=True):
        n, _, nnz = self.dims
        i = np.empty(n + 1, np.int32)
        j = np.empty(nnz, np.int32)
        a = np.empty(nnz, np.float64) if with_data else None
        _chk(_lib.b200_csr_download(self.h.p, self.p, _np_ptr(i), _np_ptr(j), _np_ptr(a) if with_data else None))
        return i, j, a

    def matvec(self, alpha, x, beta, b, y):
        _chk(_lib.b200_csr_matvec(self.h.p, alpha, self.p, x.ptr, beta, b.ptr if b is not None else None, y.ptr))

    def matvecT(self, alpha, x, beta, b, y):
        _chk(_lib.b200_csr_matvecT(self.h.p, alpha, self.p, x.ptr, beta, b.ptr if b is not None else None, y.ptr))

    def sorted_copy(self):
        p = _vp()
        _chk(_lib.b200_csr_sorted_copy(self.h.p, self.p, C.byref(p)))
        return Csr(self.h, p)

    def transpose(self):
        p = _vp()
        _chk(_lib.b200_csr_transpose(self.h.p, self.p, C.byref(p)))
        return Csr(self.h, p)

    def multiply(self, other):
        p = _vp()
        _chk(_lib.b200_csr_multiply(self.h.p, self.p, other.p, C.byref(p)))
        return Csr(self.h, p)

    def destroy(self):
        if self.owned and self.p:
            _chk(_lib.b200_csr_destroy(self.h.p, self.p))
            self.p = None


class _GmresParams(C.Structure):      # b200_gmres_params (include/hypre_b200.h)
    _fields_ = [("tol", _d), ("a_tol", _d), ("cf_tol", _d), ("max_iter", _i), ("min_iter", _i), ("k_dim", _i),
                ("rel_change", _i), ("skip_real_r_check", _i), ("precond", _i)]


class _BicgstabParams(C.Structure):   # b200_bicgstab_params
    _fields_ = [("tol", _d), ("a_tol", _d), ("cf_tol", _d), ("max_iter", _i), ("min_iter", _i), ("stop_crit", _i),
                ("precond", _i)]


def difconv_values(nx, ny, nz, c=(1.0, 1.0, 1.0), a=(1.0, 1.0, 1.0), atype=0):
    """The seven stencil values of -cx Dxx - cy Dyy - cz Dzz + ax Dx + ay Dy + az Dz as the reference driver
    computes them (src/test/ij.c:8266-8409, BuildParDifConv): centre, x-, y-, z-, x+, y+, z+.
    atype 0 forward, 1 backward, 3 upwind, anything else centred differences for the convection term."""
    v = [0.0] * 7
    for d, n in enumerate((nx, ny, nz)):
        hin = 1.0 / float(n + 1)
        sign = lambda t: (0.0 < t) - (0.0 > t)
        if atype in (0, 1, 3):
            back = atype == 1 or (atype == 3 and sign(c[d]) * sign(a[d]) == 1)
            if back:
                v[1 + d] = -c[d] / (hin * hin) - a[d] / hin
                v[4 + d] = -c[d] / (hin * hin)
                if n > 1:
                    v[0] += 2.0 * c[d] / (hin * hin) + 1.0 * a[d] / hin
            else:
                v[1 + d] = -c[d] / (hin * hin)
                v[4 + d] = -c[d] / (hin * hin) + a[d] / hin
                if n > 1:
                    v[0] += 2.0 * c[d] / (hin * hin) - 1.0 * a[d] / hin
        else:
            v[1 + d] = -c[d] / (hin * hin) - a[d] / (2.0 * hin)
            v[4 + d] = -c[d] / (hin * hin) + a[d] / (2.0 * hin)
            if n > 1:
                v[0] += 2.0 * c[d] / (hin * hin)
    return v


class ParCsr:
    def __init__(self, handle, p):
        self.h, self.p = handle, p

    @classmethod
    def laplacian(cls, handle, nx, ny, nz, P=1, Q=1, R=1, p=0, q=0, r=0, c=(1.0, 1.0, 1.0)):
        """ij.c:7788-7810: values = [2(cx+cy+cz) over active dims, -cx, -cy, -cz]"""
        v = (C.c_double * 4)()
        v[1], v[2], v[3] = -c[0], -c[1], -c[2]
        v[0] = (2.0 * c[0] if nx > 1 else 0.0) + (2.0 * c[1] if ny > 1 else 0.0) + (2.0 * c[2] if nz > 1 else 0.0)
        out = _vp()
        _chk(_lib.b200_generate_laplacian(handle.p, nx, ny, nz, P, Q, R, p, q, r, v, C.byref(out)))
        return cls(handle, out)

    @classmethod
    def laplacian27(cls, handle, nx, ny, nz, P=1, Q=1, R=1, p=0, q=0, r=0):
        """ij.c:9078-9086"""
        v = (C.c_double * 2)()
        v[0] = 26.0
        if nx == 1 or ny == 1 or nz == 1:
            v[0] = 8.0
        if nx * ny == 1 or nx * nz == 1 or ny * nz == 1:
            v[0] = 2.0
        v[1] = -1.0
        out = _vp()
        _chk(_lib.b200_generate_laplacian27(handle.p, nx, ny, nz, P, Q, R, p, q, r, v, C.byref(out)))
        return cls(handle, out)

    @classmethod
    def difconv(cls, handle, nx, ny, nz, c=(1.0, 1.0, 1.0), a=(1.0, 1.0, 1.0), atype=0, P=1, Q=1, R=1, p=0, q=0, r=0):
        """`ij -difconv -n nx ny nz -c .. -a .. -atype ..`: GenerateDifConv (par_difconv.c:15), nonsymmetric 7-point"""
        v = (C.c_double * 7)(*difconv_values(nx, ny, nz, c, a, atype))
        out = _vp()
        _chk(_lib.b200_generate_difconv(handle.p, nx, ny, nz, P, Q, R, p, q, r, v, C.byref(out)))
        return cls(handle, out)

    @classmethod
    def rotate7pt(cls, handle, nx, ny, alpha, eps, P=1, Q=1, p=0, q=0):
        """`ij -rotate -n nx ny -alpha A -eps E`: GenerateRotate7pt (par_rotate_7pt.c:15), 2-D rotated anisotropy"""
        out = _vp()
        _chk(_lib.b200_generate_rotate7pt(handle.p, nx, ny, P, Q, p, q, alpha, eps, C.byref(out)))
        return cls(handle, out)

    @classmethod
    def from_host(cls, handle, indptr, indices, data, ncols=None):
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        data = np.ascontiguousarray(data, dtype=np.float64)
        out = _vp()
        n = indptr.size - 1
        _chk(_lib.b200_parcsr_create_from_host(handle.p, n, n if ncols is None else ncols, indices.size,
                                               _np_ptr(indptr), _np_ptr(indices), _np_ptr(data), C.byref(out)))
        return cls(handle, out)

    @property
    def local(self):
        a, b, c, d = _i(), _i(), _i(), _i()
        _chk(_lib.b200_parcsr_local_rows(self.p, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    @property
    def diag(self):
        return Csr(self.h, _vp(_lib.b200_parcsr_diag(self.p)), owned=False)

    @property
    def offd(self):
        return Csr(self.h, _vp(_lib.b200_parcsr_offd(self.p)), owned=False)

    def matvec(self, alpha, x, beta, b, y):
        _chk(_lib.b200_parcsr_matvec(self.h.p, alpha, self.p, x.ptr, beta, b.ptr if b is not None else None, y.ptr))

    def destroy(self):
        if self.p:
            _chk(_lib.b200_parcsr_destroy(self.h.p, self.p))
            self.p = None


class IJAssembler:
    """HYPRE_IJMatrixSetValues / AddToValues / Assemble on the device (b200_ij_*)"""

    def __init__(self, handle, ilower, iupper, jlower=None, jupper=None, global_cols=None):
        """global_cols given: the rows [ilower, iupper] of an operator spread over several ranks (global column ids are
        kept; pass the assembler to DistMatrix.from_ij)"""
        self.h = handle
        self.p = _vp()
        if global_cols is not None:
            _chk(_lib.b200_ij_create_rows(handle.p, ilower, iupper, global_cols, C.byref(self.p)))
        else:
            _chk(_lib.b200_ij_create(handle.p, ilower, iupper, ilower if jlower is None else jlower,
                                     iupper if jupper is None else jupper, C.byref(self.p)))

    def set_values(self, ncols, rows, cols, values, add=False):
        ncols = np.ascontiguousarray(ncols, dtype=np.int32)
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        assert ncols.size == rows.size and cols.size == values.size == int(ncols.sum())
        rej = _i()
        _chk(_lib.b200_ij_set_values(self.h.p, self.p, rows.size, _np_ptr(ncols), _np_ptr(rows), _np_ptr(cols), _np_ptr(values),
                                     1 if add else 0, C.byref(rej)))
        return rej.value

    def assemble(self):
        """returns (ParCsr, number of post-assembly records whose element does not exist)"""
        out, miss = _vp(), _i()
        _chk(_lib.b200_ij_assemble(self.h.p, self.p, C.byref(out), C.byref(miss)))
        return ParCsr(self.h, out), miss.value

    def destroy(self):
        if self.p:
            _chk(_lib.b200_ij_destroy(self.h.p, self.p))
            self.p = None


class Amg:
    """BoomerAMG hierarchy (mirror of the HYPRE_BoomerAMG* solver object)."""

    def __init__(self, handle, **params):
        self.h = handle
        p = _vp()
        _chk(_lib.b200_amg_create(C.byref(p)))
        self.p = p
        for k, v in params.items():
            self.set(k, v)

    def set(self, name, value):
        if isinstance(value, float):
            _chk(_lib.b200_amg_set_real(self.p, name.encode(), value))
        else:
            _chk(_lib.b200_amg_set_int(self.p, name.encode(), int(value)))

    def setup(self, A):
        _chk(_lib.b200_amg_setup(self.h.p, self.p, A.p))

    def solve(self, f, u):
        _chk(_lib.b200_amg_solve(self.h.p, self.p, f.ptr, u.ptr))

    @property
    def num_levels(self):
        return _lib.b200_amg_num_levels(self.p)

    def level_A(self, l):
        return Csr(self.h, _vp(_lib.b200_amg_level_A(self.p, l)), owned=False)

    def level_P(self, l):
        return Csr(self.h, _vp(_lib.b200_amg_level_P(self.p, l)), owned=False)

    def level_S(self, l):
        q = _lib.b200_amg_level_S(self.p, l)
        return Csr(self.h, _vp(q), owned=False) if q else None

    def level_CF(self, l):
        n = self.level_A(l).dims[0]
        return _download_raw(self.h, _lib.b200_amg_level_CF(self.p, l), n, np.int32)

    def level_l1(self, l):
        n = self.level_A(l).dims[0]
        q = _lib.b200_amg_level_l1(self.p, l)
        return _download_raw(self.h, q, n, np.float64) if q else None

    def setup_times(self):
        t = (C.c_double * 8)()
        _chk(_lib.b200_amg_setup_times(self.p, t))
        return list(t)

    def destroy(self):
        if self.p:
            _chk(_lib.b200_amg_destroy(self.h.p, self.p))
            self.p = None


class Handle:
    def __init__(self, device=0):
        load_library()
        p = _vp()
        _chk(_lib.b200_init(device, C.byref(p)))
        self.p = p

    def sync(self):
        _chk(_lib.b200_sync(self.p))

    def array(self, a):
        return DeviceArray.from_numpy(self, a)

    def empty(self, n, dtype=np.float64):
        return DeviceArray(self, n, dtype)

    def zeros(self, n, dtype=np.float64):
        d = DeviceArray(self, n, dtype)
        _chk(_lib.b200_memset(self.p, d.ptr, 0, max(1, d.n) * d.dtype.itemsize))
        return d

    def dot(self, x, y):
        r = _d()
        _chk(_lib.b200_vec_dot(self.p, x.n, x.ptr, y.ptr, C.byref(r)))
        return r.value

    def axpy(self, a, x, y):
        _chk(_lib.b200_vec_axpy(self.p, x.n, a, x.ptr, y.ptr))

    def scale(self, a, y):
        _chk(_lib.b200_vec_scale(self.p, y.n, a, y.ptr))

    def fill(self, x, v):
        _chk(_lib.b200_vec_fill(self.p, x.n, v, x.ptr))

    def copy(self, x, y):
        _chk(_lib.b200_vec_copy(self.p, x.n, x.ptr, y.ptr))

    def timer_start(self):
        _chk(_lib.b200_timer_start(self.p))

    def timer_stop_ms(self):
        ms = _d()
        _chk(_lib.b200_timer_stop_ms(self.p, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return _lib.b200_launch_count()

    def strength(self, A, theta=0.25, max_row_sum=1.0):
        p = _vp()
        _chk(_lib.b200_strength(self.p, A.p, theta, max_row_sum, C.byref(p)))
        return Csr(self, p)

    def pmis(self, S, seed=2747):
        n = S.dims[0]
        cf = self.zeros(n, np.int32)
        _chk(_lib.b200_pmis(self.p, S.p, seed, cf.ptr))
        return cf

    def hmis(self, S, seed=2747):
        """hypre_BoomerAMGCoarsenHMIS: Ruge-Stueben first pass (one device thread) + PMIS seeded with its C points"""
        n = S.dims[0]
        cf = self.zeros(n, np.int32)
        _chk(_lib.b200_hmis(self.p, S.p, seed, cf.ptr))
        return cf

    def extpi_interp(self, A, S, cf, trunc_factor=0.0, max_elmts=4):
        p = _vp()
        _chk(_lib.b200_extpi_interp(self.p, A.p, S.p, cf.ptr, trunc_factor, max_elmts, C.byref(p)))
        return Csr(self, p)

    def create_2nd_s(self, S, cf):
        p = _vp()
        _chk(_lib.b200_create_2nd_s(self.p, S.p, cf.ptr, C.byref(p)))
        return Csr(self, p)

    def agg_coarsen(self, S, cf, seed=2747):
        """second PMIS on the distance-two graph + CorrectCFMarker; cf is updated in place"""
        _chk(_lib.b200_agg_coarsen(self.p, S.p, seed, cf.ptr))

    def multipass_interp(self, A, S, cf):
        p = _vp()
        _chk(_lib.b200_multipass_interp(self.p, A.p, S.p, cf.ptr, C.byref(p)))
        return Csr(self, p)

    def l1_norms(self, A, option=1, blocks=1):
        n = A.dims[0]
        d = self.empty(n)
        _chk(_lib.b200_l1_norms_blocks(self.p, A.p, option, blocks, d.ptr))
        return d

    def relax_gs(self, A, relax_type, f, l1, u, blocks=1):
        """hypre_BoomerAMGRelax for the Gauss-Seidel family (3/4/6 classic, 8/13/14 l1), in place on u;
        blocks = Gauss-Seidel blocks (the reference's OpenMP thread count)"""
        _chk(_lib.b200_relax_gs(self.p, A.p, relax_type, blocks, f.ptr, l1.ptr if l1 is not None else None, u.ptr))

    def pcg(self, A, amg, b, x, tol=1e-8, max_iter=100):
        its = _i()
        rel = _d()
        norms = np.zeros(max_iter + 2, np.float64)
        _chk(_lib.b200_pcg_solve(self.p, A.p, amg.p if amg is not None else None, b.ptr, x.ptr, tol, max_iter,
                                 C.byref(its), C.byref(rel), _np_ptr(norms)))
        return its.value, rel.value, norms[: its.value + 1]

    def gmres(self, A, amg, b, x, tol=1e-8, max_iter=100, k_dim=5, precond=None, a_tol=0.0, min_iter=0,
              skip_real_r_check=0):
        """hypre_GMRESSolve (krylov/gmres.c:226): restarted GMRES(k_dim), right-preconditioned by one BoomerAMG cycle
        (precond 1, default when amg is given), diagonal scaling (2) or nothing (0).
        Returns (iterations, final relative residual, norms[0..iterations], converged)."""
        prm = _GmresParams(tol, a_tol, 0.0, max_iter, min_iter, k_dim, 0, skip_real_r_check,
                           (1 if amg is not None else 0) if precond is None else precond)
        its, rel, conv = _i(), _d(), _i()
        norms = np.zeros(max_iter + 2, np.float64)
        _chk(_lib.b200_gmres_solve(self.p, A.p, amg.p if amg is not None else None, C.byref(prm), b.ptr, x.ptr,
                                   C.byref(its), C.byref(rel), _np_ptr(norms), C.byref(conv)))
        return its.value, rel.value, norms[: its.value + 1], conv.value

    def bicgstab(self, A, amg, b, x, tol=1e-8, max_iter=100, precond=None, a_tol=0.0, min_iter=0):
        """hypre_BiCGSTABSolve (krylov/bicgstab.c:207); same conventions as gmres()"""
        prm = _BicgstabParams(tol, a_tol, 0.0, max_iter, min_iter, 0,
                              (1 if amg is not None else 0) if precond is None else precond)
        its, rel, conv = _i(), _d(), _i()
        norms = np.zeros(max_iter + 2, np.float64)
        _chk(_lib.b200_bicgstab_solve(self.p, A.p, amg.p if amg is not None else None, C.byref(prm), b.ptr, x.ptr,
                                      C.byref(its), C.byref(rel), _np_ptr(norms), C.byref(conv)))
        return its.value, rel.value, norms[: its.value + 1], conv.value

    def trim(self):
        """return the pool's entirely free slabs to the driver; returns the bytes released"""
        n = _sz()
        _chk(_lib.b200_pool_trim(self.p, C.byref(n)))
        return n.value

    def pool_stats(self):
        """(reserved, in_use, peak) bytes of this handle's device pool"""
        a, b, c = _sz(), _sz(), _sz()
        _chk(_lib.b200_pool_stats(self.p, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def close(self):
        """b200_finalize: returns every slab to the driver; objects created on this handle are dead afterwards"""
        if self.p:
            p, self.p = self.p, None
            _chk(_lib.b200_finalize(p))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """Communicator: one rank per GPU (NCCL) or N host threads on one GPU (test backend)."""

    def __init__(self, handle, p):
        self.h, self.p = handle, p

    @classmethod
    def single(cls, handle):
        p = _vp()
        _chk(_lib.b200_comm_create_single(C.byref(p)))
        return cls(handle, p)

    @classmethod
    def threads(cls, handle, group, rank):
        p = _vp()
        _chk(_lib.b200_comm_create_threads(group, rank, C.byref(p)))
        if os.environ.get("B200_P2P_THREADS") == "1":      # direct peer-to-peer halos / reductions between the rank threads
            _chk(_lib.b200_comm_threads_enable_p2p(handle.p, p))
        return cls(handle, p)

    @classmethod
    def nccl(cls, handle, nranks, rank, unique_id):
        p = _vp()
        _chk(_lib.b200_comm_create_nccl(handle.p, nranks, rank, unique_id, C.byref(p)))
        return cls(handle, p)

    @staticmethod
    def nccl_unique_id():
        load_library()
        buf = C.create_string_buffer(128)
        _chk(_lib.b200_comm_nccl_unique_id(buf))
        return buf.raw

    @staticmethod
    def group_create(nranks):
        load_library()
        g = _vp()
        _chk(_lib.b200_comm_group_create(nranks, C.byref(g)))
        return g

    @staticmethod
    def group_abort(group):
        _lib.b200_comm_group_abort(group)

    @staticmethod
    def group_destroy(group):
        _lib.b200_comm_group_destroy(group)

    def abort(self):
        if self.p:
            _lib.b200_comm_abort(self.p)

    @property
    def rank(self):
        return _lib.b200_comm_rank(self.p)

    @property
    def size(self):
        return _lib.b200_comm_size(self.p)

    def destroy(self):
        if self.p:
            _chk(_lib.b200_comm_destroy(self.h.p, self.p))
            self.p = None


class DistMatrix:
    """Row-partitioned operator (ParCSR + CommPkg of the reference) of one rank."""

    def __init__(self, handle, comm, p, owned=True):
        self.h, self.c, self.p, self.owned = handle, comm, p, owned

    @property
    def stream_bytes_per_entry(self):
        return _lib.b200_dist_matrix_stream_bytes_per_entry(self.p)

    @classmethod
    def laplacian(cls, handle, comm, nx, ny, nz, P, Q, R, stencil=7, c=(1.0, 1.0, 1.0)):
        v = (C.c_double * 4)()
        if stencil == 7:
            v[1], v[2], v[3] = -c[0], -c[1], -c[2]
            v[0] = (2.0 * c[0] if nx > 1 else 0.0) + (2.0 * c[1] if ny > 1 else 0.0) + (2.0 * c[2] if nz > 1 else 0.0)
        else:
            v[0] = 26.0
            if nx == 1 or ny == 1 or nz == 1:
                v[0] = 8.0
            if nx * ny == 1 or nx * nz == 1 or ny * nz == 1:
                v[0] = 2.0
            v[1] = -1.0
        out = _vp()
        _chk(_lib.b200_dist_generate_laplacian(handle.p, comm.p, nx, ny, nz, P, Q, R, stencil, v, C.byref(out)))
        return cls(handle, comm, out)

    @classmethod
    def difconv(cls, handle, comm, nx, ny, nz, P, Q, R, c=(1.0, 1.0, 1.0), a=(1.0, 1.0, 1.0), atype=0):
        """GenerateDifConv on the P x Q x R process grid (par_difconv.c:15)"""
        v = (C.c_double * 7)(*difconv_values(nx, ny, nz, c, a, atype))
        out = _vp()
        _chk(_lib.b200_dist_generate_difconv(handle.p, comm.p, nx, ny, nz, P, Q, R, v, C.byref(out)))
        return cls(handle, comm, out)

    @classmethod
    def from_rows(cls, handle, comm, indptr, indices_global, data):
        """this rank's contiguous block of rows (global column ids, diagonal first); collective"""
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices_global = np.ascontiguousarray(indices_global, dtype=np.int32)
        data = np.ascontiguousarray(data, dtype=np.float64)
        out = _vp()
        _chk(_lib.b200_dist_matrix_create_from_host(handle.p, comm.p, indptr.size - 1, _np_ptr(indptr), _np_ptr(indices_global),
                                                    _np_ptr(data), C.byref(out)))
        return cls(handle, comm, out)

    @classmethod
    def from_ij(cls, handle, comm, assembler):
        """device-side assembly of this rank's SetValues / AddToValues records, then localization; collective"""
        out = _vp()
        _chk(_lib.b200_dist_matrix_create_from_ij(handle.p, comm.p, assembler.p, C.byref(out)))
        return cls(handle, comm, out)

    @classmethod
    def rotate7pt(cls, handle, comm, nx, ny, P, Q, alpha, eps):
        """GenerateRotate7pt on the P x Q process grid (par_rotate_7pt.c:15)"""
        out = _vp()
        _chk(_lib.b200_dist_generate_rotate7pt(handle.p, comm.p, nx, ny, P, Q, alpha, eps, C.byref(out)))
        return cls(handle, comm, out)

    @property
    def info(self):
        v = [_i() for _ in range(7)]
        _chk(_lib.b200_dist_matrix_info(self.p, *[C.byref(x) for x in v]))
        k = ["local_rows", "first_row", "global_rows", "local_nnz", "n_ghost", "first_col", "global_cols"]
        return dict(zip(k, [x.value for x in v]))

    def download(self):
        inf = self.info
        i = np.empty(inf["local_rows"] + 1, np.int32)
        j = np.empty(inf["local_nnz"], np.int32)
        a = np.empty(inf["local_nnz"], np.float64)
        _chk(_lib.b200_dist_matrix_download(self.h.p, self.p, _np_ptr(i), _np_ptr(j), _np_ptr(a)))
        return i, j, a

    def vector(self, fill=0.0):
        """device vector with room for the ghost tail"""
        inf = self.info
        d = self.h.zeros(inf["local_rows"] + inf["n_ghost"] + 8)
        d.n_owned = inf["local_rows"]
        if fill != 0.0:
            _chk(_lib.b200_vec_fill(self.h.p, d.n_owned, fill, d.ptr))
        return d

    def matvec(self, alpha, x, beta, b, y):
        _chk(_lib.b200_dist_matvec(self.h.p, self.c.p, alpha, self.p, x.ptr, beta, b.ptr if b is not None else None, y.ptr))

    def relax_gs(self, relax_type, blocks, f, u):
        """hypre_BoomerAMGRelax 8/13/14 across ranks: halo of u, then Gauss-Seidel inside each block"""
        _chk(_lib.b200_dist_relax_gs(self.h.p, self.c.p, self.p, relax_type, blocks, f.ptr, u.ptr))

    def destroy(self):
        if self.owned and self.p:
            _chk(_lib.b200_dist_matrix_destroy(self.h.p, self.p))
            self.p = None


class DistAmg:
    def __init__(self, handle, comm, params, A):
        self.h, self.c = handle, comm
        p = _vp()
        _chk(_lib.b200_dist_amg_setup(handle.p, comm.p, params.p, A.p, C.byref(p)))
        self.p = p

    @property
    def num_levels(self):
        return _lib.b200_dist_amg_num_levels(self.p)

    def _view(self, l, what):
        p = _vp()
        _chk(_lib.b200_dist_amg_level_view(self.h.p, self.c.p, self.p, l, what, C.byref(p)))
        return DistMatrix(self.h, self.c, p, owned=False)

    def level_A(self, l):
        return self._view(l, 0)

    def level_P(self, l):
        return self._view(l, 1)

    def level_cf(self, l):
        n = self.level_A(l).info["local_rows"]
        cf = np.empty(n, np.int32)
        _chk(_lib.b200_dist_amg_level_cf(self.h.p, self.p, l, _np_ptr(cf)))
        return cf

    @property
    def setup_ms(self):
        ms = _d()
        _chk(_lib.b200_dist_amg_setup_ms(self.p, C.byref(ms)))
        return ms.value

    def destroy(self):
        if self.p:
            _chk(_lib.b200_dist_amg_destroy(self.h.p, self.p))
            self.p = None


def dist_pcg(handle, comm, A, amg, b, x, tol=1e-8, max_iter=100):
    its, rel = _i(), _d()
    norms = np.zeros(max_iter + 2, np.float64)
    _chk(_lib.b200_dist_pcg_solve(handle.p, comm.p, A.p, amg.p if amg is not None else None, b.ptr, x.ptr, tol, max_iter,
                                  C.byref(its), C.byref(rel), _np_ptr(norms)))
    return its.value, rel.value, norms[: its.value + 1]


def dist_gmres(handle, comm, A, amg, b, x, tol=1e-8, max_iter=100, k_dim=5):
    """hypre_GMRESSolve across ranks; returns (iterations, final relative residual, norms, converged)"""
    prm = _GmresParams(tol, 0.0, 0.0, max_iter, 0, k_dim, 0, 0, 1 if amg is not None else 0)
    its, rel, conv = _i(), _d(), _i()
    norms = np.zeros(max_iter + 2, np.float64)
    _chk(_lib.b200_dist_gmres_solve(handle.p, comm.p, A.p, amg.p if amg is not None else None, C.byref(prm), b.ptr, x.ptr,
                                    C.byref(its), C.byref(rel), _np_ptr(norms), C.byref(conv)))
    return its.value, rel.value, norms[: its.value + 1], conv.value


def dist_bicgstab(handle, comm, A, amg, b, x, tol=1e-8, max_iter=100):
    """hypre_BiCGSTABSolve across ranks; returns (iterations, final relative residual, norms, converged)"""
    prm = _BicgstabParams(tol, 0.0, 0.0, max_iter, 0, 0, 1 if amg is not None else 0)
    its, rel, conv = _i(), _d(), _i()
    norms = np.zeros(max_iter + 2, np.float64)
    _chk(_lib.b200_dist_bicgstab_solve(handle.p, comm.p, A.p, amg.p if amg is not None else None, C.byref(prm), b.ptr, x.ptr,
                                       C.byref(its), C.byref(rel), _np_ptr(norms), C.byref(conv)))
    return its.value, rel.value, norms[: its.value + 1], conv.value
