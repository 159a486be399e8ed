"""ctypes binding of include/sq.h (the C-ABI replacing tauhost.c's OpenCL section)."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)

SQ_KERNEL_COMPAT1D, SQ_KERNEL_LATTICE = 0, 1
SQ_REAL_F32, SQ_REAL_F64 = 0, 1
SQ_MATH_ACCURATE, SQ_MATH_FAST = 0, 1
SQ_POT_HARMONIC, SQ_POT_DOUBLEWELL, SQ_POT_PHI4 = 0, 3, 4
SQ_FLAG_NO_OBSERVABLES = 1
SQ_FLAG_FORCE_STREAMING = 2
SQ_FLAG_GENERIC_KERNEL = 4
SQ_FLAG_ROWBLOCK_KERNEL = 8


class SqError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        super().__init__(f"{what}: error {code}" + (f" ({detail})" if detail else ""))


class SqParams(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("kernel", C.c_int32), ("real", C.c_int32),
                ("math", C.c_int32), ("potential", C.c_int32), ("ndim", C.c_int32),
                ("dims", C.c_int64 * 4), ("spacing", C.c_double), ("noise_c", C.c_double),
                ("m2", C.c_double), ("lam", C.c_double), ("device", C.c_int32),
                ("nchains", C.c_int32), ("slab_t0", C.c_int64), ("slab_nt", C.c_int64),
                ("reserved", C.c_int32), ("flags", C.c_int32)]


class SqObs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32),
                ("f", C.POINTER(C.c_double)), ("x", C.POINTER(C.c_double)),
                ("xx0", C.POINTER(C.c_double)), ("omega", C.c_double), ("seed", C.c_uint64),
                ("lrgEl", C.c_int32), ("stable", C.c_int32), ("lrgVl", C.c_double),
                ("runs", C.c_int64), ("mean_phi", C.c_double), ("mean_phi2", C.c_double),
                ("slice_x", C.POINTER(C.c_double)), ("slice_xx0", C.POINTER(C.c_double)),
                ("corr", C.POINTER(C.c_double)), ("nclamped", C.c_int64),
                ("nevents", C.c_uint64), ("steps_done", C.c_int64)]


class SqRngEntry(C.Structure):
    _fields_ = [("gid_start", C.c_uint64), ("seed", C.c_uint64), ("ov_gid", C.c_uint64),
                ("ov_t1", C.c_uint64), ("ov_t2", C.c_uint64)]


class SqCompatState(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("lrgEl", C.c_int32), ("seed", C.c_uint64),
                ("lrgVl", C.c_double), ("omega", C.c_double), ("newf_lrgEl", C.c_double)]


class SqFrameRec(C.Structure):
    _fields_ = [("dtau", C.c_double), ("stable", C.c_int32), ("steps", C.c_int32)]


SQ_FRAMES_MAX = 64


def library_path() -> str:
    # SQ_LIBRARY: tuning builds of the same library (tools/); the product is libsq.so
    return os.environ.get("SQ_LIBRARY") or os.path.join(PKG, "libsq.so")


def build(force: bool = False) -> None:
    """Compile libsq.so / tauhost.o in-tree (nvcc -gencode arch=compute_100a,code=sm_100a)."""
    args = ["make", "-C", ROOT, "stochquant_b200/libsq.so", "tauhost.o", "host/libtauhost_io.so"]
    if force:
        args.insert(1, "-B")
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libsq failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])


def exported_symbols() -> list[str]:
    """Entry points declared in include/sq.h."""
    txt = open(os.path.join(ROOT, "include", "sq.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sq_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def load() -> C.CDLL:
    """Load libsq.so.  Raises (never falls back) when the CUDA library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: run `make` (or __graft_entry__.build()); "
                          "stochquant_b200 has no CPU fallback")
    L = C.CDLL(path)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    pd = C.POINTER(C.c_double)
    L.sq_init.restype = i32
    L.sq_init.argtypes = [C.POINTER(vp), C.POINTER(SqParams), pd, pd, pd, dbl, u64]
    L.sq_step.restype = i32
    L.sq_step.argtypes = [vp, dbl, i32, i64, C.POINTER(i32)]
    L.sq_step_async.restype = i32
    L.sq_step_async.argtypes = [vp, dbl, i32, i64]
    L.sq_sync.restype = i32
    L.sq_sync.argtypes = [vp, C.POINTER(i32)]
    L.sq_measure.restype = i32
    L.sq_measure.argtypes = [vp, C.POINTER(SqObs)]
    L.sq_free.restype = None
    L.sq_free.argtypes = [vp]
    L.sq_strerror.restype = C.c_char_p
    L.sq_strerror.argtypes = [i32]
    L.sq_last_cuda_error.restype = C.c_char_p
    L.sq_api_version.restype = i32
    L.sq_device_count.restype = i32
    L.sq_stream.restype = vp
    L.sq_stream.argtypes = [vp]
    L.sq_launch_count.restype = i64
    L.sq_launch_count.argtypes = [vp]
    L.sq_upload_field.restype = i32
    L.sq_upload_field.argtypes = [vp, i32, vp, i32]
    L.sq_download_field.restype = i32
    L.sq_download_field.argtypes = [vp, i32, vp, i32]
    L.sq_set_chain.restype = i32
    L.sq_set_chain.argtypes = [vp, i32, u64, dbl, dbl]
    L.sq_measure_chains.restype = i32
    L.sq_measure_chains.argtypes = [vp, pd, pd, C.POINTER(u64)]
    L.sq_frame_host.restype = i32
    L.sq_frame_host.argtypes = [vp, vp, vp, i32, dbl, i32, i64, C.POINTER(SqObs), C.POINTER(i32)]
    L.sq_debug_draws.restype = i32
    L.sq_debug_draws.argtypes = [vp, i32, u64, u64, C.POINTER(u64), C.POINTER(u64)]
    L.sq_kernel_timing.restype = i32
    L.sq_kernel_timing.argtypes = [vp, i32]
    L.sq_kernel_time.restype = i32
    L.sq_kernel_time.argtypes = [vp, pd, C.POINTER(i64)]
    L.sq_lcg_jump.restype = u64
    L.sq_lcg_jump.argtypes = [u64, u64, u64]
    L.sq_controller_set.restype = i32
    L.sq_controller_set.argtypes = [vp, dbl, i64, i32]
    L.sq_controller_get.restype = i32
    L.sq_controller_get.argtypes = [vp, pd, C.POINTER(i64), C.POINTER(i32)]
    L.sq_frames.restype = i32
    L.sq_frames.argtypes = [vp, i32, i32, C.POINTER(SqFrameRec), pd]
    L.sq_compat_get_state.restype = i32
    L.sq_compat_get_state.argtypes = [vp, C.POINTER(SqCompatState)]
    L.sq_compat_set_state.restype = i32
    L.sq_compat_set_state.argtypes = [vp, C.POINTER(SqCompatState)]
    L.sq_session_open.restype = i32
    L.sq_session_open.argtypes = [C.POINTER(vp), C.c_char_p, i32, i32]
    L.sq_session_barrier.restype = i32
    L.sq_session_barrier.argtypes = [vp]
    L.sq_session_allgather_u64.restype = i32
    L.sq_session_allgather_u64.argtypes = [vp, C.POINTER(u64), i32, C.POINTER(u64)]
    L.sq_session_allgather_f64.restype = i32
    L.sq_session_allgather_f64.argtypes = [vp, pd, i32, pd]
    L.sq_session_abort.restype = None
    L.sq_session_abort.argtypes = [vp]
    L.sq_session_rank.restype = i32
    L.sq_session_rank.argtypes = [vp]
    L.sq_session_size.restype = i32
    L.sq_session_size.argtypes = [vp]
    L.sq_session_close.restype = None
    L.sq_session_close.argtypes = [vp]
    L.sq_slab_join.restype = i32
    L.sq_slab_join.argtypes = [vp, vp]
    L.sq_slab_stats.restype = i32
    L.sq_slab_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.sq_rng_resolve.restype = i32
    L.sq_rng_resolve.argtypes = [u64, C.POINTER(SqRngEntry), i32, u64, C.POINTER(SqRngEntry),
                                 C.POINTER(i32), C.POINTER(i32)]
    _lib = L
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _check(L, rc, what):
    if rc != 0:
        raise SqError(rc, what, (L.sq_strerror(rc) or b"").decode() + "; " + (L.sq_last_cuda_error() or b"").decode())


class Session:
    """Rendezvous of the ranks of one slab ring on one box (POSIX shared memory; no GPU needed)."""

    def __init__(self, name: str, rank: int, nranks: int):
        self.L = load()
        self._h = C.c_void_p()
        _check(self.L, self.L.sq_session_open(C.byref(self._h), name.encode(), rank, nranks), "sq_session_open")
        self.rank, self.nranks = rank, nranks

    def barrier(self):
        _check(self.L, self.L.sq_session_barrier(self._h), "sq_session_barrier")

    def allgather_u64(self, words) -> np.ndarray:
        a = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1)
        out = np.zeros((self.nranks, a.size), dtype=np.uint64)
        p64 = C.POINTER(C.c_uint64)
        _check(self.L, self.L.sq_session_allgather_u64(self._h, a.ctypes.data_as(p64), a.size, out.ctypes.data_as(p64)),
               "sq_session_allgather_u64")
        return out

    def allgather_f64(self, vals) -> np.ndarray:
        a = np.ascontiguousarray(vals, dtype=np.float64).reshape(-1)
        out = np.zeros((self.nranks, a.size))
        _check(self.L, self.L.sq_session_allgather_f64(self._h, _dp(a), a.size, _dp(out)), "sq_session_allgather_f64")
        return out

    def abort(self):
        if self._h.value:
            self.L.sq_session_abort(self._h)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.L.sq_session_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rng_resolve(step_seed: int, entries: list, gid: int):
    """sq_rng_resolve: literal replay of the draw at gid -> (entry dict, ndraws, plus)."""
    L = load()
    arr = (SqRngEntry * max(1, len(entries)))()
    for k, e in enumerate(entries):
        for f, _ in SqRngEntry._fields_:
            setattr(arr[k], f, int(e[f]))
    out, nd, pl = SqRngEntry(), C.c_int(), C.c_int()
    _check(L, L.sq_rng_resolve(int(step_seed) & (2**64 - 1), arr, len(entries), int(gid), C.byref(out), C.byref(nd),
                               C.byref(pl)), "sq_rng_resolve")
    return {f: int(getattr(out, f)) for f, _ in SqRngEntry._fields_}, nd.value, pl.value


class Context:
    """One sq_ctx.  kernel='compat1d' mirrors the reference's 1-D run; kernel='lattice'
    is the d-dimensional generalisation."""

    def __init__(self, dims, kernel="lattice", real="f32", math="accurate", potential=0,
                 spacing=1.0, noise_c=1.0, m2=0.0, lam=0.0, device=0, nchains=1, seed=1242608872,
                 f0=None, x0=None, xx0_0=None, omega0=0.0, slab=(0, 0), flags=0):
        self.L = load()
        dims = [int(d) for d in (dims if hasattr(dims, "__len__") else [dims])]
        p = SqParams()
        p.struct_size = C.sizeof(SqParams)
        p.kernel = SQ_KERNEL_COMPAT1D if kernel == "compat1d" else SQ_KERNEL_LATTICE
        p.real = {"f32": SQ_REAL_F32, "f64": SQ_REAL_F64}[real]
        if kernel == "compat1d":
            p.real = SQ_REAL_F64
        p.math = {"accurate": SQ_MATH_ACCURATE, "fast": SQ_MATH_FAST}[math]
        p.potential = potential
        p.ndim = len(dims)
        for k, d in enumerate(dims):
            p.dims[k] = d
        p.spacing, p.noise_c, p.m2, p.lam = spacing, noise_c, m2, lam
        p.device, p.nchains = device, nchains
        p.slab_t0, p.slab_nt = slab
        p.flags, p.reserved = flags, 0
        self.params = p
        self.dims = dims
        self.kernel = kernel
        self.dtype = np.float32 if p.real == SQ_REAL_F32 else np.float64
        self.volume = int(np.prod(dims))
        self.nt = (slab[1] or dims[-1]) if kernel != "compat1d" else dims[0]
        self.vlocal = self.volume // dims[-1] * self.nt if kernel != "compat1d" else dims[0]
        self.runs = 0
        self._h = C.c_void_p()
        conv = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        self._keep = [conv(f0), conv(x0), conv(xx0_0)]
        rc = self.L.sq_init(C.byref(self._h), C.byref(p), _dp(self._keep[0]), _dp(self._keep[1]),
                            _dp(self._keep[2]), float(omega0), int(seed) & (2**64 - 1))
        self._check(rc, "sq_init")

    # -- plumbing
    def _check(self, rc, what):
        if rc != 0:
            raise SqError(rc, what, (self.L.sq_strerror(rc) or b"").decode() + "; " +
                          (self.L.sq_last_cuda_error() or b"").decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.L.sq_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def join(self, session: "Session"):
        """Collective: make this context one slab of the ring behind `session` (sq_slab_join)."""
        self._check(self.L.sq_slab_join(self._h, session._h), "sq_slab_join")
        self._session = session  # must outlive the context

    def slab_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self.L.sq_slab_stats(self._h, C.byref(a), C.byref(b)), "sq_slab_stats")
        return {"finder_scans": a.value, "agree_rounds": b.value}

    @property
    def stream(self) -> int:
        return int(self.L.sq_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.L.sq_launch_count(self._h))

    # -- the frame loop's vocabulary
    def step(self, dtau: float, nsteps: int, runs0: int | None = None) -> bool:
        """One frame of `nsteps` (= Loops) tau-steps; returns stable (tauhost.c:481-506)."""
        r0 = self.runs if runs0 is None else runs0
        st = C.c_int(1)
        self._check(self.L.sq_step(self._h, dtau, nsteps, r0, C.byref(st)), "sq_step")
        if st.value == 1:
            self.runs = r0 + nsteps
        return st.value == 1

    def step_async(self, dtau: float, nsteps: int, runs0: int | None = None):
        r0 = self.runs if runs0 is None else runs0
        self._pending = (r0, nsteps)
        self._check(self.L.sq_step_async(self._h, dtau, nsteps, r0), "sq_step_async")

    def sync(self) -> bool:
        st = C.c_int(1)
        self._check(self.L.sq_sync(self._h, C.byref(st)), "sq_sync")
        if st.value == 1 and getattr(self, "_pending", None):
            self.runs = self._pending[0] + self._pending[1]
        self._pending = None
        return st.value == 1

    def measure(self) -> dict:
        n = self.vlocal if self.kernel == "compat1d" else self.nt
        o = SqObs()
        o.struct_size = C.sizeof(SqObs)
        out = {}
        if self.kernel == "compat1d":
            for k in ("f", "x", "xx0"):
                out[k] = np.zeros(n)
                setattr(o, k, _dp(out[k]))
        out["corr"] = np.zeros(n)
        o.corr = _dp(out["corr"])
        if self.kernel != "compat1d":
            out["slice_x"], out["slice_xx0"] = np.zeros(n), np.zeros(n)
            o.slice_x, o.slice_xx0 = _dp(out["slice_x"]), _dp(out["slice_xx0"])
        self._check(self.L.sq_measure(self._h, C.byref(o)), "sq_measure")
        for k in ("omega", "seed", "lrgEl", "stable", "lrgVl", "runs", "mean_phi", "mean_phi2",
                  "nclamped", "nevents", "steps_done"):
            out[k] = getattr(o, k)
        return out

    # -- frame controller on the device (compat1d, sq.h: sq_controller_* / sq_frames)
    def controller_set(self, dtau: float, runs: int = 0, stab_cnt: int = 0):
        self._check(self.L.sq_controller_set(self._h, dtau, runs, stab_cnt), "sq_controller_set")

    def controller_get(self):
        d, r, s_ = C.c_double(), C.c_int64(), C.c_int32()
        self._check(self.L.sq_controller_get(self._h, C.byref(d), C.byref(r), C.byref(s_)), "sq_controller_get")
        return d.value, r.value, s_.value

    def frames(self, nframes: int, nsteps: int):
        """nframes frames back to back under the device-side controller -> (records, xavg[nframes][N])."""
        recs = (SqFrameRec * max(1, nframes))()
        xavg = np.full((max(1, nframes), self.vlocal), np.nan)
        self._check(self.L.sq_frames(self._h, nframes, nsteps, recs, _dp(xavg)), "sq_frames")
        out = [(recs[k].dtau, recs[k].stable, recs[k].steps) for k in range(nframes)]
        self.runs = self.controller_get()[1]
        return out, xavg[:nframes]

    def compat_state(self) -> dict:
        """sq_compat_get_state: what the end file lacks for a bit-exact resume."""
        st = SqCompatState()
        st.struct_size = C.sizeof(SqCompatState)
        self._check(self.L.sq_compat_get_state(self._h, C.byref(st)), "sq_compat_get_state")
        return {k: getattr(st, k) for k, _ in SqCompatState._fields_ if k != "struct_size"}

    def set_compat_state(self, seed, lrgEl, lrgVl, omega, newf_lrgEl):
        st = SqCompatState()
        st.struct_size = C.sizeof(SqCompatState)
        st.seed, st.lrgEl, st.lrgVl, st.omega, st.newf_lrgEl = int(seed), int(lrgEl), lrgVl, omega, newf_lrgEl
        self._check(self.L.sq_compat_set_state(self._h, C.byref(st)), "sq_compat_set_state")

    def kernel_timing(self, enable: bool):
        self._check(self.L.sq_kernel_timing(self._h, int(enable)), "sq_kernel_timing")

    def kernel_time(self):
        ms, n = C.c_double(), C.c_int64()
        self._check(self.L.sq_kernel_time(self._h, C.byref(ms), C.byref(n)), "sq_kernel_time")
        return ms.value, n.value

    # -- lattice extras
    def upload(self, field, chain=0):
        a = np.ascontiguousarray(field).reshape(-1)
        real = SQ_REAL_F32 if a.dtype == np.float32 else SQ_REAL_F64
        if real == SQ_REAL_F64:
            a = a.astype(np.float64, copy=False)
        assert a.size == self.vlocal
        self._check(self.L.sq_upload_field(self._h, chain, a.ctypes.data, real), "sq_upload_field")

    def download(self, chain=0, dtype=None) -> np.ndarray:
        dt = np.dtype(dtype or self.dtype)
        a = np.empty(self.vlocal, dtype=dt)
        self._check(self.L.sq_download_field(self._h, chain, a.ctypes.data,
                                             SQ_REAL_F32 if dt == np.float32 else SQ_REAL_F64),
                    "sq_download_field")
        return a

    def set_chain(self, chain, seed, m2=0.0, lam=0.0):
        self._check(self.L.sq_set_chain(self._h, chain, int(seed) & (2**64 - 1), m2, lam), "sq_set_chain")

    def measure_chains(self):
        nc = self.params.nchains
        a, b, s = np.zeros(nc), np.zeros(nc), np.zeros(nc, dtype=np.uint64)
        self._check(self.L.sq_measure_chains(self._h, _dp(a), _dp(b),
                                             s.ctypes.data_as(C.POINTER(C.c_uint64))), "sq_measure_chains")
        return a, b, s

    def debug_draws(self, gid0, n, chain=0):
        t1 = np.zeros(n, dtype=np.uint64)
        t2 = np.zeros(n, dtype=np.uint64)
        p64 = C.POINTER(C.c_uint64)
        self._check(self.L.sq_debug_draws(self._h, chain, gid0, n, t1.ctypes.data_as(p64),
                                          t2.ctypes.data_as(p64)), "sq_debug_draws")
        return t1, t2

    def frame_host(self, host_in_ptr, host_out_ptr, dtau, nsteps, runs0=None, measure=False):
        """End-to-end frame through host buffers (raw pointers, ideally pinned)."""
        r0 = self.runs if runs0 is None else runs0
        st = C.c_int(1)
        o = None
        if measure:
            o = SqObs()
            o.struct_size = C.sizeof(SqObs)
        real = SQ_REAL_F32 if self.dtype == np.float32 else SQ_REAL_F64
        self._check(self.L.sq_frame_host(self._h, host_in_ptr, host_out_ptr, real, dtau, nsteps, r0,
                                         C.byref(o) if o is not None else None, C.byref(st)),
                    "sq_frame_host")
        if st.value == 1:
            self.runs = r0 + nsteps
        return st.value == 1, o
