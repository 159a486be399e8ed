"""ctypes loader for the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  See oracle/sq_oracle.h for scope and citations
(/root/reference/tau_kernel.cl:25-284, /root/reference/tauhost.c:29-621).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsq_oracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libtau_ref.so")

RNG_CHAIN, RNG_SHARED = 0, 1
FIELD_JACOBI, FIELD_INPLACE = 0, 1
F32, F64 = 0, 1


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIB_PATH) or (
        os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "sq_oracle.c"))
    ) or (os.path.isdir("/root/reference") and not os.path.exists(REF_PATH)):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)


class Draw(C.Structure):
    _fields_ = [("t1", C.c_uint64), ("t2", C.c_uint64), ("seed_after", C.c_uint64),
                ("ndraws", C.c_int), ("plus_branch", C.c_int)]


class State(C.Structure):
    _fields_ = [("N", C.c_int), ("deltaT", C.c_double), ("deltaTau", C.c_double),
                ("C", C.c_double), ("potential", C.c_int),
                ("f", C.POINTER(C.c_double)), ("x", C.POINTER(C.c_double)),
                ("xx0", C.POINTER(C.c_double)), ("newf", C.POINTER(C.c_double)),
                ("newx", C.POINTER(C.c_double)), ("newxx0", C.POINTER(C.c_double)),
                ("omega", C.c_double), ("rand1", C.c_uint64), ("stable", C.c_int),
                ("lrgEl", C.c_int), ("lrgVl", C.c_double), ("runs", C.c_int)]


class Lattice(C.Structure):
    _fields_ = [("ndim", C.c_int), ("dims", C.c_int64 * 4), ("real", C.c_int),
                ("potential", C.c_int), ("a", C.c_double), ("C", C.c_double),
                ("m2", C.c_double), ("lam", C.c_double),
                ("phi", C.c_void_p), ("phi_new", C.c_void_p), ("seed", C.c_uint64),
                ("runs", C.c_int64), ("slice_x", C.POINTER(C.c_double)),
                ("slice_xx0", C.POINTER(C.c_double)), ("slice_sum", C.POINTER(C.c_double)),
                ("sum_phi", C.c_double), ("sum_phi2", C.c_double),
                ("nclamped", C.c_int64), ("nevents", C.c_uint64)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.sqo_random.restype = C.c_double
        L.sqo_random.argtypes = [C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(Draw)]
        for name in ("sqo_clas",):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double, C.c_double, C.c_int]
        L.sqo_ddPot.restype = C.c_double
        L.sqo_ddPot.argtypes = [C.c_double, C.c_int]
        L.sqo_intConst.restype = C.c_double
        L.sqo_intConst.argtypes = [C.c_int]
        L.sqo_time_dev.restype = C.c_int
        L.sqo_time_dev.argtypes = [C.POINTER(State), C.c_int, C.c_int, C.c_int]
        L.sqo_set_trace.argtypes = [C.POINTER(Draw), C.c_int]
        L.sqo_host_init.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int,
                                    C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_uint64)]
        L.sqo_jump.restype = C.c_uint64
        L.sqo_jump.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.sqo_lattice_step.argtypes = [C.POINTER(Lattice), C.c_double]
        L.sqo_lattice_step_omp.argtypes = [C.POINTER(Lattice), C.c_double]
        L.sqo_set_threads.argtypes = [C.c_int]
        L.sqo_set_threads.restype = C.c_int
        L.sqo_lattice_draws.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.sqo_write_endfile.restype = C.c_int
        L.sqo_write_endfile.argtypes = [C.c_char_p, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 4 + [
            C.c_double, C.c_int, C.c_double]
        L.sqo_read_startfile.restype = C.c_int
        L.sqo_read_startfile.argtypes = [C.c_char_p, C.c_int, C.c_double] + [C.POINTER(C.c_double)] * 4 + [
            C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.sqo_print_frame_path.restype = C.c_int
        L.sqo_print_frame_path.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double), C.c_double, C.c_int, C.c_int]
        L.sqo_tauhost_main_path.restype = C.c_int
        L.sqo_tauhost_main_path.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def tauhost_main(args, outpath, rng=RNG_CHAIN, field=FIELD_JACOBI) -> int:
    """The whole reference program on the CPU (tauhost.c:29-621 over the oracle kernel).
    args: the 13 positional arguments as strings.  stdout stream goes to outpath."""
    argv = [b"tauhost.o"] + [str(a).encode() for a in args]
    arr = (C.c_char_p * len(argv))(*argv)
    return lib().sqo_tauhost_main_path(len(argv), arr, os.fsencode(outpath), rng, field)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def random(seed: int, gid: int):
    """tau_kernel.cl:269-284 -> (value, Draw)."""
    s = C.c_uint64(seed & (2**64 - 1))
    rec = Draw()
    v = lib().sqo_random(C.byref(s), gid, C.byref(rec))
    return v, rec


def host_init(N: int, deltat: float, deltatau: float, cold: bool = True):
    """tauhost.c:84-102,185 -> (f, omega, rand1)."""
    f = np.zeros(N, dtype=np.float64)
    om = C.c_double()
    r1 = C.c_uint64()
    lib().sqo_host_init(N, deltat, deltatau, int(cold), _dp(f), C.byref(om), C.byref(r1))
    return f, om.value, r1.value


class Compat1D:
    """The reference's 1-D kernel state + launches (oracle semantics selectable)."""

    def __init__(self, N, deltat, deltatau, potential, c, f, omega, seed, x=None, xx0=None,
                 runs=0, rng=RNG_CHAIN, field=FIELD_JACOBI):
        self.N = N
        self.rng, self.field = rng, field
        self.f = np.array(f, dtype=np.float64).copy()
        self.x = np.zeros(N) if x is None else np.array(x, dtype=np.float64).copy()
        self.xx0 = np.zeros(N) if xx0 is None else np.array(xx0, dtype=np.float64).copy()
        self.newf, self.newx, self.newxx0 = self.f.copy(), self.x.copy(), self.xx0.copy()
        s = State()
        s.N, s.deltaT, s.deltaTau, s.C, s.potential = N, deltat, deltatau, c, potential
        s.f, s.x, s.xx0 = _dp(self.f), _dp(self.x), _dp(self.xx0)
        s.newf, s.newx, s.newxx0 = _dp(self.newf), _dp(self.newx), _dp(self.newxx0)
        s.omega, s.rand1, s.stable, s.lrgEl, s.lrgVl, s.runs = omega, seed, 1, 0, 0.0, runs
        self.s = s

    def launch(self, loops: int) -> int:
        return lib().sqo_time_dev(C.byref(self.s), loops, self.rng, self.field)

    def frame(self, loops: int) -> bool:
        """One reference frame incl. the host's commit / rollback (tauhost.c:504-554),
        with the dtau controller left to the caller.  Returns stable."""
        f0, x0, xx00, om0 = self.f.copy(), self.x.copy(), self.xx0.copy(), self.s.omega
        self.launch(loops)
        ok = self.s.stable == 1
        if ok:
            self.f[:], self.x[:], self.xx0[:] = self.newf, self.newx, self.newxx0
            self.s.runs += loops
        else:
            self.f[:], self.x[:], self.xx0[:] = f0, x0, xx00
            self.s.omega = om0
            self.s.stable = 1
        return ok


def set_threads(n: int = 0) -> int:
    """OpenMP threads of LatticeOracle.step(omp=True); n <= 0 only queries.  bench.py's timing legs call
    it with the process's CPU affinity because torchrun exports OMP_NUM_THREADS=1 to its workers."""
    return int(lib().sqo_set_threads(int(n)))


class LatticeOracle:
    """d-dim generalisation; serial definition or OpenMP variant."""

    def __init__(self, dims, real=F32, potential=0, a=1.0, c=1.0, m2=0.0, lam=0.0,
                 seed=1242608872, phi0=None):
        dims = list(dims)
        self.dims = dims
        self.V = int(np.prod(dims))
        self.dtype = np.float32 if real == F32 else np.float64
        self.phi = np.zeros(self.V, dtype=self.dtype) if phi0 is None else \
            np.ascontiguousarray(np.asarray(phi0, dtype=self.dtype).reshape(-1)).copy()
        self.phi_new = np.zeros(self.V, dtype=self.dtype)
        Lt = dims[-1]
        self.slice_x = np.zeros(Lt)
        self.slice_xx0 = np.zeros(Lt)
        self.slice_sum = np.zeros(Lt)
        L = Lattice()
        L.ndim = len(dims)
        for k, d in enumerate(dims):
            L.dims[k] = d
        L.real, L.potential, L.a, L.C, L.m2, L.lam = real, potential, a, c, m2, lam
        L.phi, L.phi_new = self.phi.ctypes.data, self.phi_new.ctypes.data
        L.seed, L.runs = seed, 0
        L.slice_x, L.slice_xx0, L.slice_sum = _dp(self.slice_x), _dp(self.slice_xx0), _dp(self.slice_sum)
        self.L = L

    def step(self, dtau: float, n: int = 1, omp: bool = False):
        fn = lib().sqo_lattice_step_omp if omp else lib().sqo_lattice_step
        for _ in range(n):
            fn(C.byref(self.L), dtau)

    @property
    def field(self) -> np.ndarray:
        cur = self.phi if self.L.phi == self.phi.ctypes.data else self.phi_new
        return cur

    @property
    def seed(self) -> int:
        return self.L.seed


def lattice_draws(seed: int, V: int):
    t1 = np.zeros(V, dtype=np.uint64)
    t2 = np.zeros(V, dtype=np.uint64)
    after = C.c_uint64()
    lib().sqo_lattice_draws(seed, V, t1.ctypes.data_as(C.POINTER(C.c_uint64)),
                            t2.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(after))
    return t1, t2, after.value


# ---------------------------------------------------------------- oracle/_ref --
_ref = None


def ref_available() -> bool:
    build()
    return os.path.exists(REF_PATH)


def ref() -> C.CDLL:
    """The reference kernel compiled from /root/reference/tau_kernel.cl (oracle/_ref)."""
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(REF_PATH)
        R.sq_ref_random.restype = C.c_double
        R.sq_ref_random.argtypes = [C.POINTER(C.c_ulong), C.c_int]
        R.clas.restype = C.c_double
        R.clas.argtypes = [C.c_double, C.c_double, C.c_int]
        R.ddPot.restype = C.c_double
        R.ddPot.argtypes = [C.c_double, C.c_int]
        R.intConst.restype = C.c_double
        R.intConst.argtypes = [C.c_int]
        _ref = R
    return _ref


class RefKernel:
    """Reference `time_dev` (real source) under the serial work-item schedule."""

    def __init__(self, N, deltat, deltatau, potential, c, f, omega, seed, x=None, xx0=None, runs=0):
        self.N = N
        self.f = np.array(f, dtype=np.float64).copy()
        self.x = np.zeros(N) if x is None else np.array(x, dtype=np.float64).copy()
        self.xx0 = np.zeros(N) if xx0 is None else np.array(xx0, dtype=np.float64).copy()
        self.newf, self.newx, self.newxx0 = self.f.copy(), self.x.copy(), self.xx0.copy()
        self.omega = C.c_double(omega)
        self.rand1 = C.c_ulong(seed)
        self.stable = C.c_int(1)
        self.deltaTau = C.c_double(deltatau)
        self.lrgEl = C.c_int(0)
        self.lrgVl = C.c_double(0.0)
        self.initRun = C.c_int(1)
        self.LIST_SIZE = C.c_int(N)
        self.deltaT = C.c_double(deltat)
        self.runs = C.c_int(runs)
        self.potential = C.c_int(potential)
        self.C = C.c_double(c)

    def launch(self, loops: int):
        L = C.c_int(loops)
        ref().sq_ref_launch(_dp(self.f), _dp(self.x), _dp(self.xx0), _dp(self.newf), _dp(self.newx),
                            _dp(self.newxx0), C.byref(self.omega), C.byref(self.rand1),
                            C.byref(self.stable), C.byref(self.deltaTau), C.byref(self.lrgEl),
                            C.byref(self.lrgVl), C.byref(self.initRun), C.byref(self.LIST_SIZE),
                            C.byref(self.deltaT), C.byref(self.runs), C.byref(self.potential),
                            C.byref(self.C), C.byref(L))

    def steps_canonical(self, n: int):
        """n tau-steps as n launches with Loops=1 + the host's f=newf hand-over
        (tauhost.c:508-554): canonical {chain RNG, Jacobi} semantics."""
        for _ in range(n):
            self.launch(1)
            if self.stable.value != 1:
                return False
            self.f[:], self.x[:], self.xx0[:] = self.newf, self.newx, self.newxx0
            self.runs.value += 1
        return True
