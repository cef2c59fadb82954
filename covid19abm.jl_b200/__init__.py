"""covid19abm.jl_b200 — host-side mirror of covid19abm.jl's public API over libcabm_b200.so.

Same names and argument meaning as the reference (/root/reference/src/covid19abm.jl):
`ModelParameters` (:31-55), `HEALTH` (:5), `main(ip)` (:104-142) returning the popsize x modeltime Int16
matrix, `runsim(simnum, ip)` (:77-101). All computation happens in hand-written sm_100a kernels behind the
C ABI of include/cabm.h; there is no CPU path: without the built library or without a GPU the calls raise.

The directory name contains a dot, so import it through the repo-root shim:  `import covid19abm_b200 as cv`.
"""
import ctypes as C
import enum
import os
from dataclasses import dataclass, fields

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CABM_LIB", os.path.join(_HERE, "libcabm_b200.so"))     # CABM_LIB: e.g. the diagnostics build

PROVINCES = ("alberta", "bc", "canada", "manitoba", "newbruns", "newfdland", "nwterrito", "novasco",
             "nunavut", "ontario", "pei", "quebec", "saskat", "yukon", "newyork")     # order of :319-333


class HEALTH(enum.IntEnum):          # src/covid19abm.jl:5
    SUS = 0; LAT = 1; PRE = 2; ASYMP = 3; MILD = 4; MISO = 5; INF = 6; IISO = 7; HOS = 8; ICU = 9; REC = 10; DED = 11; UNDEF = 12


SUS, LAT, PRE, ASYMP, MILD, MISO, INF, IISO, HOS, ICU, REC, DED, UNDEF = (int(h) for h in HEALTH)
ISOVIA = ("null", "qu", "pi", "mi", "ct")            # :21
MODE_EXACT, MODE_NATIVE = 0, 1
STORE_HMATRIX = 1
FIELDS = {"health": 0, "swap": 1, "sickfrom": 2, "nextday_meetcnt": 3, "ag": 4, "tis": 5, "exp": 6, "dur": 7,
          "doi": 8, "iso": 9, "isovia": 10, "tracing": 11, "tracestart": 12, "traceend": 13, "tracedxp": 14,
          "ncontacts": 15}
ERR_BITS = {1: "swap expired, but no swap set.", 2: "No susceptibles in the population to change status",
            4: "a traced person should always be isolated until the trace is killed", 8: "empty age group sampled",
            16: "contact log capacity exceeded", 32: "sickfrom not set right", 64: "kernel implementation limit exceeded"}


TAPE_DTYPE = np.dtype([("seed", "<u8"), ("ctr", "<u4", (4,)), ("w", "<u4", (4,))])       # struct cabm_tape_entry (40 bytes)
STREAMS = {"INIT": 1, "COUNT": 2, "SEED": 3, "TRANS": 4, "CTWIN": 5, "CTCON": 6, "CT3": 7, "DT": 8}   # draw-contract streams


class CabmError(RuntimeError):
    """Stands in for Julia's ErrorException: raised with the reference's error("...") message where one exists."""


class CabmParams(C.Structure):        # include/cabm.h: cabm_params
    _fields_ = [("beta", C.c_double), ("frelasymp", C.c_double), ("seasonal", C.c_int32), ("popsize", C.c_int32),
                ("prov_id", C.c_int32), ("calibration", C.c_int32), ("modeltime", C.c_int32),
                ("initialinf", C.c_int32), ("initialhi", C.c_int32), ("tau_mild", C.c_int32),
                ("asymp_props", C.c_double * 5), ("mild_props", C.c_double * 5), ("fmild", C.c_double),
                ("fsevere", C.c_double), ("fpreiso", C.c_double), ("quar_grp_frac", C.c_double),
                ("tpreiso", C.c_int32), ("quar_grp", C.c_int32), ("ctstrat", C.c_int32), ("cdaysback", C.c_int32),
                ("strat3qdays", C.c_int32), ("fctcapture", C.c_float), ("fcontactst", C.c_float), ("_pad", C.c_int32)]


@dataclass(init=False)
class ModelParameters:
    """@with_kw mutable struct ModelParameters (src/covid19abm.jl:31-55): same fields, same defaults.

    `β` and `τmild` are valid Python identifiers and are the canonical names; `beta`/`tau_mild` are accepted as
    keyword aliases, as are the README-generation names `eldq` (quar_grp_frac) and `eldqag` (quar_grp).
    `fpre` and `cidtime` (README only, no behaviour in the source) are accepted and ignored.
    """
    β: float = 0.0
    frelasymp: float = 0.11
    seasonal: bool = False
    popsize: int = 10000
    prov: str = "newyork"
    calibration: bool = False
    modeltime: int = 300
    initialinf: int = 1
    initialhi: int = 0
    asymp_props: tuple = (0.25, 0.25, 0.14, 0.07, 0.07)
    mild_props: tuple = (0.95, 0.9, 0.85, 0.6, 0.2)
    τmild: int = 0
    fmild: float = 0.0
    fsevere: float = 0.0
    tpreiso: int = 0
    fpreiso: float = 0.0
    quar_grp: int = 5
    quar_grp_frac: float = 0.0
    ctstrat: int = 0
    fctcapture: float = 0.0
    fcontactst: float = 0.0
    cdaysback: int = 0
    strat3qdays: int = 0

    _ALIASES = {"beta": "β", "tau_mild": "τmild", "eldq": "quar_grp_frac", "eldqag": "quar_grp"}
    _IGNORED = ("fpre", "cidtime")

    def __init__(self, **kw):
        for f in fields(self):
            setattr(self, f.name, f.default)
        for k, v in kw.items():
            if k in self._IGNORED:
                continue
            k = self._ALIASES.get(k, k)
            if k not in {f.name for f in fields(self)}:
                raise TypeError(f"ModelParameters has no field {k}")
            setattr(self, k, v)

    @property
    def beta(self):
        return self.β

    @property
    def tau_mild(self):
        return self.τmild

    def to_c(self):
        p = CabmParams()
        prov = str(self.prov).lstrip(":")
        p.beta, p.frelasymp, p.seasonal, p.popsize = float(self.β), float(self.frelasymp), int(bool(self.seasonal)), int(self.popsize)
        p.prov_id = PROVINCES.index(prov) if prov in PROVINCES else -1
        p.calibration, p.modeltime = int(bool(self.calibration)), int(self.modeltime)
        p.initialinf, p.initialhi, p.tau_mild = int(self.initialinf), int(self.initialhi), int(self.τmild)
        p.asymp_props = (C.c_double * 5)(*[float(x) for x in self.asymp_props])
        p.mild_props = (C.c_double * 5)(*[float(x) for x in self.mild_props])
        p.fmild, p.fsevere, p.fpreiso, p.quar_grp_frac = float(self.fmild), float(self.fsevere), float(self.fpreiso), float(self.quar_grp_frac)
        p.tpreiso, p.quar_grp, p.ctstrat = int(self.tpreiso), int(self.quar_grp), int(self.ctstrat)
        p.cdaysback, p.strat3qdays = int(self.cdaysback), int(self.strat3qdays)
        p.fctcapture = float(np.float16(self.fctcapture))      # Float16 fields :51-52
        p.fcontactst = float(np.float16(self.fcontactst))
        return p


_lib = None


def lib():
    """The C-ABI library. Fails loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CabmError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.cabm_last_error.restype = C.c_char_p; L.cabm_last_error.argtypes = [vp]
    L.cabm_version.restype = C.c_char_p
    L.cabm_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32]
    L.cabm_destroy.argtypes = [vp]; L.cabm_destroy.restype = None
    L.cabm_configure.argtypes = [vp, vp, i32, vp, i32, vp, i32]
    L.cabm_run.argtypes = [vp]; L.cabm_sync.argtypes = [vp]
    L.cabm_set_strict.argtypes = [vp, i32]
    L.cabm_set_tape.argtypes = [vp, vp, C.c_uint64]
    L.cabm_last_run_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.cabm_set_host_curve_rows.argtypes = [vp, vp]
    L.cabm_mark.argtypes = [vp]; L.cabm_ms_since_mark.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.cabm_launch_count.argtypes = [vp]; L.cabm_launch_count.restype = i64
    L.cabm_set_profiling.argtypes = [vp, i32]
    L.cabm_set_path.argtypes = [vp, i32]
    L.cabm_get_debug.argtypes = [vp, vp]
    L.cabm_get_last_path.argtypes = [vp]
    L.cabm_get_profile.argtypes = [vp, vp, vp]
    L.cabm_step_initialize.argtypes = [vp]
    L.cabm_step_insert_infected.argtypes = [vp, i32, i32, i32, i32]
    L.cabm_step_dyntrans.argtypes = [vp]; L.cabm_step_time_update.argtypes = [vp]
    L.cabm_get_day.argtypes = [vp]
    L.cabm_get_dyntrans_counters.argtypes = [vp, vp, vp]
    L.cabm_get_hmatrix.argtypes = [vp, i32, vp]; L.cabm_get_hmatrix_all.argtypes = [vp, vp]; L.cabm_get_hmatrix_all_u8.argtypes = [vp, vp]
    L.cabm_get_curves.argtypes = [vp, vp, vp]
    L.cabm_get_curves_i16.argtypes = [vp, vp, vp]
    L.cabm_get_mean_curves.argtypes = [vp, vp, vp]
    L.cabm_set_host_curves_i16.argtypes = [vp, vp, vp]
    L.cabm_get_scalars.argtypes = [vp, vp, vp, vp, vp]
    L.cabm_get_field.argtypes = [vp, i32, i32, vp]
    L.cabm_set_field.argtypes = [vp, i32, i32, i32, C.c_int32]
    L.cabm_reduce_hmatrix.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    L.cabm_host_alloc.argtypes = [C.c_uint64]; L.cabm_host_alloc.restype = vp
    L.cabm_host_free.argtypes = [vp]; L.cabm_host_free.restype = None
    L.cabm_get_table.argtypes = [C.c_char_p, C.POINTER(C.c_int32), vp, i32]
    L.cabm_get_gpw.argtypes = [i32, i32, vp]; L.cabm_get_gpw.restype = None
    _lib = L
    return L


def get_table(name):
    """(base, thresholds) of one draw-contract table; host-only."""
    buf = np.zeros(512, dtype=np.uint32)
    base = C.c_int32(0)
    n = lib().cabm_get_table(name.encode(), C.byref(base), buf.ctypes.data, 512)
    if n < 0:
        raise KeyError(name)
    return base.value, buf[:n].copy()


def get_gpw(ag, cnts):
    out = np.zeros(5, dtype=np.int32)
    lib().cabm_get_gpw(ag, cnts, out.ctypes.data)
    return out


class PinnedArray:
    """numpy view over cudaMallocHost memory (cabm_host_alloc) so device->host copies run at full PCIe rate."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib().cabm_host_alloc(self.nbytes)
        if not self.ptr:
            raise CabmError("cabm_host_alloc failed")
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().cabm_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        self.free()


class Population:
    """A device-resident ensemble: `humans`, `p`, `BETAS`, `ct_data` of the reference (:69-73) x n replicates."""

    def __init__(self, max_replicates, popsize, modeltime, store_hmatrix=False, device=0):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.cabm_create(C.byref(self.h), device, max_replicates, popsize, modeltime, STORE_HMATRIX if store_hmatrix else 0)
        if rc:
            self.h = None
            raise CabmError(self.L.cabm_last_error(None).decode())
        self.R, self.N, self.T, self.store_hmatrix = max_replicates, popsize, modeltime, store_hmatrix
        self.n_rep = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.cabm_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc:
            raise CabmError(self.L.cabm_last_error(self.h).decode())

    def configure(self, psets, seeds, pset_of_rep=None, mode=MODE_EXACT):
        """reset_params :146-168 for every replicate. psets: ModelParameters or list of them."""
        if isinstance(psets, ModelParameters):
            psets = [psets]
        arr = (CabmParams * len(psets))(*[p.to_c() for p in psets])
        sd = np.ascontiguousarray(seeds, dtype=np.uint64)
        por = None if pset_of_rep is None else np.ascontiguousarray(pset_of_rep, dtype=np.int32)
        self._chk(self.L.cabm_configure(self.h, arr, len(psets), None if por is None else por.ctypes.data,
                                        len(sd), sd.ctypes.data, mode))
        self.n_rep = len(sd)
        return self

    def run(self, sync=True):
        self._chk(self.L.cabm_run(self.h))
        if sync:
            self._chk(self.L.cabm_sync(self.h))

    def sync(self):
        self._chk(self.L.cabm_sync(self.h))

    def set_strict(self, on=True):
        """strict (default): after a run, sync() and the result getters raise CabmError with the reference's error("...") text
        when any replicate hit one of the in-model errors (:308,:371-372,:418,:719) or an implementation limit; off: the caller
        inspects errors() itself."""
        self._chk(self.L.cabm_set_strict(self.h, int(on)))

    def set_tape(self, entries=None):
        """Replay tape (include/cabm.h cabm_set_tape): `entries` is a numpy array of TAPE_DTYPE — (seed, ctr[4] = (index, day,
        stream, sub), w[4]) — listing draws that take the given words instead of their Philox value; None removes the tape.
        With a tape the per-day kernels run (path "multikernel")."""
        if entries is None or len(entries) == 0:
            self._chk(self.L.cabm_set_tape(self.h, None, 0))
            return
        a = np.ascontiguousarray(entries, dtype=TAPE_DTYPE)
        self._chk(self.L.cabm_set_tape(self.h, a.ctypes.data, len(a)))

    @property
    def last_run_ms(self):
        ms = C.c_float(0)
        self._chk(self.L.cabm_last_run_ms(self.h, C.byref(ms)))
        return ms.value

    def mark(self):
        """Record a device time stamp on this handle's stream (cabm_mark)."""
        self._chk(self.L.cabm_mark(self.h))

    def ms_since_mark(self, marked=None):
        """Device milliseconds from `marked`'s mark (default: this handle's) to the end of this handle's last run."""
        ms = C.c_float(0)
        self._chk(self.L.cabm_ms_since_mark((marked or self).h, self.h, C.byref(ms)))
        return ms.value

    @property
    def launch_count(self):
        return int(self.L.cabm_launch_count(self.h))

    def set_profiling(self, on=True):
        self._chk(self.L.cabm_set_profiling(self.h, int(on)))

    def profile(self):
        """{class: (milliseconds, launches)} of the last profiled run."""
        ms = np.zeros(6, dtype=np.float32)
        n = np.zeros(6, dtype=np.int32)
        self._chk(self.L.cabm_get_profile(self.h, ms.ctypes.data, n.ctypes.data))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(["dyntrans", "time_update", "ct_identify", "other", "fused", "time_update_events"])}

    def set_path(self, path):
        """'auto' | 'multikernel' | 'fused' | 'fused-global' (include/cabm.h CABM_PATH_*)."""
        self._chk(self.L.cabm_set_path(self.h, {"auto": 0, "multikernel": 1, "fused": 2, "fused-global": 3}[path]))

    @property
    def last_path(self):
        return {0: None, 1: "multikernel", 2: "fused", 3: "fused-global"}[self.L.cabm_get_last_path(self.h)]

    @property
    def day(self):
        return self.L.cabm_get_day(self.h)

    # --- the phases, as test/runtests.jl drives them
    def initialize(self):
        self._chk(self.L.cabm_step_initialize(self.h))

    def insert_infected(self, health, num, ag, call_id=0):
        self._chk(self.L.cabm_step_insert_infected(self.h, int(health), num, ag, call_id))

    def dyntrans(self):
        self._chk(self.L.cabm_step_dyntrans(self.h))
        met = np.zeros(self.n_rep, dtype=np.int32)
        inf = np.zeros(self.n_rep, dtype=np.int32)
        self._chk(self.L.cabm_get_dyntrans_counters(self.h, met.ctypes.data, inf.ctypes.data))
        return met, inf

    def time_update(self):
        self._chk(self.L.cabm_step_time_update(self.h))

    def field(self, name, replicate=0):
        n = self.N * (4 if name == "dur" else 1)
        out = np.zeros(n, dtype=np.int32)
        self._chk(self.L.cabm_get_field(self.h, FIELDS[name], replicate, out.ctypes.data))
        return out.reshape(self.N, 4) if name == "dur" else out

    def set_field(self, name, agent, value, replicate=0):
        self._chk(self.L.cabm_set_field(self.h, FIELDS[name], replicate, agent, int(value)))

    # --- results
    def hmatrix(self, replicate=0):
        """Matrix{Int16}(popsize, modeltime), column-major like Julia (:111,:141)."""
        buf = np.zeros((self.T, self.N), dtype=np.int16)
        self._chk(self.L.cabm_get_hmatrix(self.h, replicate, buf.ctypes.data))
        return buf.T

    def hmatrix_all(self, out=None, dtype=np.int16):
        """[n_rep][T][N] state matrices; dtype int16 as main returns them (:111), or uint8 as the device keeps them (half the bytes)."""
        if out is None:
            out = np.zeros((self.n_rep, self.T, self.N), dtype=dtype)
        if out.dtype == np.uint8:
            self._chk(self.L.cabm_get_hmatrix_all_u8(self.h, out.ctypes.data))
        else:
            self._chk(self.L.cabm_get_hmatrix_all(self.h, out.ctypes.data))
        return out

    def curves(self, inc=None, prev=None, dtype=np.int32):
        """(inc, prev) [n_rep][6][T][12]; dtype int32, or int16 (narrowed on the device, popsize <= 32767)."""
        shape = (self.n_rep, 6, self.T, 12)
        inc = np.zeros(shape, dtype=dtype) if inc is None else inc
        prev = np.zeros(shape, dtype=dtype) if prev is None else prev
        if inc.dtype == np.int16:
            self._chk(self.L.cabm_get_curves_i16(self.h, inc.ctypes.data, prev.ctypes.data))
        else:
            self._chk(self.L.cabm_get_curves(self.h, inc.ctypes.data, prev.ctypes.data))
        return inc, prev

    def stream_curves_to(self, inc, prev, rows=None):
        """Register two PinnedArray-backed Int16 arrays [R][6][T][12]: the fused kernel writes each finished replicate's curves
        into them directly; a later curves(inc, prev) with the same arrays costs only a synchronise. None, None unregisters.
        rows: optional PinnedArray-backed int32 array [R]. With it the kernel sends only the leading rows (days) of a replicate
        that differ — after an epidemic has died out incidence is 0 and prevalence repeats — and stores their number in
        rows[r]; expand_curves(inc, prev, rows) fills the rest in on the host when dense arrays are wanted."""
        self._chk(self.L.cabm_set_host_curve_rows(self.h, None if rows is None else rows.ctypes.data))
        self._chk(self.L.cabm_set_host_curves_i16(self.h, None if inc is None else inc.ctypes.data, None if prev is None else prev.ctypes.data))

    def mean_curves(self):
        """compute_yearly_average (scripts/local_run.jl:75-103) on the device: (inc_mean, prev_mean) [6][T][12] float64."""
        a = np.zeros((6, self.T, 12), dtype=np.float64)
        b = np.zeros((6, self.T, 12), dtype=np.float64)
        self._chk(self.L.cabm_get_mean_curves(self.h, a.ctypes.data, b.ctypes.data))
        return a, b

    def scalars(self):
        inf = np.zeros((self.n_rep, 4), dtype=np.int64)
        tiso = np.zeros(self.n_rep, dtype=np.int64)
        lat = np.zeros(self.n_rep, dtype=np.int64)
        err = np.zeros(self.n_rep, dtype=np.uint32)
        self._chk(self.L.cabm_get_scalars(self.h, inf.ctypes.data, tiso.ctypes.data, lat.ctypes.data, err.ctypes.data))
        return dict(infectors=inf, totalisolated=tiso, lat_inc_sum=lat, err=err)

    def totalisolated(self):
        tiso = np.zeros(self.n_rep, dtype=np.int64)
        self._chk(self.L.cabm_get_scalars(self.h, None, tiso.ctypes.data, None, None))
        return tiso

    def errors(self):
        err = np.zeros(self.n_rep, dtype=np.uint32)
        self._chk(self.L.cabm_get_scalars(self.h, None, None, None, err.ctypes.data))
        return err

    def check_errors(self):
        err = self.errors()
        bad = np.nonzero(err)[0]
        if len(bad):
            bits = int(err[bad[0]])
            msgs = [m for b, m in ERR_BITS.items() if bits & b]
            raise CabmError(f"replicate {int(bad[0])}: " + "; ".join(msgs))

    def reduce_hmatrix(self, hm, ags):
        """_get_column_incidence :270-280 / _get_column_prevalence :282-293 on device for caller matrices.
        hm: [M][T][N] int16 (or (N, T) for one matrix), ags: [M][N] in 1..5 -> (inc, prev) [M][6][T][12]."""
        hm = np.asarray(hm)
        if hm.ndim == 2:
            hm = np.ascontiguousarray(hm.T)[None]
            ags = np.asarray(ags)[None]
        hm = np.ascontiguousarray(hm, dtype=np.int16)
        ags = np.ascontiguousarray(ags, dtype=np.int32)
        M, T, N = hm.shape
        inc = np.zeros((M, 6, T, 12), dtype=np.int32)
        prev = np.zeros((M, 6, T, 12), dtype=np.int32)
        self._chk(self.L.cabm_reduce_hmatrix(self.h, hm.ctypes.data, M, N, T, ags.ctypes.data, inc.ctypes.data, prev.ctypes.data))
        return inc, prev


# ------------------------------------------------------------------ the reference's entry points
_last = {}


def main(ip, seed=0, device=0, path="auto"):
    """main(ip::ModelParameters) :104-142 -> Matrix{Int16}(popsize, modeltime).

    Julia seeds its global RNG outside main (runsim :78); here the replicate's Philox key is `seed`.
    The final agent state stays readable through `humans()` like the reference's global `humans`."""
    pop = Population(1, ip.popsize, ip.modeltime, store_hmatrix=True, device=device)
    pop.configure(ip, [seed])
    pop.set_path(path)
    pop.run()
    pop.check_errors()
    old = _last.pop("pop", None)
    if old is not None:
        old.close()
    _last["pop"] = pop
    return pop.hmatrix(0)


def humans():
    """The population left behind by the last main()/runsim() call (the reference's global `humans` :69)."""
    return _last.get("pop")


_STATE_NAMES = [h.name for h in HEALTH][:12]


def collectdf(inc, prev):
    """_collectdf :226-245 for one group: dict of columns time, secondarys, <S>_INC, <S>_PREV."""
    T = inc.shape[0]
    out = {"time": np.arange(1, T + 1)}
    lat_inc = inc[:, LAT].astype(np.float64)
    inf_prv = prev[:, PRE:IISO + 1].sum(axis=1).astype(np.float64)      # mdf_prev[:, 3:8] :238
    with np.errstate(divide="ignore", invalid="ignore"):
        sec = lat_inc / inf_prv
    sec[~np.isfinite(sec)] = 0.0                                          # :240-241
    out["secondarys"] = sec
    for s, name in enumerate(_STATE_NAMES):
        out[name + "_INC"] = inc[:, s].astype(np.int64)
    for s, name in enumerate(_STATE_NAMES):
        out[name + "_PREV"] = prev[:, s].astype(np.int64)
    return out


def runsim(simnum, ip, device=0):
    """runsim(simnum, ip) :77-101: seed simnum*726 (:78), main, and the result tuple
    (a, g1..g5, infectors, ct_numbers) with each group as a dict of columns (DataFrame-shaped)."""
    hm = main(ip, seed=simnum * 726, device=device)
    pop = _last["pop"]
    inc, prev = pop.curves()
    sc = pop.scalars()
    res = {}
    for k, name in enumerate(["a", "g1", "g2", "g3", "g4", "g5"]):
        df = collectdf(inc[0, k], prev[0, k])
        df["sim"] = np.full(ip.modeltime, simnum)
        res[name] = df
    res["infectors"] = tuple(int(x) for x in sc["infectors"][0])
    res["ct_numbers"] = (0, 0, int(sc["totalisolated"][0]), 0, 0, 0, 0)     # :86-87; only totalisolated is ever written
    res["hmatrix"] = hm
    return res


def run_ensemble(ip, nsims, seeds=None, pset_of_rep=None, mode=MODE_EXACT, want_hmatrix=False, device=0, pop=None, path="auto", strict=True):
    """pmap(x -> runsim(x, ip), 1:nsims) (scripts/local_run.jl:28-30) as one device launch sequence.
    ip may be a list of ModelParameters with pset_of_rep choosing one per replicate (β sweeps).
    A replicate that hits one of the reference's error("...") conditions raises CabmError (strict=False: returned in "err")."""
    first = ip[0] if isinstance(ip, (list, tuple)) else ip
    if seeds is None:
        seeds = [726 * (i + 1) for i in range(nsims)]
    own = pop is None
    if own:
        pop = Population(nsims, first.popsize, first.modeltime, store_hmatrix=want_hmatrix, device=device)
    try:
        pop.configure(ip, seeds, pset_of_rep=pset_of_rep, mode=mode)
        pop.set_path(path)
        pop.set_strict(strict)
        pop.run()
        inc, prev = pop.curves()
        out = dict(inc=inc, prev=prev, **pop.scalars())
        if want_hmatrix:
            out["hmatrix"] = pop.hmatrix_all()
        out["device_ms"] = pop.last_run_ms
        out["path"] = pop.last_path
    finally:
        if own:
            pop.close()
    return out


def expand_curves(inc, prev, rows):
    """Dense curves from the rows-only form of stream_curves_to(..., rows): for every replicate the rows from rows[r] on are
    incidence 0 and the prevalence of row rows[r] - 1. In place; returns (inc, prev)."""
    T = inc.shape[2]
    for r in np.nonzero(np.asarray(rows) < T)[0]:
        d = int(rows[r])
        inc[r, :, d:, :] = 0
        prev[r, :, d:, :] = prev[r, :, d - 1:d, :]
    return inc, prev


class EnsemblePipeline:
    """pmap(x -> runsim(x, ip), 1:nsims) for ensembles larger than one batch: `handles` Population handles of `batch` replicates
    on one GPU, used round-robin. Each handle has its own stream, so consecutive batches overlap: the replicates of batch k+1
    that die out fill the SMs which the few long epidemics of batch k leave idle (the reference's pmap balances replicates over
    workers dynamically in the same way, scripts/local_run.jl:28-30). Results are identical to running the batches one by one.

        with cv.EnsemblePipeline(1000, ip.popsize, ip.modeltime) as pipe:
            for k in range(20):
                pipe.submit(ip, seeds_of_batch[k])          # waits for (and returns) the batch that last used this slot
            rest = pipe.drain()
    """

    def __init__(self, batch, popsize, modeltime, handles=3, store_hmatrix=False, device=0, path="auto", pinned=True, copy=True):
        self.batch, self.T, self.copy = batch, modeltime, copy   # copy=False: results are views of the slot's pinned buffers, valid until it is reused
        self.pops = [Population(batch, popsize, modeltime, store_hmatrix=store_hmatrix, device=device) for _ in range(handles)]
        self.bufs = []
        shape = (batch, 6, modeltime, 12)
        for p in self.pops:
            p.set_path(path)
            if pinned and popsize <= 32767:          # Int16 curves streamed into pinned memory by the fused kernel
                a, b = PinnedArray(shape, np.int16), PinnedArray(shape, np.int16)
                p.stream_curves_to(a.array, b.array)
                self.bufs.append((a, b))
            else:
                self.bufs.append(None)
        self.pending = [None] * handles           # user tag of the batch in flight on each handle
        self.outs = [None] * handles              # caller-supplied pinned result arrays of that batch (submit(out=...))
        self.own = [b is not None for b in self.bufs]   # the handle currently streams into its own buffers
        self.k = 0

    def _collect(self, j):
        p, tag = self.pops[j], self.pending[j]
        if tag is None:
            return None
        self.pending[j] = None
        if self.outs[j] is not None:               # the kernel wrote straight into the caller's pinned arrays
            inc, prev = p.curves(self.outs[j][0], self.outs[j][1])
            rows = self.outs[j][2] if len(self.outs[j]) > 2 else None
            self.outs[j] = None
            res = dict(tag=tag, inc=inc, prev=prev, **p.scalars())
            if rows is not None:
                res["rows"] = rows
            return res
        elif self.bufs[j] is not None:
            inc, prev = p.curves(self.bufs[j][0].array, self.bufs[j][1].array)      # same pointers: waits for the run only
            if self.copy:
                inc, prev = inc.copy(), prev.copy()
        else:
            inc, prev = p.curves()
        return dict(tag=tag, inc=inc, prev=prev, **p.scalars())

    def submit(self, psets, seeds, pset_of_rep=None, tag=None, mode=MODE_EXACT, collect=True, out=None):
        """Start one batch; returns the results of the batch that used this handle before (None the first time round, or
        with collect=False, where the previous results are dropped — timing loops). out=(inc, prev): Int16 arrays
        [batch][6][T][12] in pinned host memory (e.g. slices of one PinnedArray for the whole ensemble) that this batch's
        curves are written into by the kernel — no copy on the host at all."""
        j = self.k % len(self.pops)
        prev = self._collect(j) if collect else None
        p = self.pops[j]
        if out is not None:                                      # (inc, prev) or (inc, prev, rows): see Population.stream_curves_to
            p.stream_curves_to(*out)
            self.outs[j], self.own[j] = out, False
        elif self.bufs[j] is not None and not self.own[j]:       # back to the slot's own buffers
            p.stream_curves_to(self.bufs[j][0].array, self.bufs[j][1].array, None)
            self.own[j] = True
        p.configure(psets, seeds, pset_of_rep, mode)
        p.run(sync=False)
        self.pending[j] = self.k if tag is None else tag
        self.k += 1
        return prev

    def drain(self):
        """Results of the batches still in flight, oldest first."""
        n = len(self.pops)
        order = [(self.k + i) % n for i in range(n)]
        return [r for r in (self._collect(j) for j in order) if r is not None]

    def sync(self):
        for p in self.pops:
            p.sync()

    def close(self):
        for p, b in zip(self.pops, self.bufs):
            p.close()
            if b is not None:
                b[0].free(); b[1].free()
        self.pops, self.bufs = [], []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def run_ensemble_pipelined(ip, nsims, batch=1000, handles=3, seeds=None, device=0, path="auto", pipe=None, pinned=None, rows_only=False):
    """run_ensemble for nsims > batch: the batches are pipelined over `handles` handles (EnsemblePipeline). With popsize <= 32767
    the curves come back as Int16 views of two pinned arrays for the whole ensemble (kept alive by the returned dict), written
    by the kernels themselves: the host never copies them. Creating the handles (several GB of device memory) and the pinned
    arrays takes a few hundred milliseconds — far longer than the simulation of a few thousand replicates — so repeated
    calls (β sweeps, calibration) should pass their own `pipe` (an EnsemblePipeline(batch, popsize, modeltime, pinned=False))
    and `pinned` (two PinnedArray((nsims, 6, modeltime, 12), int16)) and keep them. rows_only=True (even modeltime): the kernels
    send only the rows of each replicate's curves before its epidemic died out; the result then has "rows" [nsims] and the arrays
    hold valid data in rows < rows[r] only (expand_curves fills the rest in)."""
    first = ip[0] if isinstance(ip, (list, tuple)) else ip
    if seeds is None:
        seeds = [726 * (i + 1) for i in range(nsims)]
    if nsims % batch:
        raise ValueError("nsims must be a multiple of batch")
    T = first.modeltime
    direct = first.popsize <= 32767
    if direct and pinned is None:
        pinned = (PinnedArray((nsims, 6, T, 12), np.int16), PinnedArray((nsims, 6, T, 12), np.int16))
    rows = PinnedArray((nsims,), np.int32) if (direct and rows_only) else None
    own_pipe = pipe is None
    if own_pipe:
        pipe = EnsemblePipeline(batch, first.popsize, T, handles=handles, device=device, path=path, pinned=False)
    parts = []
    try:
        base = pipe.k
        for k in range(nsims // batch):
            sl = slice(k * batch, (k + 1) * batch)
            out = None
            if direct:
                out = (pinned[0].array[sl], pinned[1].array[sl]) + ((rows.array[sl],) if rows is not None else ())
            r = pipe.submit(ip, seeds[sl], tag=base + k, out=out)
            if r is not None:
                parts.append(r)
        parts += pipe.drain()
    finally:
        if own_pipe:
            pipe.close()
    parts.sort(key=lambda r: r["tag"])
    res = {k: np.concatenate([r[k] for r in parts]) for k in parts[0] if k not in ("tag", "inc", "prev", "rows")}
    if direct:
        res["inc"], res["prev"], res["_pinned"] = pinned[0].array, pinned[1].array, pinned
        if rows is not None:
            res["rows"], res["_pinned_rows"] = rows.array, rows
    else:
        res["inc"], res["prev"] = np.concatenate([r["inc"] for r in parts]), np.concatenate([r["prev"] for r in parts])
    return res


def bind_host_thread_to_gpu(device):
    """One process per GPU: keep this process (and the pinned result buffers it allocates afterwards) on the CPU socket the GPU
    hangs off, so that eight ranks streaming curves to the host do not all cross the same inter-socket link. Best effort:
    returns False when NVML is not available or the affinity cannot be set."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(device))
        return True
    except Exception:
        return False


# ------------------------------------------------------------------ multi-GPU: replicate sharding
def shard_indices(nsims, rank, world):
    """Replicate r runs on rank r mod world (interleaved, so a β sweep spreads evenly): no hot-path traffic."""
    return list(range(rank, nsims, world))


def gather_ensemble(local, nsims, rank, world, dist=None, device="cpu", keys=("inc", "prev", "infectors", "totalisolated", "lat_inc_sum", "err"),
                    keep_on_device=False, events=None):
    """The path's single collective: end-of-run gather of the per-replicate summaries (pmap's implicit gather,
    scripts/local_run.jl:28-30,40-45) over torch.distributed (NCCL all_gather over NVLink on GPUs, gloo in the CPU tests).
    `local[k]` holds this rank's replicates in shard order (replicate r lives on rank r mod world); every rank returns the full
    arrays in replicate order — numpy arrays, or torch tensors left on `device` with keep_on_device=True (a further reduction
    on the GPU, e.g. compute_yearly_average, then needs no second copy). events=(start, end): CUDA events recorded around the
    collectives themselves (the host->device staging of the local arrays stays outside)."""
    if world == 1 or dist is None:
        return {k: np.asarray(local[k]) for k in keys if k in local}
    import torch
    nmax = (nsims + world - 1) // world
    staged = {}
    for k in keys:
        if k not in local:
            continue
        a = np.ascontiguousarray(local[k])
        if a.dtype == np.uint32:
            a = a.astype(np.int64)
        pad = np.zeros((nmax,) + a.shape[1:], dtype=a.dtype)
        pad[:a.shape[0]] = a
        staged[k] = torch.from_numpy(pad).to(device)
    if events is not None:
        events[0].record()
    gathered = {}
    for k, t in staged.items():
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        # NCCL has no 16-bit integer type: the Int16 curves travel as bytes
        ti, to = (t.view(torch.uint8), out.view(torch.uint8)) if t.dtype in (torch.int16, torch.uint16) else (t, out)
        dist.all_gather_into_tensor(to.reshape(-1), ti.reshape(-1))
        gathered[k] = out
    if events is not None:
        events[1].record()
    res = {}
    for k, out in gathered.items():
        # out[rk, j] is replicate rk + j * world: replicate order is the transpose of (rank, slot)
        full = out.transpose(0, 1).reshape((nmax * world,) + tuple(out.shape[2:]))[:nsims]
        res[k] = full.contiguous() if keep_on_device else full.cpu().numpy()
    return res


def run_ensemble_sharded(ip, nsims, rank, world, dist=None, device_index=0, seeds=None, pset_of_rep=None, **kw):
    """run_ensemble over `world` processes, one GPU each: this rank runs replicates rank, rank+world, ... and the
    summaries are gathered once at the end."""
    if seeds is None:
        seeds = [726 * (i + 1) for i in range(nsims)]
    idx = shard_indices(nsims, rank, world)
    loc = run_ensemble(ip, len(idx), seeds=[seeds[i] for i in idx],
                       pset_of_rep=None if pset_of_rep is None else [pset_of_rep[i] for i in idx], device=device_index, **kw)
    return gather_ensemble(loc, nsims, rank, world, dist=dist, device=f"cuda:{device_index}" if dist is not None and dist.get_backend() == "nccl" else "cpu")
