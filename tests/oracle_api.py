"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/abm_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

PROVINCES = ["alberta", "bc", "canada", "manitoba", "newbruns", "newfdland", "nwterrito", "novasco",
             "nunavut", "ontario", "pei", "quebec", "saskat", "yukon", "newyork"]
SUS, LAT, PRE, ASYMP, MILD, MISO, INF, IISO, HOS, ICU, REC, DED, UNDEF = range(13)


class OrcParams(C.Structure):
    _fields_ = [("beta", C.c_double), ("frelasymp", C.c_double), ("seasonal", C.c_int32), ("popsize", C.c_int32),
                ("prov_id", C.c_int32), ("calibration", C.c_int32), ("modeltime", C.c_int32),
                ("initialinf", C.c_int32), ("initialhi", C.c_int32), ("tau_mild", C.c_int32),
                ("asymp_props", C.c_double * 5), ("mild_props", C.c_double * 5), ("fmild", C.c_double),
                ("fsevere", C.c_double), ("fpreiso", C.c_double), ("quar_grp_frac", C.c_double),
                ("tpreiso", C.c_int32), ("quar_grp", C.c_int32), ("ctstrat", C.c_int32), ("cdaysback", C.c_int32),
                ("strat3qdays", C.c_int32), ("fctcapture", C.c_float), ("fcontactst", C.c_float), ("_pad", C.c_int32)]


class OrcHuman(C.Structure):
    _fields_ = [("idx", C.c_int32), ("health", C.c_int32), ("swap", C.c_int32), ("sickfrom", C.c_int32),
                ("nextday_meetcnt", C.c_int32), ("ag", C.c_int32), ("tis", C.c_int32), ("exp", C.c_int32),
                ("dur", C.c_int32 * 4), ("doi", C.c_int32), ("iso", C.c_int32), ("isovia", C.c_int32),
                ("tracing", C.c_int32), ("tracestart", C.c_int32), ("traceend", C.c_int32), ("tracedxp", C.c_int32),
                ("tracecontacts", C.POINTER(C.c_uint8))]


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(ORACLE_DIR, "abm_oracle.c")):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcParams), C.c_uint64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_params_ptr.restype = C.POINTER(OrcParams)
        L.orc_params_ptr.argtypes = [C.c_void_p]
        L.orc_humans.restype = C.POINTER(OrcHuman)
        L.orc_humans.argtypes = [C.c_void_p]
        L.orc_betas.restype = C.POINTER(C.c_double)
        L.orc_betas.argtypes = [C.c_void_p]
        L.orc_totalisolated.restype = C.c_int64
        L.orc_totalisolated.argtypes = [C.c_void_p]
        L.orc_day.argtypes = [C.c_void_p]
        L.orc_set_day.argtypes = [C.c_void_p, C.c_int]
        L.orc_initialize.argtypes = [C.c_void_p]
        L.orc_insert_infected.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
        L.orc_dyntrans.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_time_update.argtypes = [C.c_void_p]
        L.orc_get_nextday_counts.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_betavalue.restype = C.c_double
        L.orc_get_betavalue.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_main.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_column_incidence.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_get_column_prevalence.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_count_infectors.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_collect_curves.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_run_ensemble.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_tape.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_table.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint32))]
        L.orc_gpw.argtypes = [C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def make_params(beta=0.0, frelasymp=0.11, seasonal=False, popsize=10000, prov="newyork", calibration=False,
                modeltime=300, initialinf=1, initialhi=0, asymp_props=(0.25, 0.25, 0.14, 0.07, 0.07),
                mild_props=(0.95, 0.9, 0.85, 0.6, 0.2), tau_mild=0, fmild=0.0, fsevere=0.0, tpreiso=0, fpreiso=0.0,
                quar_grp=5, quar_grp_frac=0.0, ctstrat=0, fctcapture=0.0, fcontactst=0.0, cdaysback=0,
                strat3qdays=0):
    """Defaults are the reference's (src/covid19abm.jl:31-55)."""
    p = OrcParams()
    p.beta, p.frelasymp, p.seasonal, p.popsize = beta, frelasymp, int(seasonal), popsize
    p.prov_id = PROVINCES.index(prov) if prov in PROVINCES else -1
    p.calibration, p.modeltime, p.initialinf, p.initialhi, p.tau_mild = int(calibration), modeltime, initialinf, initialhi, tau_mild
    p.asymp_props = (C.c_double * 5)(*asymp_props)
    p.mild_props = (C.c_double * 5)(*mild_props)
    p.fmild, p.fsevere, p.fpreiso, p.quar_grp_frac = fmild, fsevere, fpreiso, quar_grp_frac
    p.tpreiso, p.quar_grp, p.ctstrat, p.cdaysback, p.strat3qdays = tpreiso, quar_grp, ctstrat, cdaysback, strat3qdays
    p.fctcapture = float(np.float16(fctcapture))     # Float16 fields, src/covid19abm.jl:51-52
    p.fcontactst = float(np.float16(fcontactst))
    return p


class Sim:
    """One replicate of the oracle: the module-level `humans`, `p`, `BETAS`, `ct_data` of the reference."""

    def __init__(self, params, seed):
        self.L = lib()
        self.h = self.L.orc_create(C.byref(params), C.c_uint64(seed))
        if not self.h:
            raise OracleError(self.L.orc_last_error().decode())
        self.n = params.popsize
        self.T = params.modeltime

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc:
            raise OracleError(self.L.orc_last_error().decode())

    @property
    def p(self):
        return self.L.orc_params_ptr(self.h).contents

    @property
    def humans(self):
        return self.L.orc_humans(self.h)

    @property
    def betas(self):
        return np.ctypeslib.as_array(self.L.orc_betas(self.h), shape=(self.T,)).copy()

    @property
    def totalisolated(self):
        return self.L.orc_totalisolated(self.h)

    @property
    def day(self):
        return self.L.orc_day(self.h)

    def field(self, name):
        hs = self.humans
        if name == "dur":
            return np.array([[hs[i].dur[k] for k in range(4)] for i in range(self.n)], dtype=np.int32)
        return np.array([getattr(hs[i], name) for i in range(self.n)], dtype=np.int32)

    def initialize(self):
        self._chk(self.L.orc_initialize(self.h))

    def insert_infected(self, health, num, ag, call_id=0):
        out = (C.c_int32 * max(num, 1))()
        self._chk(self.L.orc_insert_infected(self.h, health, num, ag, call_id, out))
        return list(out)[:num]

    def dyntrans(self):
        a, b = C.c_int(0), C.c_int(0)
        self._chk(self.L.orc_dyntrans(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def time_update(self):
        self._chk(self.L.orc_time_update(self.h))

    def get_nextday_counts(self, i):
        return self.L.orc_get_nextday_counts(self.h, i)

    def get_betavalue(self, sys_time, xhealth):
        return self.L.orc_get_betavalue(self.h, sys_time, xhealth)

    def main(self, want_hmatrix=True):
        """main(ip) :104-142 — returns the popsize x modeltime Int16 matrix (Fortran order, like Julia)."""
        hm = np.zeros((self.T, self.n), dtype=np.int16)
        self._chk(self.L.orc_main(self.h, hm.ctypes.data if want_hmatrix else None))
        return hm.T  # (popsize, modeltime) view, column-major storage

    def count_infectors(self):
        out = np.zeros(4, dtype=np.int64)
        self._chk(self.L.orc_count_infectors(self.h, out.ctypes.data))
        return out


def _hm_storage(hm):
    """hm: (popsize, modeltime) array -> contiguous [modeltime][popsize] int16 buffer."""
    return np.ascontiguousarray(np.asarray(hm, dtype=np.int16).T)


def get_column_incidence(hm, hcol):
    s = _hm_storage(hm)
    out = np.zeros(s.shape[0], dtype=np.int64)
    lib().orc_get_column_incidence(s.ctypes.data, s.shape[1], s.shape[0], hcol, out.ctypes.data)
    return out


def get_column_prevalence(hm, hcol):
    s = _hm_storage(hm)
    out = np.zeros(s.shape[0], dtype=np.int64)
    lib().orc_get_column_prevalence(s.ctypes.data, s.shape[1], s.shape[0], hcol, out.ctypes.data)
    return out


def collect_curves(hm, ags):
    s = _hm_storage(hm)
    T, n = s.shape
    a = np.ascontiguousarray(ags, dtype=np.int32)
    inc = np.zeros((6, T, 12), dtype=np.int32)
    prev = np.zeros((6, T, 12), dtype=np.int32)
    lib().orc_collect_curves(s.ctypes.data, n, T, a.ctypes.data, inc.ctypes.data, prev.ctypes.data)
    return inc, prev


def run_ensemble(psets, pset_of_rep, seeds, nthreads=1, want_hmatrix=False):
    """pmap(runsim) restated: returns dict(inc, prev, infectors, totalisolated[, hmatrix])."""
    psets = list(psets)
    arr = (OrcParams * len(psets))(*psets)
    nrep = len(seeds)
    T, n = psets[0].modeltime, psets[0].popsize
    por = np.ascontiguousarray(pset_of_rep, dtype=np.int32)
    sd = np.ascontiguousarray(seeds, dtype=np.uint64)
    inc = np.zeros((nrep, 6, T, 12), dtype=np.int32)
    prev = np.zeros((nrep, 6, T, 12), dtype=np.int32)
    infectors = np.zeros((nrep, 4), dtype=np.int64)
    tiso = np.zeros(nrep, dtype=np.int64)
    hm = np.zeros((nrep, T, n), dtype=np.int16) if want_hmatrix else None
    failed = lib().orc_run_ensemble(arr, len(psets), por.ctypes.data, nrep, sd.ctypes.data, nthreads,
                                    inc.ctypes.data, prev.ctypes.data, infectors.ctypes.data, tiso.ctypes.data,
                                    hm.ctypes.data if want_hmatrix else None)
    if failed:
        raise OracleError(f"{failed} replicates failed")
    out = dict(inc=inc, prev=prev, infectors=infectors, totalisolated=tiso)
    if want_hmatrix:
        out["hmatrix"] = hm
    return out


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def table(name):
    base = C.c_int(0)
    ptr = C.POINTER(C.c_uint32)()
    n = lib().orc_table(name.encode(), C.byref(base), C.byref(ptr))
    if n < 0:
        raise KeyError(name)
    return base.value, np.ctypeslib.as_array(ptr, shape=(n,)).copy()


def gpw(ag, cnts):
    out = np.zeros(5, dtype=np.int32)
    lib().orc_gpw(ag, cnts, out.ctypes.data)
    return out


TAPE_DTYPE = np.dtype([("seed", "<u8"), ("ctr", "<u4", (4,)), ("w", "<u4", (4,))])       # orc_tape_entry


def set_tape(entries=None):
    """Replay tape of the oracle (process-wide): same layout and meaning as cabm_set_tape. None removes it."""
    if entries is None or len(entries) == 0:
        lib().orc_set_tape(None, 0)
        return
    a = np.ascontiguousarray(entries, dtype=TAPE_DTYPE)
    if lib().orc_set_tape(a.ctypes.data, len(a)):
        raise OracleError(lib().orc_last_error().decode())
