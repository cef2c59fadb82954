"""One face over the two implementations so that the reference's invariant tests (test/runtests.jl) can be
re-expressed once and run against the CPU oracle (here, no GPU) and against the CUDA path (-m gpu)."""
import numpy as np

import oracle_api as O


class OracleBackend:
    name = "oracle"

    def __init__(self, seed=1, **kw):
        self.kw = dict(kw)
        self.seed = seed
        self.sim = O.Sim(O.make_params(**self._okw()), seed)
        self.n = self.sim.n

    def _okw(self):
        kw = dict(self.kw)
        for a, b in (("β", "beta"), ("τmild", "tau_mild")):
            if a in kw:
                kw[b] = kw.pop(a)
        return kw

    def initialize(self):
        self.sim.initialize()

    # main seeds with the user's fpreiso, zeroes it, and restores it on day tpreiso (:117-125,:133-135). The CUDA
    # step API applies that rule itself; the oracle follows the reference's mutable global `p` literally, so the
    # adapter plays main's part here.
    def insert_infected(self, health, num, ag, call_id=0):
        self.sim.p.fpreiso = self.kw.get("fpreiso", 0.0)
        self.sim.insert_infected(health, num, ag, call_id)

    def dyntrans(self):
        return self.sim.dyntrans()

    def time_update(self):
        tp = self.sim.p.tpreiso
        self.sim.p.fpreiso = self.kw.get("fpreiso", 0.0) if (tp >= 1 and self.sim.day >= tp) else 0.0
        self.sim.time_update()

    def field(self, name):
        if name == "ncontacts":
            hs, nb = self.sim.humans, (self.n + 7) // 8
            out = np.zeros(self.n, dtype=np.int32)
            for i in range(self.n):
                if hs[i].tracecontacts:
                    out[i] = int(np.unpackbits(np.ctypeslib.as_array(hs[i].tracecontacts, shape=(nb,))).sum())
            return out
        return self.sim.field(name)

    def set_field(self, name, i, value):
        setattr(self.sim.humans[i], name, int(value))

    def totalisolated(self):
        return self.sim.totalisolated

    def betavalue(self, t, health):
        return self.sim.get_betavalue(t, health)

    def close(self):
        self.sim.close()


class CudaBackend:
    name = "cuda"

    def __init__(self, seed=1, **kw):
        import covid19abm_b200 as cv
        self.cv = cv
        ip = cv.ModelParameters(**kw)
        self.pop = cv.Population(1, ip.popsize, ip.modeltime)
        self.pop.configure(ip, [seed])
        self.n = ip.popsize

    def initialize(self):
        self.pop.initialize()

    def insert_infected(self, health, num, ag, call_id=0):
        self.pop.insert_infected(health, num, ag, call_id)

    def dyntrans(self):
        met, inf = self.pop.dyntrans()
        return int(met[0]), int(inf[0])

    def time_update(self):
        self.pop.time_update()

    def field(self, name):
        return self.pop.field(name)

    def set_field(self, name, i, value):
        self.pop.set_field(name, i, value)

    def totalisolated(self):
        return int(self.pop.totalisolated()[0])

    def close(self):
        self.pop.close()
