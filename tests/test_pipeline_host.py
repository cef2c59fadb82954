"""Host logic of cv.EnsemblePipeline (round-robin over handles, collecting the batch that used a slot before, drain order)
with the device side replaced by a recording stand-in: no GPU needed."""
import numpy as np
import pytest


class FakePopulation:
    made = []

    def __init__(self, n_rep, popsize, modeltime, store_hmatrix=False, device=0):
        self.n_rep, self.N, self.T = n_rep, popsize, modeltime
        self.log, self.seeds, self.closed, self.host = [], None, False, None
        FakePopulation.made.append(self)

    def set_path(self, path): self.log.append(("path", path))
    def stream_curves_to(self, inc, prev, rows=None): self.host = (inc, prev)
    def configure(self, psets, seeds, pset_of_rep=None, mode=0): self.seeds = np.asarray(seeds); self.log.append(("configure", int(self.seeds[0])))
    def run(self, sync=True): self.log.append(("run", sync))
    def sync(self): self.log.append(("sync",))
    def close(self): self.closed = True

    def curves(self, inc=None, prev=None):
        shape = (self.n_rep, 6, self.T, 12)
        inc = np.zeros(shape, np.int32) if inc is None else inc
        prev = np.zeros(shape, np.int32) if prev is None else prev
        inc[:] = (self.seeds % 1000)[:, None, None, None]          # "results" that identify the batch
        prev[:] = inc + 1
        return inc, prev

    def scalars(self):
        return dict(infectors=np.tile(self.seeds[:, None], (1, 4)), totalisolated=self.seeds * 0, lat_inc_sum=self.seeds * 2,
                    err=np.zeros(self.n_rep, np.uint32))


class FakePinned:
    def __init__(self, shape, dtype): self.array = np.zeros(shape, dtype); self.freed = False
    def free(self): self.freed = True


@pytest.fixture
def fake(cv, monkeypatch):
    import sys
    FakePopulation.made = []
    mod = sys.modules[cv.EnsemblePipeline.__module__]       # the package module behind the import shim
    monkeypatch.setattr(mod, "Population", FakePopulation)
    monkeypatch.setattr(mod, "PinnedArray", FakePinned)
    return cv


def test_submit_round_robin_and_collect(fake):
    cv = fake
    with cv.EnsemblePipeline(4, 100, 5, handles=2) as pipe:
        assert len(FakePopulation.made) == 2 and all(p.host is not None for p in FakePopulation.made)   # Int16 buffers registered
        out = [pipe.submit("ip", [10 * k + 1] * 4) for k in range(5)]
        assert out[0] is None and out[1] is None                               # nothing to collect the first time round
        assert [o["tag"] for o in out[2:]] == [0, 1, 2]                        # then the batch that used the slot before
        assert (out[2]["inc"] == 1).all() and (out[3]["inc"] == 11).all() and (out[4]["prev"] == 22).all()
        assert out[2]["inc"].dtype == np.int16
        rest = pipe.drain()
        assert [o["tag"] for o in rest] == [3, 4] and (rest[1]["lat_inc_sum"] == 82).all()
        assert pipe.drain() == []
        assert [p.log.count(("run", False)) for p in FakePopulation.made] == [3, 2]
    assert all(p.closed for p in FakePopulation.made)


def test_copy_false_returns_views_of_the_slot_buffers(fake):
    cv = fake
    pipe = cv.EnsemblePipeline(2, 100, 3, handles=2, copy=False)
    pipe.submit("ip", [5, 6]); pipe.submit("ip", [7, 8])
    a = pipe.submit("ip", [9, 9])                                              # collects batch 0 from slot 0
    assert a["inc"] is pipe.bufs[0][0].array
    pipe.close()
    assert pipe.pops == []


def test_run_ensemble_pipelined_concatenates_in_replicate_order(fake):
    cv = fake
    ip = cv.ModelParameters(modeltime=4, popsize=100)
    seeds = list(range(1, 13))
    out = cv.run_ensemble_pipelined(ip, 12, batch=3, handles=3, seeds=seeds)
    assert (out["lat_inc_sum"] == 2 * np.arange(1, 13)).all()
    assert out["inc"].shape == (12, 6, 4, 12) and (out["inc"][:, 0, 0, 0] == np.arange(1, 13)).all()
    with pytest.raises(ValueError):
        cv.run_ensemble_pipelined(ip, 10, batch=3)
