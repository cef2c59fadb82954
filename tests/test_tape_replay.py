"""Replay tape (include/cabm.h cabm_set_tape; north_star: "a replay mode feeds each kernel the reference's own uniform draws").

The reference draws from one sequential stream; here every draw has a key (seed, index, day, stream, sub). A tape overrides
draws by key. Three properties are tested, on the oracle (CPU) and on the CUDA path (GPU):
  1. with the same tape both implementations produce the same state matrix, bit for bit;
  2. a tape that covers every draw a run consumes makes the result independent of the seed: every draw site consults the tape;
  3. a tape changes the result (it is not ignored), and removing it restores the seeded run."""
import numpy as np
import pytest

from test_gpu_parity import to_oracle

N, T, KMAX = 160, 24, 32
STREAM = {"INIT": 1, "COUNT": 2, "SEED": 3, "TRANS": 4, "CTWIN": 5, "CTCON": 6, "CT3": 7, "DT": 8}
KW = dict(β=0.25, prov="ontario", popsize=N, modeltime=T, initialinf=4, fmild=0.5, τmild=1, fsevere=0.7, fpreiso=0.4, tpreiso=3,
          ctstrat=3, strat3qdays=4, fctcapture=0.8, fcontactst=0.7, cdaysback=3, quar_grp_frac=0.3, quar_grp=3)


def full_tape(dtype, seed, rng_seed):
    """Words for every key a run of KW can consume (and many it does not). The words depend on rng_seed only."""
    rng = np.random.default_rng(rng_seed)
    keys = []

    def add(idx, day, stream, sub):
        g = np.stack(np.meshgrid(idx, day, [stream], sub, indexing="ij"), axis=-1).reshape(-1, 4)
        keys.append(g.astype(np.uint32))

    ag = np.arange(N)
    add(ag, [0], STREAM["INIT"], [0, 1])
    add(np.arange((N + 3) // 4), np.arange(1, T + 2), STREAM["COUNT"], [0])
    add(np.arange(KW["initialinf"]), [0], STREAM["SEED"], [0, 1])
    add(ag, np.arange(0, T + 1), STREAM["TRANS"], [0, 1])
    add(ag, np.arange(1, T + 1), STREAM["CTWIN"], [0])
    add(ag, np.arange(1, T + 1), STREAM["CT3"], [0])
    add(ag, np.arange(1, T + 1), STREAM["CTCON"], ag)
    add(ag, np.arange(1, T + 1), STREAM["DT"], [(g << 16) | k for g in range(5) for k in range(KMAX)])
    ctr = np.concatenate(keys)
    tape = np.zeros(len(ctr), dtype=dtype)
    tape["seed"] = seed
    tape["ctr"] = ctr
    w = rng.integers(0, 2 ** 32, size=(len(ctr), 4), dtype=np.uint64).astype(np.uint32)
    cnt = ctr[:, 2] == STREAM["COUNT"]
    w[cnt] = (w[cnt] >> 1) + (w[cnt] >> 2)          # contact-count words stay below 0.75 * 2^32: budgets within the KMAX slots listed per group
    tape["w"] = w
    return tape


def oracle_run(O, ip, seed, tape):
    O.set_tape(tape)
    try:
        sim = O.Sim(to_oracle(O, ip), seed)
        hm = np.asarray(sim.main()).copy()
        tiso = sim.totalisolated
        sim.close()
    finally:
        O.set_tape(None)
    return hm, tiso


def test_oracle_tape_replay_properties(O, cv):
    ip = cv.ModelParameters(**KW)
    plain, _ = oracle_run(O, ip, 11, None)
    a, ta = oracle_run(O, ip, 11, full_tape(O.TAPE_DTYPE, 11, rng_seed=5))
    b, tb = oracle_run(O, ip, 987654321012, full_tape(O.TAPE_DTYPE, 987654321012, rng_seed=5))
    assert (a == b).all() and ta == tb                       # the tape decides every draw: the seed no longer matters
    assert (a != plain).any()                                # ... and it is not ignored
    c, _ = oracle_run(O, ip, 11, full_tape(O.TAPE_DTYPE, 11, rng_seed=6))
    assert (a != c).any()                                    # other words, other trajectory
    assert (a[:, -1] != 0).sum() > 20                        # an epidemic did happen
    again, _ = oracle_run(O, ip, 11, None)
    assert (again == plain).all()                            # removing the tape restores the seeded run
    part = full_tape(O.TAPE_DTYPE, 11, rng_seed=5)
    part = part[part["ctr"][:, 2] == STREAM["TRANS"]]        # a partial tape overrides those draws only
    d, _ = oracle_run(O, ip, 11, part)
    assert (d != plain).any() and (d != a).any()


@pytest.mark.gpu
def test_gpu_replays_the_same_tape_bit_exactly(cv, O):
    ip = cv.ModelParameters(**KW)
    for seed, rs in ((11, 5), (987654321012, 5), (11, 6)):
        tape = full_tape(cv.TAPE_DTYPE, seed, rng_seed=rs)
        ref, tiso = oracle_run(O, ip, seed, tape)
        with cv.Population(1, N, T, store_hmatrix=True) as pop:
            pop.set_tape(tape)
            pop.configure(ip, [seed])
            pop.run()
            assert pop.last_path == "multikernel"
            got = np.asarray(pop.hmatrix(0))
            assert (got == ref).all(), (seed, rs)
            assert pop.scalars()["totalisolated"][0] == tiso
            if (seed, rs) == (11, 5):
                first = got.copy()
            if (seed, rs) == (987654321012, 5):
                assert (got == first).all()                  # seed-independent under a full tape, on the device as well
            # a partial tape: only the transition draws; the rest stay Philox
            part = tape[tape["ctr"][:, 2] == STREAM["TRANS"]]
            pref, _ = oracle_run(O, ip, seed, part)
            pop.set_tape(part); pop.configure(ip, [seed]); pop.run()
            assert (np.asarray(pop.hmatrix(0)) == pref).all()
            # without a tape the handle is back on the fused path and the seeded result
            pop.set_tape(None); pop.configure(ip, [seed]); pop.run()
            assert pop.last_path == "fused"
            plain, _ = oracle_run(O, ip, seed, None)
            assert (np.asarray(pop.hmatrix(0)) == plain).all()
            with pytest.raises(cv.CabmError, match="duplicate key"):
                pop.set_tape(np.concatenate([part[:3], part[:1]]))


@pytest.mark.gpu
def test_step_api_consults_the_tape(cv, O):
    """The phases driven one by one (cabm_step_*), as test/runtests.jl drives them: initialize under a tape."""
    ip = cv.ModelParameters(**KW)
    tape = full_tape(cv.TAPE_DTYPE, 42, rng_seed=9)
    O.set_tape(tape)
    try:
        sim = O.Sim(to_oracle(O, ip), 42)
        sim.initialize()
        ref = {f: sim.field(f) for f in ("ag", "nextday_meetcnt", "iso")}
        rdur = sim.field("dur")
        sim.close()
    finally:
        O.set_tape(None)
    with cv.Population(1, N, T) as pop:
        pop.set_tape(tape)
        pop.configure(ip, [42])
        pop.initialize()
        for f, v in ref.items():
            assert (pop.field(f) == v).all(), f
        assert (pop.field("dur") == rdur).all()
