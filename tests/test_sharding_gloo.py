"""The N>1 host logic on CPU: world_size-2 gloo processes shard an ensemble (replicate r -> rank r mod 2), each computes
its shard (with the oracle standing in for the device, as the checker of the plumbing only) and the single end-of-run
gather (covid19abm_b200.gather_ensemble) must reproduce the single-process result in replicate order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NSIMS, N, T = 7, 400, 25


def _params(O):
    return O.make_params(beta=0.3, prov="ontario", popsize=N, modeltime=T, initialinf=3, fmild=0.5, tau_mild=1)


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import covid19abm_b200 as cv
    import oracle_api as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = [726 * (i + 1) for i in range(NSIMS)]
    idx = cv.shard_indices(NSIMS, rank, world)
    loc = O.run_ensemble([_params(O)], [0] * len(idx), [seeds[i] for i in idx], nthreads=1)
    loc["lat_inc_sum"] = loc["inc"][:, 0, :, 1].sum(axis=1).astype(np.int64)
    loc["err"] = np.zeros(len(idx), dtype=np.uint32)
    loc["inc"] = loc["inc"].astype(np.int16)          # the transport form of the curves (Int16: travels as bytes, NCCL has no 16-bit integers)
    full = cv.gather_ensemble(loc, NSIMS, rank, world, dist=dist, device="cpu")
    assert full["inc"].dtype == np.int16
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path, O, cv):
    assert cv.shard_indices(7, 0, 2) == [0, 2, 4, 6] and cv.shard_indices(7, 1, 2) == [1, 3, 5]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    seeds = [726 * (i + 1) for i in range(NSIMS)]
    ref = O.run_ensemble([_params(O)], [0] * NSIMS, seeds, nthreads=2)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        assert (got["inc"] == ref["inc"]).all() and (got["prev"] == ref["prev"]).all()
        assert (got["infectors"] == ref["infectors"]).all() and (got["totalisolated"] == ref["totalisolated"]).all()
        assert (got["lat_inc_sum"] == ref["inc"][:, 0, :, 1].sum(axis=1)).all()


# ------------------------------------------------------------------------------------------------ the same plumbing on GPUs
import pytest


def _nccl_worker(rank, world, port, out_dir, nsims):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import covid19abm_b200 as cv
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ip = cv.ModelParameters(β=0.07, prov="ontario", modeltime=120, initialinf=2, fmild=0.5, τmild=1, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3)
    full = cv.run_ensemble_sharded(ip, nsims, rank, world, dist=dist, device_index=rank)
    np.savez(os.path.join(out_dir, f"nccl_rank{rank}.npz"), **full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_sharded_ensemble_over_nccl(tmp_path, cv):
    """run_ensemble_sharded on two GPUs (one process each, replicate r on rank r mod 2) with the end-of-run all_gather over
    NCCL/NVLink: every rank ends up with the whole ensemble in replicate order, equal to the one-GPU run."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nsims = 37
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path), nsims), nprocs=2, join=True)
    ip = cv.ModelParameters(β=0.07, prov="ontario", modeltime=120, initialinf=2, fmild=0.5, τmild=1, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3)
    ref = cv.run_ensemble(ip, nsims)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"nccl_rank{rank}.npz"))
        for k in ("inc", "prev", "infectors", "totalisolated", "lat_inc_sum"):
            assert (got[k] == ref[k]).all(), (rank, k)
        assert (got["err"] == 0).all()
