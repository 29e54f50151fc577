"""Multi-rank path.  CPU (gloo, world_size 2 and 4): the exchange logic of fusion_sim_b200/dist.py --
slab bounds, the fixed-region all-to-all of particle records (and the exact all-to-all-v), 5-row halo
of the per-cell sums -- driven with an oracle-backed rank, must reproduce the single-process oracle BIT
FOR BIT (particles by global id, counts and running average by global cell).  GPU (nccl, 2 / 4 / 8
devices): the same check with the CUDA SlabPusher."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import assert_same
import dist_helpers as dh

FRAMES = 6
ONLY_WORLD = int(os.environ.get("FSIM_TEST_WORLD", "0"))  # a multi-GPU box is paid per GPU: run one world size per call


def need_gpus(world):
    if ONLY_WORLD and world != ONLY_WORLD:
        pytest.skip(f"FSIM_TEST_WORLD={ONLY_WORLD}")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def check_against_single(path, solve=False):
    got = np.load(path)
    o = dh.single_oracle(FRAMES, solve)
    if solve:  # the slab-decomposed field solve equals the single-process one bit for bit
        assert_same(got["phi"], o.getField("phi"), "potential")
        assert_same(got["E"], o.getField("E"), "E")
        assert np.abs(got["phi"]).max() > 0 and np.abs(o.A).max() > 0
    assert_same(got["ids"], np.arange(o.n, dtype=got["ids"].dtype), "every particle exactly once")
    assert_same(got["pos"], o.getPosition(), "position")
    assert_same(got["vel"], o.getVelocity(), "velocity")
    assert_same(got["rnd"], o.getRand(), "rand")
    assert_same(got["cnt"], o.cell_count, "cell counts")
    assert_same(got["avg"], o.moments01_avg, "running average")
    assert int(got["moved"]) > 0  # particles really crossed the slab boundary / respawned across it


def test_exchange_region_layout():
    """Host side of the fixed-region exchange: both ends of a pair compute the same sizes without talking."""
    from fusion_sim_b200.dist import HEADER_BYTES, region_bytes, region_capacities
    for world in (2, 4, 8):
        caps = [region_capacities(r, world, 1000, 10) for r in range(world)]
        for a in range(world):
            assert caps[a][a] == 0
            for b in range(world):
                assert caps[a][b] == caps[b][a]  # what a sends to b is what b expects from a
                if a != b:
                    assert caps[a][b] == (1000 if abs(a - b) == 1 else 10)
    assert region_bytes(0, 88) == HEADER_BYTES and region_bytes(3, 88) % 16 == 0 and region_bytes(3, 88) >= HEADER_BYTES + 3 * 88


@pytest.mark.gpu
def test_gpus_exchange_overflow_is_reported(tmp_path):
    """A send region too small for a frame's leavers: sync() raises FSIM_ERR_RANGE on the rank it happened on."""
    need_gpus(2)
    path = str(tmp_path / "msg.txt")
    mp.spawn(dh.gpu_worker_overflow, args=(2, free_port(), path), nprocs=2, join=True)
    msgs = open(path).read().split("\n")
    assert any("send region" in m for m in msgs), msgs


def test_slab_bounds():
    from fusion_sim_b200.dist import slab_bounds
    assert slab_bounds(800, 8) == [0, 100, 200, 300, 400, 500, 600, 700, 800]
    b = slab_bounds(10, 3)
    assert b[0] == 0 and b[-1] == 10 and all(x < y for x, y in zip(b, b[1:]))


@pytest.mark.parametrize("world,exchange", [(2, "fixed"), (2, "exact"), (4, "fixed")])
def test_ranks_gloo_match_single_oracle(tmp_path, world, exchange):
    path = str(tmp_path / "res.npz")
    mp.spawn(dh.cpu_worker, args=(world, free_port(), FRAMES, path, False, exchange), nprocs=world, join=True)
    check_against_single(path)
    if world > 2:  # a respawn crossed more than one slab boundary: the all-to-all is not neighbour-only
        assert int(np.load(path)["far"]) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("world,exchange", [(2, "fixed"), (2, "exact"), (4, "fixed"), (8, "fixed"), (8, "exact")])
def test_gpus_nccl_match_single_oracle(tmp_path, world, exchange):
    """world ranks on world GPUs against the single-process oracle, bit for bit.  With 4 and 8 slabs the
    middle ranks have two neighbours (both halos), the source region (rows 28..36 of 64) lies outside
    the cell table of the outer ranks, and particles absorbed there respawn into a NON-neighbouring slab."""
    need_gpus(world)
    path = str(tmp_path / "res.npz")
    mp.spawn(dh.gpu_worker, args=(world, free_port(), FRAMES, path, False, exchange), nprocs=world, join=True)
    check_against_single(path)


def test_two_ranks_gloo_self_consistent_fields(tmp_path):
    """EXTENSION: step -> density -> solveFields on two slabs (halo rows of charge
    source, potential and E exchanged between the stages) against the single-process oracle."""
    path = str(tmp_path / "res.npz")
    mp.spawn(dh.cpu_worker, args=(2, free_port(), FRAMES, path, True), nprocs=2, join=True)
    check_against_single(path, solve=True)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_gpus_nccl_self_consistent_fields(tmp_path, world):
    need_gpus(world)
    path = str(tmp_path / "res.npz")
    mp.spawn(dh.gpu_worker, args=(world, free_port(), FRAMES, path, True), nprocs=world, join=True)
    check_against_single(path, solve=True)


@pytest.mark.gpu
def test_two_gpus_replicated_alternative(tmp_path):
    """The measured alternative: particles bit-identical (the push does not communicate), counts exact,
    running average equal to 1e-12 (cross-rank sums are not in id order) and identical on both ranks."""
    need_gpus(2)
    path = str(tmp_path / "res.npz")
    mp.spawn(dh.gpu_worker_replicated, args=(2, free_port(), FRAMES, path), nprocs=2, join=True)
    got = np.load(path)
    o = dh.single_oracle(FRAMES)
    assert_same(got["ids"], np.arange(o.n, dtype=got["ids"].dtype), "every particle exactly once")
    assert_same(got["pos"], o.getPosition(), "position")
    assert_same(got["vel"], o.getVelocity(), "velocity")
    assert_same(got["cnt"], o.cell_count, "cell counts")
    ref = o.moments01_avg
    ok = ~np.isnan(ref)
    assert np.abs(got["avg"][ok] - ref[ok]).max() <= 1e-12 * np.abs(ref[ok]).max()
    assert_same(got["avg"], got["avg_other"], "both ranks hold the same reduced field")
