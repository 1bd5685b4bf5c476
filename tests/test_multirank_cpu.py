"""World-size-2 test of the multi-GPU host logic on CPU (gloo): member sharding and the single end-of-run
gather.  The per-member compute is stood in for by the CPU oracle (the checker), because the product's
compute path is CUDA-only; what is under test is that every member lands in the right place exactly once and
that results do not depend on how the ensemble is cut (members are independent)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from flow_sim_b200.ensemble import gather_members, gather_packed, shard_bounds, shard_members


def test_shard_bounds_cover_the_ensemble_exactly_once():
    for total in (1, 5, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 4, 4)
    for layout in ("strided", "block"):
        for total, world in ((5, 2), (65536, 8), (7, 3)):
            idx = np.concatenate([shard_members(total, r, world, layout) for r in range(world)])
            assert sorted(idx.tolist()) == list(range(total))
    assert shard_members(10, 1, 4).tolist() == [1, 5, 9]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, layout, out_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle_py

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    flat = util.golden_inputs("gerd_calib_m0")
    n_all = np.linspace(0.02, 0.06, total)
    mine = shard_members(total, rank, world, layout)
    flat.member_n_main = n_all[mine]
    h, q, _ = oracle_py.gvf(flat, flat.meta["initial_flow"], flat.meta["downstream_depth"], n_members=len(mine))
    flat.ic_depth, flat.ic_flow = h, q
    res = oracle_py.run(flat, n_members=len(mine), out_mode=1)
    lv, rm = oracle_py.objective(flat.n_levels, res["flow"], res["depth"], flat.meta["z0"],
                                 [1562.5, 3850, 6000, 10000, 14000, 21000], [497.5, 500, 502, 505, 507, 510])
    rmse = gather_members(torch.from_numpy(rm), total, rank, world, layout)
    iters = gather_members(torch.from_numpy(res["iters"].astype(np.int32)), total, rank, world, layout)
    # the strong-scaling gather of bench.py: RMSE + iteration counts + status + upstream series as ONE message
    local = dict(rmse=torch.from_numpy(rm), iters=torch.from_numpy(res["iters"].astype(np.int32)),
                 status=torch.from_numpy(res["status"].astype(np.int32)), depth=torch.from_numpy(res["depth"]),
                 flow=torch.from_numpy(res["flow"]))
    packed = gather_packed(local, total, rank, world, layout)
    if rank == 0:
        np.savez(out_path, rmse=rmse.numpy(), iters=iters.numpy(), p_rmse=packed["rmse"].numpy(), p_iters=packed["iters"].numpy(),
                 p_status=packed["status"].numpy(), p_depth=packed["depth"].numpy(), p_flow=packed["flow"].numpy(),
                 p_bytes=packed["bytes_per_member"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total,layout", [(4, "strided"), (5, "strided"), (5, "block")])
def test_two_rank_gather_equals_single_process(tmp_path, total, layout):
    import oracle_py

    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), total, layout, out), nprocs=2, join=True)
    got = np.load(out)
    flat = util.golden_inputs("gerd_calib_m0")
    flat.member_n_main = np.linspace(0.02, 0.06, total)
    h, q, _ = oracle_py.gvf(flat, flat.meta["initial_flow"], flat.meta["downstream_depth"], n_members=total)
    flat.ic_depth, flat.ic_flow = h, q
    res = oracle_py.run(flat, n_members=total, out_mode=1)
    _, rm = oracle_py.objective(flat.n_levels, res["flow"], res["depth"], flat.meta["z0"],
                                [1562.5, 3850, 6000, 10000, 14000, 21000], [497.5, 500, 502, 505, 507, 510])
    assert np.array_equal(got["rmse"], rm)            # bit-identical: sharding must not change any member
    assert np.array_equal(got["iters"], res["iters"])
    # the packed gather carries the same bytes: 8 + 4 (L-1) + 4 + 16 L per member
    L = flat.n_levels
    assert int(got["p_bytes"]) == 8 + 4 * (L - 1) + 4 + 16 * L == 668
    assert np.array_equal(got["p_rmse"], rm) and np.array_equal(got["p_iters"], res["iters"])
    assert np.array_equal(got["p_status"], res["status"])
    assert np.array_equal(got["p_depth"], res["depth"]) and np.array_equal(got["p_flow"], res["flow"])
