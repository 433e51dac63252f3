"""N>1 host logic on CPU: world_size-2 gloo run of the batch sharding + the path's single all-reduce."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hesic_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_forward(pair_ids):
    """Deterministic stand-in for one rank's forward: per-pair partial sums (6 values each)."""
    g = torch.Generator().manual_seed(7)
    table = torch.rand(64, sharding.N_PARTIALS, generator=g, dtype=torch.float64)
    table[:, :4] = -1000.0 * table[:, :4]          # log2-likelihood sums are negative
    return table[pair_ids].sum(0)


def _worker(rank, world, port, n_pairs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.shard_range(n_pairs, rank, world)
    part = _fake_forward(torch.arange(b, e))
    sharding.reduce_partials(part)
    if rank == 0:
        torch.save(part, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [16, 7])
def test_two_rank_gloo_reduce_matches_single_process(tmp_path, n_pairs):
    out = str(tmp_path / "p.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_pairs, out), nprocs=2, join=True)
    got = torch.load(out)
    want = _fake_forward(torch.arange(n_pairs))
    assert torch.allclose(got, want, rtol=0, atol=1e-9)
    m = sharding.metrics_from_partials(got, n_pairs, 512, 512)
    pix = n_pairs * 512 * 512
    assert math.isclose(m["bpp"], -float(want[:4].sum()) / pix, rel_tol=1e-12)
    assert math.isclose(m["bpp1"] + m["bpp2"], m["bpp"], rel_tol=1e-12)
    assert math.isclose(m["psnr2"], 10 * math.log10(3 * pix / float(want[5])), rel_tol=1e-12)


def test_shard_range_partitions_every_batch():
    for n in (0, 1, 5, 16, 128, 131):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_reduce_partials_validates_and_is_identity_without_a_group():
    p = torch.arange(6, dtype=torch.float64)
    assert torch.equal(sharding.reduce_partials(p.clone()), p)
    with pytest.raises(ValueError):
        sharding.reduce_partials(torch.zeros(5, dtype=torch.float64))
    with pytest.raises(ValueError):
        sharding.reduce_partials(torch.zeros(6))
