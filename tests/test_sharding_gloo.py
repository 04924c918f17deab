"""CPU, world_size 2 (gloo): the batch x heads partition covers every unit exactly once, and
per-rank results of the (oracle-evaluated) hot path reassemble to the single-process result -
i.e. the path needs no data-path collective.  The product kernels are GPU-only; here the oracle
stands in as the per-unit function, which is legitimate for checking the partition logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mx_quantization_b200.sharding import gather_for_verification, shard_batch_heads, shard_bounds, take_shard


def test_shard_bounds_cover_exactly():
    for n in (1, 2, 7, 16, 255, 256):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                lo, hi = shard_bounds(n, w, r)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    assert shard_batch_heads(256, 12, 8, 3) == ("batch", 96, 128)
    assert shard_batch_heads(1, 16, 8, 7) == ("heads", 14, 16)      # long-sequence sweep: B < G
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, B, H, q, k, v, full_out, full_mask):
    from oracle import mxint8_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        qs, ks, vs = (take_shard(t, B, H, world, rank) for t in (q, k, v))
        r = O.pruned_attention(qs, ks, vs, 6, integer_scores=True)
        masks = O.idx_to_mask_words(r["idx"], q.shape[2])
        kind, lo, hi = shard_batch_heads(B, H, world, rank)
        outs = gather_for_verification(r["out"].reshape(hi - lo, -1), world)
        mks = gather_for_verification(masks.reshape(hi - lo, -1), world)
        if rank == 0:
            got = torch.cat(outs, 0).reshape(full_out.shape if kind == "batch" else (-1,) + full_out.shape[2:])
            assert torch.equal(got.reshape(full_out.shape), full_out)
            assert torch.equal(torch.cat(mks, 0).reshape(full_mask.shape), full_mask)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,H", [(4, 2), (1, 4), (3, 2)])
def test_two_rank_sharding_reassembles(B, H):
    from oracle import mxint8_oracle as O
    g = torch.Generator().manual_seed(5)
    q, k, v = (torch.randn(B, H, 24, 64, generator=g) for _ in range(3))
    ref = O.pruned_attention(q, k, v, 6, integer_scores=True)
    full_mask = O.idx_to_mask_words(ref["idx"], 24)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, B, H, q, k, v, ref["out"], full_mask), nprocs=2, join=True)
