"""Multi-GPU parity (needs >= 2 GPUs on the box, skipped otherwise): one process per GPU, NCCL.

* batch sharding: every rank runs the layer on its slice of the batch, no collective in the data path;
  the concatenation equals the single-GPU result bit for bit;
* bank sharding: every rank correlates against its slice of the bank columns, ONE all-reduce MAX of the packed
  int64 (score, ~index) keys over NCCL, then the blend / paste run replicated; equals the unsharded result.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port):
    import torch.distributed as dist
    from deepinpainting_b200 import shift_ops
    from deepinpainting_b200.sharding import allreduce_max_keys, shard_bank, shard_batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        gen = torch.Generator().manual_seed(11)                    # same data on every rank
        B, C, H = max(4, world), 256, 32                        # at least one image per rank
        N = H * H
        x = torch.randn(B, C, H, H, generator=gen).to(dev)
        ref = (torch.relu(torch.randn(B, C, H, H, generator=gen)) * 3).to(dev)
        g = torch.randn(B, C, H, H, generator=gen).to(dev)
        flag = torch.zeros(H, H, dtype=torch.int64)
        flag[8:24, 6:20] = 1
        mi = shift_ops.mask_index_from_flag(flag.view(-1), dev)
        for mode in ("tensor", "exact"):
            full_out, full_saved = shift_ops.shift_forward(x, ref, mi, need_grad=True, mode=mode)
            full_gin = shift_ops.shift_backward(g, full_saved, 1.0)
            # ---- batch sharding: no collective; gather only to compare ----
            b0, b1 = shard_batch(B, world, rank)
            out_s, saved_s = shift_ops.shift_forward(x[b0:b1].contiguous(), ref[b0:b1].contiguous(), mi, need_grad=True, mode=mode)
            gin_s = shift_ops.shift_backward(g[b0:b1].contiguous(), saved_s, 1.0)
            assert torch.equal(out_s, full_out[b0:b1]) and torch.equal(gin_s, full_gin[b0:b1])
            assert torch.equal(saved_s.ind, full_saved.ind[b0:b1])
            # ---- bank sharding: one NCCL all-reduce MAX of int64 keys ----
            cb, ce = shard_bank(N, world, rank)
            out_k, saved_k = shift_ops.shift_forward_sharded(x, ref, mi, cb, ce, allreduce_max_keys, need_grad=True, mode=mode)
            gin_k = shift_ops.shift_backward(g, saved_k, 1.0)
            torch.cuda.synchronize()
            assert torch.equal(saved_k.ind, full_saved.ind), mode
            assert torch.equal(out_k, full_out) and torch.equal(gin_k, full_gin), mode
        # ---- patch mode (BASELINE.json configs[3]: 3 x 3 patches, bank sharded, NCCL all-reduce MAX of the keys) ----
        # both routes: rows of C*9 <= 1024 values through the 1 x 1 pipeline on patch maps, wider rows through the
        # wide kernels; exact mode (P = nH*nW is not a multiple of 128), column shards of any size
        for (Bp, Cp, Hp) in ((2, 64, 16), (1, 128, 12)):
            xp = torch.randn(Bp, Cp, Hp, Hp, generator=gen).abs().to(dev)
            rp = (torch.relu(torch.randn(Bp, Cp, Hp, Hp, generator=gen)) * 3).to(dev)
            feat = torch.zeros(Hp, Hp, dtype=torch.uint8, device=dev)
            feat[4:9, 3:10] = 1
            mip = shift_ops.build_flags(feat, 3, 1, 1)
            P = mip.flag.numel()
            full_out, full_ind = shift_ops.shift_forward_patches(xp, rp, mip, 3, 1, mode="exact")
            cb, ce = shard_bank(P, world, rank, align=1)
            out_k, ind_k = shift_ops.shift_forward_patches(xp, rp, mip, 3, 1, mode="exact", col_begin=cb, col_end=ce,
                                                           reduce_max=allreduce_max_keys)
            torch.cuda.synchronize()
            assert torch.equal(ind_k, full_ind) and torch.equal(out_k, full_out), (Cp, Hp)
        # ---- more ranks than 128-column tiles: the surplus ranks own an EMPTY shard, contribute identity keys and must not
        # fall out of the collective (16 x 16 map: 2 tiles) ----
        Hs = 16
        xs = torch.randn(2, 64, Hs, Hs, generator=gen).to(dev)
        rs = (torch.relu(torch.randn(2, 64, Hs, Hs, generator=gen)) * 3).to(dev)
        fs = torch.zeros(Hs, Hs, dtype=torch.int64)
        fs[4:12, 3:11] = 1
        mis = shift_ops.mask_index_from_flag(fs.view(-1), dev)
        full_out, full_saved = shift_ops.shift_forward(xs, rs, mis, need_grad=False, mode="tensor")
        cb, ce = shard_bank(Hs * Hs, world, rank)
        out_k, saved_k = shift_ops.shift_forward_sharded(xs, rs, mis, cb, ce, allreduce_max_keys, need_grad=False, mode="tensor")
        torch.cuda.synchronize()
        assert torch.equal(saved_k.ind, full_saved.ind) and torch.equal(out_k, full_out), (rank, cb, ce)
        # ---- long patch rows on the tensor route (K = 1152, P = 324 padded to 3 tiles of 128 columns) ----
        Bp, Cp, Hp = 1, 128, 20
        xp = torch.randn(Bp, Cp, Hp, Hp, generator=gen).abs().to(dev)
        rp = (torch.relu(torch.randn(Bp, Cp, Hp, Hp, generator=gen)) * 3).to(dev)
        feat = torch.zeros(Hp, Hp, dtype=torch.uint8, device=dev)
        feat[5:12, 4:13] = 1
        mip = shift_ops.build_flags(feat, 3, 1, 1)
        P = mip.flag.numel()
        full_out, full_ind = shift_ops.shift_forward_patches(xp, rp, mip, 3, 1, mode="tensor")
        cb, ce = shard_bank(-(-P // 128) * 128, world, rank)
        cb, ce = min(cb, P), min(ce, P)
        out_k, ind_k = shift_ops.shift_forward_patches(xp, rp, mip, 3, 1, mode="tensor", col_begin=cb, col_end=ce,
                                                       reduce_max=allreduce_max_keys)
        torch.cuda.synchronize()
        assert torch.equal(ind_k, full_ind) and torch.equal(out_k, full_out), (rank, cb, ce)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_batch_and_bank_sharding_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    mp.spawn(_worker, args=(world, _free_port()), nprocs=world, join=True)
