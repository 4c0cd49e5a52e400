"""GPU: the multi-GPU layer (kmg/dist.py) on real devices.  Single-process checks always run; the NCCL checks need
>= 2 GPUs on the box (one process per GPU) and compare the sharded build bit-for-bit with the single-GPU build."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle_c as oc
import oracle_np as onp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available(), "these tests need a GPU"
    return torch


def test_block_row_builders_single_process(torch_cuda, dna):
    from kmg import device as kd
    from kmg import dist as kdist
    codes, _ = dna
    c = codes[:600]
    planes = kd.pack(c, 0)
    r0, r1, blk = kdist.spectrum_block_row(planes, 101, [1, 2, 3, 4], 600)
    assert (r0, r1) == (0, 600) and np.array_equal(blk.cpu().numpy(), oc.spectrum_block(c, c, [1, 2, 3, 4]))
    _, _, blk = kdist.wd_block_row(planes, 101, 7, 600)
    wd = oc.wd_block(c, c, 7)
    assert np.array_equal(blk.cpu().numpy(), wd)
    cen = kdist.center_block_row(blk, 600).cpu().numpy()
    assert np.abs(cen - onp.center_K(wd)).max() <= 1e-12 * np.abs(wd).max()
    _, _, blk = kdist.mismatch_block_row(planes[:200], 101, 10, 1, 200)
    want = onp.normalize_K(oc.mismatch_raw_block(c[:200], c[:200], 10, 1).astype(np.float64))
    assert np.array_equal(blk.cpu().numpy(), want)
    _, _, blk = kdist.la_block_row(planes[:64], 101, -11, -1, 0.5, 64)
    want = oc.la_block(c[:64], c[:64], -11, -1, 0.5, 0)
    assert np.all(np.abs(blk.cpu().numpy() - want) <= 1e-12 * np.abs(want))
    _, _, blk = kdist.wds_block_row(planes[:48], 101, 4, 2, 48)
    assert np.array_equal(blk.cpu().numpy(), onp.wds_gram(c[:48], 4, 2))
    # sharded pieces against the fused single-device centring
    sub = kd.wd_block(planes[128:384], planes, 101, 7, row_index0=128)
    rs, cs = kd.row_sums(sub), kd.col_sums(sub)
    assert np.allclose(rs.cpu().numpy(), wd[128:384].sum(1), rtol=1e-13) and np.allclose(cs.cpu().numpy(), wd[128:384].sum(0), rtol=1e-13)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q):
    for p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import oracle_np as onp
        from kmg import device as kd
        from kmg import dist as kdist
        n = 3000
        codes = onp.synthetic_codes(n, 101, seed=3)
        planes = kd.pack(codes, 0)
        ks = list(range(1, 8))
        r0, r1, blk = kdist.spectrum_block_row(planes, 101, ks, n)
        phi = kd.spectrum_phi(planes, 101, ks)
        single = kd.gram_i8(phi, phi, symmetric=True)
        assert torch.equal(blk, single[r0:r1])                      # sharded block == rows of the 1-GPU Gram
        full = kdist.gather_rows(blk, n)
        assert torch.equal(full, single)                            # NCCL all-gather materialises it everywhere
        # sharded SYMMETRIC build: half the MMA work per rank, mirror stores into the peers' buffers over CUDA IPC
        # TMA stores into peer memory / peer copies of staged blocks / thread-issued stores into peer memory
        for staged in ("direct", "staged", "single"):
            shards = kdist.SymmetricShards(n, exchange=staged)
            for _ in range(2):
                shards.block.fill_(-1.0)
                torch.cuda.synchronize()
                dist.barrier()
                computed = shards.build_spectrum(phi)
                shards.finish()
                assert torch.equal(shards.block, single[shards.r0:shards.r1]), staged
                assert computed < 0.75 * (shards.r1 - shards.r0) * n
            shards.close()
        lr0, lr1, lblk = kdist.la_block_row(planes[:512], 101, 11, 1, 0.5, 512)   # rows of ONE symmetric matrix across ranks
        lfull = kd.la_block(planes[:512], planes[:512], 101, 11, 1, 0.5, 0, symmetric=True)
        assert torch.equal(lblk, lfull[lr0:lr1])
        _, _, wblk = kdist.wd_block_row(planes, 101, 10, n)
        wfull = kd.wd_block(planes, planes, 101, 10, symmetric=True)
        cen = kdist.center_block_row(wblk, n)                       # all-reduce of n+1 doubles
        ref = kd.center(wfull)[r0:r1]
        assert float((cen - ref).abs().max()) <= 1e-12 * float(wfull.abs().max())
        q.put((rank, "ok"))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_nccl_sharded_equals_single_gpu(torch_cuda):
    torch = torch_cuda
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = min(world, 4)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
