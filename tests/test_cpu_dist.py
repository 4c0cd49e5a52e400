"""CPU: the multi-GPU host logic (kmg/dist.py) with gloo, world size 2 and 3: block-row partition, all-gather
of ragged block-rows, sharded centring and sharded Frobenius products.  The per-block arithmetic is supplied by the
oracle here (on the GPU box it is kmg.device); what is under test is the partition / collective plumbing."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    for p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_np as onp
        from kmg import dist as kdist
        codes = onp.synthetic_codes(n, 31, seed=9)
        full = onp.spectrum_gram(codes, [1, 2, 3])
        wd = onp.wd_gram(codes, 4)
        # --- block rows: no communication, ragged last block (align 16 so that n=100 splits unevenly)
        r0, r1, blk = kdist.build_block_row(lambda a, b: torch.from_numpy(full[a:b].copy()), n, align=16)
        spans = kdist.all_block_rows(n, world, 16)
        assert spans[rank] == (r0, r1) and spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        # --- all-gather of ragged block rows reproduces the full Gram bit for bit on every rank
        got = kdist.gather_rows(blk, n, align=16)
        assert got.shape == (n, n) and np.array_equal(got.numpy(), full)
        # --- sharded centring == center_K of the full matrix (normwise 1e-12)
        wblk = torch.from_numpy(wd[r0:r1].copy())
        cen = kdist.center_sharded(
            wblk, n,
            row_sum_fn=lambda b: b.sum(1), col_sum_fn=lambda b: b.sum(0),
            apply_fn=lambda b, rs, cs, g, nn: b - cs[None, :] / nn - rs[:, None] / nn + g / (nn * nn))
        want = onp.center_K(wd)[r0:r1]
        assert np.abs(cen.numpy() - want).max() <= 1e-12 * np.abs(wd).max()
        # --- sharded Frobenius product (exact: integer-valued Grams)
        fro = kdist.frobenius_sharded(blk, blk, lambda a, b: (a * b).sum())
        assert float(fro) == float((full * full).sum())
        # --- the shared symmetric build has no CPU path: without a GPU every rank must raise together (the failure is
        # exchanged before anyone enters the next collective), none may hang
        from kmg._cabi import KmgError
        if not torch.cuda.is_available():
            try:
                kdist.SymmetricShards(256 * world)
                raise AssertionError("SymmetricShards must not succeed without a CUDA device")
            except (KmgError, RuntimeError) as exc:
                assert "no CPU fallback" in str(exc) or "another rank" in str(exc), str(exc)
            dist.barrier()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover - reported to the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 100), (3, 100), (2, 32)])
def test_gloo_block_row_sharding(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_block_rows_partition_properties():
    sys.path.insert(0, os.path.join(ROOT, "kernel-methods-for-genomics_b200"))
    from kmg import dist as kdist
    for n in (0, 1, 255, 256, 257, 9000, 200_000):
        for world in (1, 2, 4, 8):
            spans = kdist.all_block_rows(n, world)
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(r0 % 256 == 0 or r0 == n for r0, _ in spans)
    assert kdist.all_block_rows(200_000, 8) == [(i * 25088, min(200_000, (i + 1) * 25088)) for i in range(8)]


def test_symmetric_shard_assignment_covers_every_tile_once():
    """kmg_gram_sharded_takes_host (the rule the sharded symmetric GEMM builds its tile lists from): for every pair of
    tiles (I, J), I != J, exactly one of part(I) computing (I, J) and part(J) computing (J, I) holds; diagonal tiles are
    computed by their owner; every part does about half of its block-row."""
    from kmg import dist as kdist
    for n, world in ((2048, 2), (3000, 2), (3000, 3), (5000, 4), (7000, 5), (200000 // 8, 8), (9000, 8), (4096, 1)):
        bounds = kdist.sym_bounds(n, world)
        assert bounds[0] == 0 and bounds[-1] == n and all(b % 256 == 0 for b in bounds[:-1])
        assert all(bounds[p + 1] > bounds[p] for p in range(world))
        tiles = -(-n // 256)
        owner = [max(p for p in range(world) if bounds[p] <= 256 * t) for t in range(tiles)]
        done = np.zeros((world,), np.int64)
        for I in range(tiles):
            for J in range(tiles):
                t = kdist.sym_takes(bounds, owner[I], owner[J], I, J)
                if I == J:
                    assert t == 1
                elif I < J:
                    assert t + kdist.sym_takes(bounds, owner[J], owner[I], J, I) == 1, (n, world, I, J)
                done[owner[I]] += t
        rows = np.array([sum(1 for t in range(tiles) if owner[t] == p) for p in range(world)])
        share = done / (rows * tiles)
        if world > 1 and tiles >= 8 * world:
            assert share.min() > 0.40 and share.max() < 0.62, (n, world, share)
    with pytest.raises(ValueError):
        kdist.sym_bounds(500, 4)


def test_sharded_plan_matches_the_assignment_rule():
    """The staged sharded build launches one GEMM per piece of a peer block (api.cu sharded_plan); the staging it asks for
    (kmg_gram_sharded_stage_bytes, no GPU needed) must hold exactly the entries the assignment rule gives the part in
    other parts' columns -- the two descriptions of the same partition cannot drift apart."""
    import ctypes as C
    from kmg import _cabi
    from kmg import dist as kdist
    lib = _cabi.lib()
    for n, world in ((2048, 2), (3000, 2), (3000, 3), (5000, 4), (9000, 8), (25000, 8), (7000, 5)):
        bounds = kdist.sym_bounds(n, world)
        bd = np.ascontiguousarray(bounds, np.int64)
        tiles = -(-n // 256)
        owner = [max(p for p in range(world) if bounds[p] <= 256 * t) for t in range(tiles)]
        width = [min(256, n - 256 * t) for t in range(tiles)]
        for a in range(world):
            want = 0
            for I in range(tiles):
                if owner[I] != a:
                    continue
                for J in range(tiles):
                    if owner[J] != a and kdist.sym_takes(bounds, a, owner[J], I, J):
                        want += width[I] * width[J]
            got = C.c_int64(-1)
            _cabi.check(lib.kmg_gram_sharded_stage_bytes(world, bd.ctypes.data_as(C.c_void_p), a, 1, C.byref(got)))
            assert want * 8 <= got.value <= want * 8 + 256 * 2 * world, (n, world, a, want * 8, got.value)
            # launches of one call: 2 column pieces per full block (1 when a block is a single tile wide), the half
            # block at distance world/2, the diagonal block; the single-launch exchange is one launch
            launches = C.c_int(-1)
            _cabi.check(lib.kmg_gram_sharded_launches(world, bd.ctypes.data_as(C.c_void_p), a, _cabi.KMG_EXCH_STAGED, C.byref(launches)))
            full = sum(1 for d in range(1, world) if 2 * d < world)
            assert 1 + full <= launches.value <= 1 + 2 * full + (1 if world % 2 == 0 else 0), (n, world, a, launches.value)
            _cabi.check(lib.kmg_gram_sharded_launches(world, bd.ctypes.data_as(C.c_void_p), a, _cabi.KMG_EXCH_SINGLE, C.byref(launches)))
            assert launches.value == 1  # <= 2 pieces per peer block, each padded to 256 B


def test_needed_phi_rows_cover_exactly_what_the_launches_read():
    """SymmetricShards.needed_row_ranges (the Phi rows a rank builds in bench.py): every row or column tile of every tile the
    assignment rule gives the rank lies inside the ranges, and the ranges hold at most one tile more than that."""
    from kmg import dist as kdist

    class Fake:
        pass
    for n, world in ((200000, 8), (100000, 4), (50000, 2), (9000, 3), (7000, 5), (4096, 1)):
        bounds = kdist.sym_bounds(n, world)
        tiles = -(-n // 256)
        owner = [max(p for p in range(world) if bounds[p] <= 256 * t) for t in range(tiles)]
        for a in range(world):
            f = Fake()
            f.world, f.rank, f.bounds = world, a, bounds
            ranges = kdist.SymmetricShards.needed_row_ranges(f)
            assert all(lo < hi for lo, hi in ranges) and all(r[1] < s[0] for r, s in zip(ranges, ranges[1:]))
            need = np.zeros(tiles, bool)
            for I in range(tiles):
                if owner[I] != a:
                    continue
                for J in range(tiles):
                    if kdist.sym_takes(bounds, a, owner[J], I, J):
                        need[I] = need[J] = True
            got = np.zeros(tiles, bool)
            for lo, hi in ranges:
                got[lo // 256: -(-hi // 256)] = True
            assert np.all(got[need]), (n, world, a, ranges)
            assert got.sum() - need.sum() <= 1, (n, world, a)
