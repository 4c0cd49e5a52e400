"""
dist.py -- multi-GPU layer: one process per GPU, block-row sharding of the n x n Gram.

Every Gram entry depends on two sequences only, so each rank holds ALL packed sequences (6.4 MB at n = 200 000),
builds whatever feature matrix it needs locally and computes its own block-row: the construction needs no
collective at all.  torch.distributed (NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU tests) is used
only where the path has a real exchange:

  * gather_rows      all-gather of block-rows when a consumer (the reference's solvers) needs the whole matrix on
                     one device -- capacity limited (fp64: n <~ 140 000 per 180 GB GPU);
  * center_sharded   centring of a sharded Gram: row sums are local (a block-row has every column); the column sums
                     and the grand sum are one all-reduce of n + 1 doubles;
  * frobenius_sharded  ALIGNF's <K_i, K_j>_F style reductions: one all-reduce of a scalar per pair.

  * SymmetricShards  the mirror K[j,i] = K[i,j] (kernels.py:45) across GPUs: each rank computes only half of its
                     block-row and the GEMM epilogue stores every tile twice, into its own buffer and -- transposed, by
                     TMA bulk stores -- into staging that the copy engine ships to the owner over NVLink peer memory
                     (CUDA IPC) piece by piece, or straight into the owner's buffer ("direct"); no collective, one
                     barrier at the end.

The partition / collective logic is backend agnostic (it is what the gloo tests cover); the arithmetic is supplied
by the caller: `kmg.device` functions on the GPU, oracle functions in the CPU tests.
"""
import torch
import torch.distributed as dist

TILE_ALIGN = 256  # block-row boundaries are multiples of the GEMM tile height


def block_rows(n, world, rank, align=TILE_ALIGN):
    """Contiguous block-row [r0, r1) of rank `rank`: ceil(n / world) rounded up to `align`, last ranks may be short or empty."""
    per = -(-n // world)
    per = -(-per // align) * align
    r0 = min(n, rank * per)
    r1 = min(n, r0 + per)
    return r0, r1


def all_block_rows(n, world, align=TILE_ALIGN):
    return [block_rows(n, world, r, align) for r in range(world)]


def build_block_row(build_fn, n, group=None, align=TILE_ALIGN):
    """Run `build_fn(r0, r1) -> tensor (r1-r0, n)` for this rank's block-row.  No communication."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    r0, r1 = block_rows(n, world, rank, align)
    return r0, r1, build_fn(r0, r1)


def gather_rows(block, n, group=None, align=TILE_ALIGN):
    """All-gather the block-rows into the full (n, n) matrix on every rank (block: (r1-r0, n))."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return block
    world = dist.get_world_size(group)
    spans = all_block_rows(n, world, align)
    per = max(r1 - r0 for r0, r1 in spans)
    pad = torch.zeros((per, block.shape[1]), dtype=block.dtype, device=block.device)
    pad[: block.shape[0]] = block
    out = torch.empty((world * per, block.shape[1]), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(r1 - r0 == per for r0, r1 in spans):
        return out[:n]
    return torch.cat([out[r * per: r * per + (r1 - r0)] for r, (r0, r1) in enumerate(spans)], dim=0)


def center_sharded(block, n, row_sum_fn, col_sum_fn, apply_fn, group=None):
    """center_K (kernels.py:387-395) on a sharded Gram.  block: this rank's (rows, n) block-row.
    row_sum_fn(block) -> (rows,), col_sum_fn(block) -> (n,) partial column sums of the block,
    apply_fn(block, rs, cs, g, n) -> centred block  (K - cs/n - rs/n + g/n^2)."""
    rs = row_sum_fn(block)
    cs = col_sum_fn(block)
    packed = torch.cat([cs, rs.sum().reshape(1)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return apply_fn(block, rs, packed[:-1].contiguous(), packed[-1:].contiguous(), n)


def frobenius_sharded(block_a, block_b, dot_fn, group=None):
    """<A, B>_F of two identically sharded matrices: local partial + one scalar all-reduce."""
    part = dot_fn(block_a, block_b).reshape(1)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part


# ---------------------------------------------------------------------------------------------------
# GPU conveniences (thin: they only bind kmg.device to the functions above)
# ---------------------------------------------------------------------------------------------------
def spectrum_block_row(planes, L, ks, n, group=None, out_dtype=1):
    """This rank's block-row of the (summed) spectrum Gram, device resident.  Phi is built locally."""
    from . import device as kd
    phi = kd.spectrum_phi(planes, L, ks)

    def build(r0, r1):
        return kd.gram_i8(phi[r0:r1], phi, row_index0=r0, col_index0=0, out_dtype=out_dtype)
    return build_block_row(build, n, group)


def wd_block_row(planes, L, d, n, group=None):
    from . import device as kd
    return build_block_row(lambda r0, r1: kd.wd_block(planes[r0:r1], planes, L, d, row_index0=r0), n, group)


def wds_block_row(planes, L, d, S, n, group=None):
    from . import device as kd
    return build_block_row(lambda r0, r1: kd.wds_block(planes[r0:r1], planes, L, d, S, row_index0=r0), n, group)


def la_block_row(planes, L, e, d, beta, n, smith=0, group=None, align=8):
    """This rank's block-row of the local-alignment Gram (BASELINE configs[4]).  K[i, j] is evaluated with x = the sequence
    of smaller index (kernels.py:289-291), which the kernel derives from the global row / column indices, so block-rows
    built on different ranks are the rows of the one symmetric matrix.  One warp handles four pairs: no tile alignment."""
    from . import device as kd
    return build_block_row(lambda r0, r1: kd.la_block(planes[r0:r1], planes, L, e, d, beta, smith, row_index0=r0), n, group, align)


def mismatch_block_row(planes, L, k, m, n, normalize=True, group=None):
    from . import device as kd
    sd = kd.mismatch_diag_sqrt(planes, L, k, m) if normalize else None  # every rank computes all n diagonals: no collective
    if sd is not None and n > 0 and float(sd[0].item()) == 1.0:
        sd = None  # normalize_K's early-out (kernels.py:404-405): a raw K[0,0] of exactly 1 leaves the Gram unnormalised

    def build(r0, r1):
        return kd.mismatch_block(planes[r0:r1], planes, L, k, m, row_index0=r0,
                                 sd_rows=None if sd is None else sd[r0:r1], sd_cols=sd)
    return build_block_row(build, n, group)


def center_block_row(block, n, group=None):
    from . import device as kd
    return center_sharded(block, n, kd.row_sums, kd.col_sums, kd.center_apply, group)


# ------------------------------------------------------------------------------------------------
# Sharded SYMMETRIC build: half the tensor-core work per GPU, mirror stores over peer memory
# ------------------------------------------------------------------------------------------------
def sym_bounds(n, world, align=TILE_ALIGN):
    """Boundaries of `world` non-empty block-rows, multiples of `align` (except the last, = n), balanced to one tile."""
    tiles = -(-n // align)
    if tiles < world:
        raise ValueError(f"n = {n} is too small for {world} block-rows of at least {align} rows")
    return [min(n, align * (p * tiles // world)) for p in range(world)] + [n]


def sym_takes(bounds, a, b, I, J):
    """The assignment rule (kmg_gram_sharded_takes_host): does part a compute tile (I, J) whose columns part b owns?"""
    import ctypes as C
    import numpy as np
    from . import _cabi
    bd = np.ascontiguousarray(bounds, np.int64)
    return _cabi.check(_cabi.lib().kmg_gram_sharded_takes_host(len(bounds) - 1, bd.ctypes.data_as(C.c_void_p), a, b, I, J))


class SymmetricShards:
    """Block-row buffers of one n x n fp64 (or s32) Gram, one per rank of a single node, each visible to every other rank
    through CUDA IPC.  `build_spectrum` fills them with the sharded symmetric GEMM."""

    def __init__(self, n, dtype=torch.float64, group=None, exchange=None, ldo=None):
        """n: size of the symmetric square the ranks share; ldo >= n: row stride (and width) of every block-row buffer --
        columns [n, ldo) are the caller's (e.g. the rest of a wider block-row, filled by a plain cross-Gram launch).
        exchange: "staged" (default: transposed pieces into local staging, one copy-engine peer copy per piece while the
        next launch runs), "direct" (the GEMM epilogue's TMA stores write the owner's buffer over NVLink: compute and
        exchange in one kernel) or "single" (one launch, thread-issued peer stores); KMG_SYM_EXCHANGE overrides the
        default.  Measured on 2 GPUs, n = 100 000 (profiles/r2_exchange_modes.txt): staged 32.8 ms, direct 39.7 ms, single
        40.0 ms against 29.9 ms with no link at all -- SM-issued peer stores, by threads or by the TMA engine alike, slow
        the GEMM itself although a store-only kernel reaches 717 GB/s on the same link (profiles/r2_p2p_store_probe.txt)."""
        import ctypes as C
        import os
        self.exchange = exchange or os.environ.get("KMG_SYM_EXCHANGE", "staged")
        assert self.exchange in ("direct", "staged", "single")
        from . import _cabi
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n, self.dtype = n, dtype
        self.ldo = n if ldo is None else int(ldo)
        assert self.ldo >= n
        self.bounds = sym_bounds(n, self.world)
        self.r0, self.r1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        esz = 8 if dtype == torch.float64 else 4
        lib = _cabi.lib()
        from . import device as kd
        self._own, self._stage = C.c_void_p(), C.c_void_p()
        self.ptrs = [None] * self.world
        self._opened = []
        # Every step that can fail (allocation, IPC export / open) is followed by an exchange of the outcome, so a rank
        # that fails never leaves the others waiting inside a collective: all ranks raise together.
        payload, err = None, None
        try:
            _cabi.check(lib.kmg_dev_malloc((self.r1 - self.r0) * self.ldo * esz, C.byref(self._own)))
            nbytes = kd.sharded_stage_bytes(self.bounds, self.rank, 1 if esz == 8 else 0) if self.exchange == "staged" else 0
            if nbytes:  # local staging for the transposed blocks that the copy engine ships to their owners
                _cabi.check(lib.kmg_dev_malloc(nbytes, C.byref(self._stage)))
            if self.world > 1:
                handle = (C.c_uint8 * 64)()
                _cabi.check(lib.kmg_ipc_export(self._own, handle))
                payload = bytes(handle)
        except Exception as exc:  # noqa: BLE001
            err = exc
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, payload, group=group)
            if err is None and all(h is not None for h in handles):
                try:
                    for r, h in enumerate(handles):
                        if r == self.rank:
                            continue
                        p = C.c_void_p()
                        _cabi.check(lib.kmg_ipc_open((C.c_uint8 * 64).from_buffer_copy(h), C.byref(p)))
                        self.ptrs[r] = p.value
                        self._opened.append(p)
                except Exception as exc:  # noqa: BLE001
                    err = exc
            elif err is None:
                err = RuntimeError("another rank could not allocate or export its block-row buffer")
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None, group=group)
            if err is None and not all(oks):
                err = RuntimeError("another rank could not map the peer buffers")
        if err is not None:
            self._release()
            raise err
        self.ptrs[self.rank] = self._own.value
        self.launches = kd.sharded_launches(self.bounds, self.rank, self.exchange)
        # this rank's block-row as a tensor (no copy): torch reads the CUDA array interface
        holder = type("_Buf", (), {})()
        holder.__cuda_array_interface__ = {"shape": (self.r1 - self.r0, self.ldo), "typestr": "<f8" if esz == 8 else "<i4",
                                           "data": (self._own.value, False), "version": 3}
        self._holder = holder
        self.block = torch.as_tensor(holder, device=torch.device("cuda", torch.cuda.current_device()))

    def build_spectrum(self, phi, sd=None, defer_join=False):
        """All ranks call this with their local copy of Phi (n x W int8).  Returns entries this rank issued to the MMA.
        defer_join: do not make the stream wait for the outgoing peer copies yet -- the caller launches more work (the
        plain remainder of a wider block-row) and then calls join(), so that work runs under the copies."""
        from . import device as kd
        assert phi.shape[0] == self.n
        computed = kd.gram_i8_sharded(phi, self.bounds, self.rank, self.ptrs, self.ldo,
                                      out_dtype=1 if self.dtype == torch.float64 else 0, sd=sd,
                                      stage=self._stage.value if self.exchange == "staged" else None,
                                      exchange=self.exchange if (self.exchange != "staged" or self._stage.value) else "single",
                                      defer_join=defer_join)
        return computed

    def join(self):
        """The stream waits for this rank's outgoing peer copies (after build_spectrum(defer_join=True))."""
        from . import device as kd
        if self.exchange == "staged" and self._stage.value:
            kd.sharded_join()

    def needed_row_ranges(self):
        """Rows of Phi this rank's launches read: its own block and the column blocks at cyclic distance 1 .. world/2 (the
        part of the block at distance world/2 it computes), merged into ascending (lo, hi) ranges.  The rows of the other
        blocks only enter tiles that other ranks compute and deliver, so this rank need not build their features."""
        g, a, b = self.world, self.rank, self.bounds
        need = [(b[a], b[a + 1])]
        for d in range(1, g):
            if 2 * d > g:
                break
            p = (a + d) % g
            lo, hi = b[p], b[p + 1]
            if 2 * d == g and a > p:  # the higher-numbered part computes the columns from the split on
                tiles = -(-(b[p + 1] - b[p]) // TILE_ALIGN)
                lo = min(b[p] + (tiles + 1) // 2 * TILE_ALIGN, b[p + 1])
            if hi > lo:
                need.append((lo, hi))
        need.sort()
        merged = [list(need[0])]
        for lo, hi in need[1:]:
            if lo <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], hi)
            else:
                merged.append([lo, hi])
        return [tuple(r) for r in merged]

    def finish(self):
        """Every buffer is complete once every rank's launch has finished: stream sync, then a barrier."""
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)

    def _release(self):
        from . import _cabi
        lib = _cabi.lib()
        for p in self._opened:
            lib.kmg_ipc_close(p)
        self._opened = []
        for name in ("_own", "_stage"):
            p = getattr(self, name, None)
            if p is not None and p.value:
                lib.kmg_dev_free(p)
            setattr(self, name, None)

    def close(self):
        self.block = None
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)  # nobody still writes into a buffer that is about to be unmapped
        self._release()
