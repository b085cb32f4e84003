"""Column-sharded SpMM across the GPUs of one box (BASELINE config 5).

The feature axis shards naturally: every output column depends on one column of B only, for every
reduce op.  The CSR is replicated; rank r owns columns [r*n_local, (r+1)*n_local) of B and of C.  The
output panel is then made available on every rank, in one of two ways:

  mode="peer"  (default when the ranks can map each other's memory): ONE kernel does both — the SpMM
               epilogue stores each finished n_local-wide row segment straight into every rank's
               row-major C[M, n_total] at column offset r*n_local through NVLink-mapped peer pointers
               (CUDA IPC), so the transfer overlaps the remaining row segments; one stream-ordered
               NCCL all-reduce of a single int afterwards is the completion barrier.
  mode="mcast" (tried first when the ranks share an NVSwitch): the same fused epilogue, but C lives in torch symmetric
               memory and every finished row segment is stored ONCE to the NVLS multicast address
               (multimem.st, dgs_spmm_csr_mcast): the switch replicates it into every rank's C, so a rank sends
               M*n_local*4 bytes per step instead of (world-1) times that.  The end-of-step barrier is ours too: one
               multimem.red on an arrival counter in symmetric memory + a spin on the local copy (dgs_mcast_barrier),
               instead of a one-int NCCL all-reduce.
  mode="nccl"  the baseline: local SpMM, then one ncclAllGather of the [M, n_local] panel into a
               panel-major [world, M, n_local] buffer (panels_to_row_major() permutes when needed).

No reduction is involved, so every rank's panel is bit-identical to the single-GPU kernel run on that panel.  (One single-GPU
launch over the FULL width is a different launch: the kernel cuts the nnz stream into segments according to the launch's
number of column panels, and a row cut by a segment boundary is then summed in a different order — equal to fp32 round-off
for sum / mean, still bit-identical for max / min.  tests/mgpu_worker.py checks both statements.)

Output buffering (peer / mcast): remote ranks store into this rank's C, and the only cross-rank synchronisation is the
completion barrier AFTER the stores.  C is therefore DOUBLE-BUFFERED and alternates per call: step k+1 writes the buffer
step k did not use, so a fast rank's step-k+1 stores cannot overwrite the step-k result a slow rank's consumer kernels are
still reading; by the time a buffer is written again (step k+2) the writer has passed step k+1's barrier, which every rank
enqueues on its stream behind its step-k consumers.  Contract: consume the returned tensor on the stream the op was called
on (or order other streams behind it) and before the call after next.
One process per GPU (torchrun); torch.distributed is used for the rendezvous, the handle exchange
and the barrier only.
"""
import ctypes
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_columns(n_total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, equal column panels; n_total must divide evenly (panels stay vector-aligned)."""
    if world < 1 or n_total % world != 0:
        raise ValueError(f"feature width {n_total} does not split evenly over {world} ranks")
    n = n_total // world
    return [(r * n, (r + 1) * n) for r in range(world)]


def panels_to_row_major(panels: torch.Tensor) -> torch.Tensor:
    """[world, M, n_local] (what an allgather of row panels yields) -> row-major [M, world*n_local]."""
    w, m, n = panels.shape
    return panels.permute(1, 0, 2).reshape(m, w * n)


def exchange_objects(obj, group=None) -> list:
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


class ColumnShardedSpMM:
    def __init__(self, rowptr, col, values, n_local: int, reduce: int = 0, compute: int = 2, group=None,
                 mode: Optional[str] = None):
        from . import _lib
        self._lib = _lib
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rowptr, self.col, self.values = rowptr, col, values
        self.M, self.nnz = rowptr.numel() - 1, col.numel()
        self.n_local, self.n_total = n_local, n_local * self.world
        self.lo, self.hi = shard_columns(self.n_total, self.world)[self.rank]
        self.reduce, self.compute = reduce, compute
        dev = col.device
        self.ws = torch.empty(max(256, _lib.lib.dgs_spmm_workspace_bytes(n_local, self.nnz, 0)), dtype=torch.uint8, device=dev)
        self.C = None          # [2, M, n_total]: double-buffered output (see the module docstring), index = call parity
        self._calls = 0
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._opened = []
        self.mode = mode or ("mcast" if self.world > 1 else "local")
        if self.world > 1 and self.mode == "mcast":
            try:
                self._map_multicast(dev)
            except Exception as e:  # no NVLS / symmetric memory on this box: every rank must agree on the fallback
                self._mcast_error = str(e)
                self.mode = "peer"
            agreed = exchange_objects(self.mode, group)
            if any(m != "mcast" for m in agreed):
                self.mode = "peer"
        if self.C is None or self.mode != "mcast":
            nbuf = 2 if self.world > 1 and self.mode == "peer" else 1
            self.C = torch.empty((nbuf, self.M, self.n_total), dtype=torch.float32, device=dev)
        if self.world > 1 and self.mode == "peer":
            try:
                self._map_peers()
            except Exception as e:  # no IPC on this box: every rank must agree on the fallback
                self._peer_error = str(e)
                self.mode = "nccl"
            agreed = exchange_objects(self.mode, group)
            if any(m != "peer" for m in agreed):
                self.mode = "nccl"
        if self.mode == "nccl":
            self.C_local = torch.empty((self.M, n_local), dtype=torch.float32, device=dev)
            self.panels = torch.empty((self.world, self.M, n_local), dtype=torch.float32, device=dev)

    def _map_multicast(self, dev):
        """C in torch symmetric memory (plumbing: allocation + rendezvous); the kernel only needs the multicast address."""
        import torch.distributed._symmetric_memory as symm
        group = self.group if self.group is not None else dist.group.WORLD
        C = symm.empty((2, self.M, self.n_total), dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(C, group)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if mc == 0:
            raise RuntimeError("symmetric memory has no multicast mapping on this box")
        self.C, self._symm = C, hdl
        self._mc_dst = mc + self.lo * 4
        # arrival counter of the end-of-step barrier (dgs_mcast_barrier): its own symmetric allocation, zeroed everywhere first
        self._bar = symm.empty((64,), dtype=torch.int32, device=dev)
        self._bar.zero_()
        bh = symm.rendezvous(self._bar, group)
        self._bar_mc = int(getattr(bh, "multicast_ptr", 0) or 0)
        self._bar_hdl = bh
        self._epoch = 0
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        if self._bar_mc == 0 or os.environ.get("DGS_MCAST_BARRIER") == "nccl":
            self._bar_mc = 0

    def _map_peers(self):
        L = self._lib
        handle = ctypes.create_string_buffer(64)
        off = ctypes.c_int64(0)
        L.check(L.lib.dgs_ipc_export(self.C.data_ptr(), handle, ctypes.byref(off)), "dgs_ipc_export")
        infos = exchange_objects((bytes(handle.raw), int(off.value)), self.group)
        ptrs = []
        for r, (h, o) in enumerate(infos):
            if r == self.rank:
                base = self.C.data_ptr()
            else:
                p = ctypes.c_void_p()
                L.check(L.lib.dgs_ipc_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)), "dgs_ipc_open")
                self._opened.append(p.value)
                base = p.value + o
            ptrs.append(base + self.lo * 4)          # every destination gets THIS rank's column panel
        # dst[0] must be the local C
        ptrs = [ptrs[self.rank]] + [p for r, p in enumerate(ptrs) if r != self.rank]
        buf_bytes = self.M * self.n_total * 4
        self._dst = [(ctypes.c_void_p * len(ptrs))(*[p + b * buf_bytes for p in ptrs]) for b in range(2)]

    def __call__(self, B_local: torch.Tensor) -> torch.Tensor:
        """B_local: [K, n_local] fp32 (this rank's column panel).  Returns row-major C[M, n_total] (peer /
        local mode) or the panel-major [world, M, n_local] gather (nccl mode)."""
        L = self._lib
        lib, ptr = L.lib, L.ptr
        stream = torch.cuda.current_stream(B_local.device).cuda_stream
        if self.mode == "nccl":
            L.check(lib.dgs_spmm_csr(self.M, self.n_local, self.nnz, ptr(self.rowptr), ptr(self.col), ptr(self.values),
                                     ptr(B_local), B_local.stride(0), ptr(self.C_local), self.n_local, None, 0,
                                     self.reduce, self.compute, ptr(self.ws), self.ws.numel(), stream), "dgs_spmm_csr")
            dist.all_gather_into_tensor(self.panels, self.C_local, group=self.group)
            return self.panels
        buf = self._calls & 1 if self.C.shape[0] == 2 else 0
        self._calls += 1
        if self.mode == "mcast":
            L.check(lib.dgs_spmm_csr_mcast(self.M, self.n_local, self.nnz, ptr(self.rowptr), ptr(self.col), ptr(self.values),
                                           ptr(B_local), B_local.stride(0), self._mc_dst + buf * self.M * self.n_total * 4,
                                           self.n_total, self.reduce,
                                           self.compute, ptr(self.ws), self.ws.numel(), stream), "dgs_spmm_csr_mcast")
            self.barrier(stream)                            # completion barrier: every rank's multicast stores have landed
            return self.C[buf]
        if self.mode == "local":
            dst = (ctypes.c_void_p * 1)(self.C.data_ptr())
        else:
            dst = self._dst[buf]
        L.check(lib.dgs_spmm_csr_multi(self.M, self.n_local, self.nnz, ptr(self.rowptr), ptr(self.col), ptr(self.values),
                                       ptr(B_local), B_local.stride(0), len(dst), dst, self.n_total, self.reduce,
                                       self.compute, ptr(self.ws), self.ws.numel(), stream), "dgs_spmm_csr_multi")
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)   # completion barrier: all peers' stores have landed
        return self.C[buf]

    def barrier(self, stream=None):
        """The end-of-step barrier alone (stream-ordered): dgs_mcast_barrier in mcast mode, else a one-int NCCL all-reduce."""
        if self.world == 1:
            return
        if self.mode == "mcast" and getattr(self, "_bar_mc", 0):
            if stream is None:
                stream = torch.cuda.current_stream(self._bar.device).cuda_stream
            self._epoch += 1
            L = self._lib
            L.check(L.lib.dgs_mcast_barrier(self._bar_mc, self._bar.data_ptr(), (self._epoch * self.world) & 0xffffffff, stream),
                    "dgs_mcast_barrier")
        else:
            dist.all_reduce(self._flag, group=self.group)

    def close(self):
        for p in self._opened:
            self._lib.lib.dgs_ipc_close(p)
        self._opened = []


def nnz_chunks(nnz: int, world: int):
    """Equal, padded slices of the nonzero stream: (chunk, [(lo, hi)] per rank with hi clipped to nnz)."""
    chunk = (nnz + world - 1) // world
    return chunk, [(min(nnz, r * chunk), min(nnz, (r + 1) * chunk)) for r in range(world)]


class EdgeShardedSDDMM:
    """SDDMM across the GPUs of one box by splitting the EDGE stream: the pattern and both dense
    operands are replicated, rank r computes the dots of edges [r*chunk, (r+1)*chunk) with the single-GPU kernel (COO
    form: the rows of a slice come from the expanded row array, so no rank searches the whole rowptr), and one NCCL
    all-gather of the [chunk] result slices (4 B per edge) gives every rank the full [1, nnz] output.  Every dot is
    computed by exactly one rank with the single-GPU arithmetic, so the result is bit-identical to one GPU — unlike
    the alternative split along K, whose all-reduce changes the summation order."""

    def __init__(self, rowptr, col, group=None, _slice_kernel=None):
        # _slice_kernel(row, col, D1, D2, out): test hook (tests/test_distributed_cpu.py drives the sharding / exchange logic
        # over gloo with a stand-in); the product always runs the CUDA kernel, which raises on non-CUDA tensors
        if _slice_kernel is None:
            from . import _kernels
            _slice_kernel = lambda row, col, D1, D2, out: _kernels.sddmm_coo(row, col, D1, D2, out=out)
        self._slice_kernel = _slice_kernel
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.M, self.nnz = rowptr.numel() - 1, col.numel()
        self.chunk, spans = nnz_chunks(self.nnz, self.world)
        self.lo, self.hi = spans[self.rank]
        deg = (rowptr[1:] - rowptr[:-1]).long()
        row = torch.repeat_interleave(torch.arange(self.M, dtype=torch.int32, device=col.device), deg)
        self.row = row[self.lo:self.hi].contiguous()
        self.col = col[self.lo:self.hi].contiguous()
        self.out = torch.zeros(self.chunk * self.world, dtype=torch.float32, device=col.device)

    def __call__(self, D1: torch.Tensor, D2: torch.Tensor) -> torch.Tensor:
        """D1 [M, K], D2 [ncols, K] fp32, replicated.  Returns [1, nnz] (the torch-face shape) on every rank."""
        mine = self.out[self.rank * self.chunk:(self.rank + 1) * self.chunk]
        self._slice_kernel(self.row, self.col, D1, D2, mine[:self.hi - self.lo])
        if self.world > 1:
            dist.all_gather_into_tensor(self.out, mine, group=self.group)   # in place: the slice is already at its offset
        return self.out[:self.nnz].reshape(1, self.nnz)


class HostColumnShardedSpMM:
    """The column-sharded SpMM for HOST-resident operands (the multi-GPU counterpart of dgs_spmm_csr_host).

    Every rank needs the whole CSR, but the CSR needs to cross PCIe only once per BOX: rank r uploads the r-th
    1/world slice of col / val (plus rowptr and its own B panel), one NCCL all-gather per array replicates the slices
    over NVLink (~15x the bandwidth of one PCIe link), then the fused peer-store SpMM runs and the rank's own C panel
    returns to the host.  Per rank and step at world = 8 on the reddit-like config: 182 MB over PCIe instead of 977 MB.
    """

    def __init__(self, M: int, nnz: int, n_local: int, has_value: bool, device, reduce: int = 0, compute: int = 2,
                 group=None, mode: Optional[str] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.M, self.nnz, self.n_local = M, nnz, n_local
        self.chunk, spans = nnz_chunks(nnz, self.world)
        self.lo_nnz, self.hi_nnz = spans[self.rank]
        padded = self.chunk * self.world
        self.d_rowptr = torch.empty(M + 1, dtype=torch.int32, device=device)
        self.d_col = torch.zeros(padded, dtype=torch.int32, device=device)
        self.d_val = torch.zeros(padded, dtype=torch.float32, device=device) if has_value else None
        self.d_B = None
        self.op = ColumnShardedSpMM(self.d_rowptr, self.d_col[:nnz], self.d_val[:nnz] if has_value else None, n_local,
                                    reduce=reduce, compute=compute, group=group, mode=mode)
        self.h2d_bytes = 4 * (M + 1) + (self.hi_nnz - self.lo_nnz) * (8 if has_value else 4)
        self.d2h_bytes = 4 * M * n_local

    def __call__(self, h_rowptr, h_col, h_val, h_B_local, h_C_local):
        """All arguments are host tensors (pinned for asynchronous copies); h_C_local [M, n_local] receives this
        rank's panel.  Synchronises before returning."""
        lo, hi, r = self.lo_nnz, self.hi_nnz, self.rank
        self.d_rowptr.copy_(h_rowptr, non_blocking=True)
        mine_c = self.d_col[r * self.chunk:(r + 1) * self.chunk]
        mine_c[:hi - lo].copy_(h_col[lo:hi], non_blocking=True)
        if self.d_val is not None:
            mine_v = self.d_val[r * self.chunk:(r + 1) * self.chunk]
            mine_v[:hi - lo].copy_(h_val[lo:hi], non_blocking=True)
        if self.d_B is None:
            self.d_B = torch.empty(h_B_local.shape, dtype=torch.float32, device=self.d_col.device)
        self.d_B.copy_(h_B_local, non_blocking=True)
        if self.world > 1:   # in-place all-gather: every rank's slice already sits at its final offset
            dist.all_gather_into_tensor(self.d_col, mine_c, group=self.group)
            if self.d_val is not None:
                dist.all_gather_into_tensor(self.d_val, mine_v, group=self.group)
        out = self.op(self.d_B)
        if self.op.mode == "nccl":
            panel = out[self.rank]
        else:
            panel = out[:, self.op.lo:self.op.hi]
        h_C_local.copy_(panel, non_blocking=True)
        torch.cuda.current_stream(self.d_col.device).synchronize()
        return h_C_local

    def close(self):
        self.op.close()
