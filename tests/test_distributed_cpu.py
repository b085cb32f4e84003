"""CPU, world_size 2 over gloo: the host-side logic of the column-sharded path (panel arithmetic, the
object exchange used for IPC handles, panel-major -> row-major layout).  The kernels themselves need
a GPU (tests -m gpu / bench.py --gpus N)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, "dgsparse-lib_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dgsparse.distributed import shard_columns, panels_to_row_major, exchange_objects
    n_total, M = 8, 5
    lo, hi = shard_columns(n_total, world)[rank]
    full = torch.arange(M * n_total, dtype=torch.float32).reshape(M, n_total)
    mine = full[:, lo:hi].contiguous()                     # what this rank's SpMM would produce
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok = torch.equal(panels_to_row_major(torch.stack(gathered)), full)
    infos = exchange_objects((bytes([rank]) * 64, rank * 256))
    ok = ok and [o for _, o in infos] == [r * 256 for r in range(world)] and all(len(h) == 64 for h, _ in infos)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_column_shard_host_logic_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_nnz_chunks_cover_the_stream():
    sys.path.insert(0, os.path.join(ROOT, "dgsparse-lib_b200"))
    from dgsparse.distributed import nnz_chunks
    for nnz, world in ((114615892, 8), (10, 4), (3, 8), (0, 2), (16, 4)):
        chunk, spans = nnz_chunks(nnz, world)
        assert chunk * world >= nnz and len(spans) == world
        assert spans[0][0] == 0 and spans[-1][1] == nnz
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and all(hi - lo <= chunk for lo, hi in spans)


def test_shard_columns_rejects_ragged_split():
    sys.path.insert(0, os.path.join(ROOT, "dgsparse-lib_b200"))
    from dgsparse.distributed import shard_columns
    assert shard_columns(512, 8) == [(64 * r, 64 * (r + 1)) for r in range(8)]
    with pytest.raises(ValueError):
        shard_columns(100, 8)


def _sddmm_worker(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, "dgsparse-lib_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dgsparse.distributed import EdgeShardedSDDMM
    g = torch.Generator().manual_seed(5)
    M, K = 37, 8
    deg = torch.randint(0, 6, (M,), generator=g)
    deg[3] = 0; deg[4] = 0; deg[11] = 40                     # empty rows and a hub that straddles the shard boundary
    rowptr = torch.zeros(M + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(deg, 0).int()
    nnz = int(rowptr[-1])
    col = torch.randint(0, M, (nnz,), generator=g).int()
    D1, D2 = torch.rand(M, K, generator=g), torch.rand(M, K, generator=g)

    def stand_in(row, c, A, B, out):                           # what the CUDA slice kernel computes, in torch on the CPU
        out.copy_((A[row.long()] * B[c.long()]).sum(1))
    op = EdgeShardedSDDMM(rowptr, col, _slice_kernel=stand_in)
    got = op(D1, D2)
    row_full = torch.repeat_interleave(torch.arange(M), deg)
    want = (D1[row_full] * D2[col.long()]).sum(1).reshape(1, nnz)
    spans_ok = op.lo == min(nnz, rank * op.chunk) and op.hi == min(nnz, (rank + 1) * op.chunk)
    ret[rank] = bool(torch.equal(got, want)) and tuple(got.shape) == (1, nnz) and spans_ok
    dist.destroy_process_group()


def test_edge_sharded_sddmm_exchange_logic_world2():
    """Edge split, in-place all-gather of the slices and the [1, nnz] assembly of EdgeShardedSDDMM over gloo (the slice kernel
    itself is CUDA-only and is checked on the GPU: tests/test_sddmm_csr2csc_gpu.py, tests/mgpu_worker.py)."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31000 + os.getpid() % 2000
    mp.spawn(_sddmm_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
