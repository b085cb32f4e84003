"""Worker for tests/test_multigpu_gpu.py (launched under torchrun, one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dgsparse-lib_b200")):
    sys.path.insert(0, p)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    import dgsparse._kernels as K
    import dgsparse._lib as L
    from dgsparse.distributed import ColumnShardedSpMM, panels_to_row_major
    from tools import graphs
    M, n_local = 20000, 64
    rowptr, col = graphs.random_csr(M, M, 600000, 31, empty_frac=0.2, hub=2)
    val = graphs.uniform(col.size, 1)
    Bfull = graphs.uniform(M * n_local * world, 2, -1, 1).reshape(M, n_local * world)
    rp, cc, vv = (torch.from_numpy(a).to(dev) for a in (rowptr, col, val))
    Bf = torch.from_numpy(Bfull).to(dev)
    # The oracle of the exchange: the single-GPU kernel on every rank's column panel, side by side.  No reduction crosses ranks,
    # so the sharded result must equal it BIT FOR BIT.  (The single-GPU kernel on the full width is a different launch: how
    # the nnz stream is cut into segments depends on the number of column panels of the launch, and a row cut by a segment
    # boundary is summed in a different order — equal to fp32 round-off for sum, bit-identical for max, checked below.)
    def per_panel(Bmat, reduce=L.SUM):
        return torch.cat([K.spmm(rp, cc, vv, Bmat[:, r * n_local:(r + 1) * n_local].contiguous(), reduce, L.MUL)
                          for r in range(world)], dim=1)
    ref = per_panel(Bf)
    B_local = Bf[:, rank * n_local:(rank + 1) * n_local].contiguous()
    for reduce in (L.SUM, L.MAX):
        refr = per_panel(Bf, reduce)
        full = K.spmm(rp, cc, vv, Bf, reduce, L.MUL)               # one launch over the full width
        if reduce == L.MAX:
            assert torch.equal(full, refr)
        else:
            assert torch.allclose(full, refr, rtol=1e-5, atol=1e-5)
        for mode in ("mcast", "peer", "nccl"):
            op = ColumnShardedSpMM(rp, cc, vv, n_local, reduce=reduce, mode=mode)
            for it in range(3):
                out = op(B_local)
            torch.cuda.synchronize()
            C = out if op.mode in ("peer", "mcast") else panels_to_row_major(out)
            ok = torch.equal(C, refr)                              # no reduction across ranks: bit-identical
            flag = torch.tensor([int(ok)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"mode requested={mode} used={op.mode} reduce={reduce} identical={bool(flag.item())}", flush=True)
            assert flag.item() == 1, (mode, reduce)
            if mode == "peer" and rank == 0 and op.mode != "peer":
                print("PEER MAPPING UNAVAILABLE:", getattr(op, "_peer_error", "?"), flush=True)
            if mode == "mcast" and rank == 0 and op.mode != "mcast":
                print("MULTICAST UNAVAILABLE:", getattr(op, "_mcast_error", "?"), flush=True)
            dist.barrier()
            op.close()
    # write-after-read across ranks: a different B every step, a SLOW consumer on rank 0 (it sleeps on the stream
    # before it reads the step's C) and fast producers elsewhere.  The output is double-buffered and every step ends in a
    # barrier, so step k+1's remote stores must never land in the buffer rank 0 is still reading for step k.
    for mode in ("mcast", "peer"):
        op = ColumnShardedSpMM(rp, cc, vv, n_local, reduce=L.SUM, mode=mode)
        steps = 6
        stash = []
        for k in range(steps):
            Bk = (Bf + float(k)).contiguous()
            out = op(Bk[:, rank * n_local:(rank + 1) * n_local].contiguous())
            if rank == 0:
                torch.cuda._sleep(30_000_000)                      # ~15 ms: the consumer lags, the other ranks run ahead
            C = out if op.mode in ("peer", "mcast") else panels_to_row_major(out)
            stash.append(C.clone())
        torch.cuda.synchronize()
        ok = all(torch.equal(stash[k], per_panel((Bf + float(k)).contiguous())) for k in range(steps))
        flag = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"slow-consumer stress mode={op.mode} every step identical={bool(flag.item())}", flush=True)
        assert flag.item() == 1, mode
        dist.barrier()
        op.close()
    # host-resident operands: sliced upload + NVLink all-gather of the CSR, own panel back to the host
    from dgsparse.distributed import HostColumnShardedSpMM
    hp = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_rp, h_cc, h_vv = hp(rowptr), hp(col), hp(val)
    h_B = hp(Bfull[:, rank * n_local:(rank + 1) * n_local])
    h_C = torch.empty(M, n_local, dtype=torch.float32).pin_memory()
    hop = HostColumnShardedSpMM(M, int(col.size), n_local, True, dev)
    for it in range(2):
        h_C.fill_(float("nan"))
        hop(h_rp, h_cc, h_vv, h_B, h_C)
    ok = torch.equal(h_C, ref[:, rank * n_local:(rank + 1) * n_local].cpu())
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"host-sharded path identical={bool(flag.item())} mode={hop.op.mode}", flush=True)
    assert flag.item() == 1
    hop.close()
    dist.barrier()
    # edge-sharded SDDMM: every rank ends up with the single-GPU result, bit for bit
    from dgsparse.distributed import EdgeShardedSDDMM
    for Kd in (32, 256):
        D1 = torch.from_numpy(graphs.uniform(M * Kd, 5, -1, 1).reshape(M, Kd)).to(dev)
        D2 = torch.from_numpy(graphs.uniform(M * Kd, 6, -1, 1).reshape(M, Kd)).to(dev)
        want = K.sddmm_csr(rp, cc, D1, D2)
        sd = EdgeShardedSDDMM(rp, cc)
        for it in range(2):
            got = sd(D1, D2)
        torch.cuda.synchronize()
        flag = torch.tensor([int(torch.equal(got, want))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"edge-sharded sddmm K={Kd} identical={bool(flag.item())}", flush=True)
        assert flag.item() == 1, Kd
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK", flush=True)


if __name__ == "__main__":
    main()
