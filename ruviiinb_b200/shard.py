"""How a problem is split over ranks (one process per GPU).

Genes are sharded in contiguous row ranges for every per-gene quantity (alpha, gmean, Mb, psi, log-likelihood);
the control-gene rows are replicated so that the per-cell W update can shard by cell instead.  These helpers are
pure Python (no CUDA) so that the N > 1 host logic is testable with the gloo backend on CPU."""
import numpy as np


def gene_range(G, world, rank):
    """Rows [g0, g1) of rank `rank`: ceil(G / world) rows per rank (the layout the library's all-gather of alpha expects), the last
    rank may hold fewer.  A rank without rows cannot hold a context (`ruvnb_set_design` refuses g0 >= g1): that is an error here."""
    per = (G + world - 1) // world
    g0, g1 = min(G, rank * per), min(G, (rank + 1) * per)
    if g0 >= g1:
        raise ValueError(f"{G} genes in shards of {per} rows leave rank {rank} of {world} without a row: use at most {(G + per - 1) // per} ranks")
    return g0, g1


def chunk_range(nchunks, world, rank):
    """128-cell chunks [c0, c1) whose cells rank `rank` updates in the W step (csrc/irls.cu: irls_body)."""
    per = (nchunks + world - 1) // world
    return min(nchunks, rank * per), min(nchunks, (rank + 1) * per)


def local_counts(Y, ctl, world, rank):
    """(Y_local, Y_ctl, (g0, g1)) for `ruvnb_set_design(..., g0, g1)` + `ruvnb_upload_counts_dense(Y_local, ..., Y_ctl)`."""
    g0, g1 = gene_range(Y.shape[0], world, rank)
    Yctl = Y[np.asarray(ctl, bool)] if world > 1 else None
    return Y[g0:g1], Yctl, (g0, g1)


def broadcast_unique_id(dist, lib, device=None):
    """NCCL unique id created on rank 0 (`ruvnb_comm_unique_id`) and broadcast with torch.distributed."""
    import ctypes as C
    import torch
    buf = (C.c_ubyte * 128)()
    if dist.get_rank() == 0:
        rc = lib.ruvnb_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise RuntimeError(f"ruvnb_comm_unique_id failed with {rc}")
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=device)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())
