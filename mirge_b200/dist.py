"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Reads shard by file chunk / sample across ranks; after the local trim+collapse
every unique sequence is sent to its owner rank (hash partition) with ONE
variable-size all-to-all, so mapping and tallying run on disjoint sets with the
libraries replicated; the small count tables are then summed with an
all-reduce.  The reference has no counterpart (samples are processed serially,
src/mirge/__main__.py:346).  Pure tensor plumbing: works on NCCL (GPU) and gloo
(CPU tests) alike.
"""
import numpy as np
import torch
import torch.distributed as dist


def exchange_records(send_buf, part_counts, record_bytes, group=None):
    """send_buf: uint8 tensor of records grouped by destination rank;
    part_counts[r]: number of records for rank r.  Returns (recv_buf, n_recv)."""
    world = dist.get_world_size(group)
    dev = send_buf.device
    sc = torch.as_tensor(np.asarray(part_counts, dtype=np.int64), device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    rc_h = rc.cpu().tolist()
    sc_h = [int(x) for x in part_counts]
    n_recv = int(sum(rc_h))
    recv = torch.empty(max(n_recv, 1) * record_bytes, dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv[:n_recv * record_bytes], send_buf[:sum(sc_h) * record_bytes],
                           output_split_sizes=[c * record_bytes for c in rc_h],
                           input_split_sizes=[c * record_bytes for c in sc_h], group=group)
    return recv, n_recv


def reduce_tables(tables, device, group=None):
    """Sum dict-of-numpy integer tables over ranks (every rank gets the result)."""
    out = {}
    for k, v in tables.items():
        t = torch.as_tensor(np.ascontiguousarray(v).astype(np.int64), device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        out[k] = t.cpu().numpy()
    return out


def hash_partition(engine, stream=None, group=None):
    """Re-home every unique sequence of `engine` on its owner rank."""
    world = dist.get_world_size(group)
    _, _, U = engine.stats()
    rb = engine.record_bytes()
    dev = torch.device('cuda', torch.cuda.current_device())
    send = torch.empty(max(U, 1) * rb, dtype=torch.uint8, device=dev)
    pc = engine.export_records_device(world, send.data_ptr(), max(U, 1))
    recv, n_recv = exchange_records(send, pc, rb, group)
    torch.cuda.current_stream().synchronize()
    engine.reset_unique()
    engine.import_records_device(recv.data_ptr(), n_recv)
    _, _, n_owned = engine.stats()      # records of one sequence from several ranks merge here
    return n_owned
