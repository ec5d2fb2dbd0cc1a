"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Reads shard by sample or by file chunk of a sample across ranks; after the local trim+collapse
every (unique sequence, sample) count is sent to the sequence's owner rank (hash partition) with
ONE variable-size all-to-all (X1), so mapping and tallying run on disjoint sets with the libraries
replicated; the small integer tables are then summed with an all-reduce (X2).  The reference has
no counterpart (samples are processed serially, src/mirge/__main__.py:346).  Pure tensor plumbing:
works on NCCL (GPU) and gloo (CPU tests) alike.

Host round trips per exchange: one -- the G x G record counts (the split sizes of the all-to-all
have to be Python integers).  Everything else is queued on the engine's stream.
"""
import numpy as np
import torch
import torch.distributed as dist


def exchange_counts(send_counts, group=None):
    """send_counts: int64 tensor [world] (records bound for each rank).  Returns (send, recv) as Python lists:
    the counts all-to-all followed by the exchange's single device-to-host copy."""
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    both = torch.stack([send_counts, recv_counts]).cpu()
    return [int(x) for x in both[0]], [int(x) for x in both[1]]


def exchange_records(send_buf, part_counts, record_bytes, group=None):
    """send_buf: uint8 tensor of records grouped by destination rank; part_counts[r]: number of records for
    rank r (tensor, list or array).  Returns (recv_buf, n_recv)."""
    dev = send_buf.device
    if not torch.is_tensor(part_counts):
        part_counts = torch.as_tensor(np.asarray(part_counts, dtype=np.int64), device=dev)
    sc_h, rc_h = exchange_counts(part_counts.to(dev), group)
    n_recv = int(sum(rc_h))
    recv = torch.empty(max(n_recv, 1) * record_bytes, dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv[:n_recv * record_bytes], send_buf[:sum(sc_h) * record_bytes],
                           output_split_sizes=[c * record_bytes for c in rc_h],
                           input_split_sizes=[c * record_bytes for c in sc_h], group=group)
    return recv, n_recv


def reduce_tables(tables, device, group=None):
    """Sum dict-of-numpy integer tables over ranks (every rank gets the result); host tables, one all-reduce."""
    keys = list(tables)
    flat = np.concatenate([np.ascontiguousarray(tables[k]).astype(np.int64).ravel() for k in keys]) if keys else np.zeros(0, np.int64)
    t = torch.as_tensor(flat, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    flat = t.cpu().numpy()
    out, off = {}, 0
    for k in keys:
        n = int(np.asarray(tables[k]).size)
        out[k] = flat[off:off + n].reshape(np.asarray(tables[k]).shape)
        off += n
    return out


def _scratch(engine, name, numel, dtype, dev):
    cache = engine.__dict__.setdefault('_dist_bufs', {})
    t = cache.get(name)
    if t is None or t.numel() < numel or t.dtype != dtype or t.device != dev:
        t = torch.empty(numel, dtype=dtype, device=dev)
        cache[name] = t
    return t


def hash_partition(engine, group=None, cap_records=None):
    """X1: re-home every unique sequence of `engine` on its owner rank.  The engine's stream must be torch's
    current stream.  Returns the number of records received (a sequence seen by k ranks / samples arrives k times)."""
    world = dist.get_world_size(group)
    dev = torch.device('cuda', torch.cuda.current_device())
    rb = engine.record_bytes()
    cap = int(cap_records or engine.max_unique)
    send = _scratch(engine, 'send', cap * rb, torch.uint8, dev)
    counts = _scratch(engine, 'counts', world, torch.int64, dev)
    engine.export_records_device(world, send.data_ptr(), cap, counts.data_ptr())
    recv, n_recv = exchange_records(send, counts[:world], rb, group)
    engine.reset_unique()
    engine.import_records_device(recv.data_ptr(), n_recv)
    engine.__dict__['_dist_recv'] = recv        # keep the buffer alive until the import has run
    return n_recv


def reduce_engine_tables(engine, group=None):
    """X2 on the device: mir_quant, mir_iscan, lib_totals, len_hist, kept and processed of every rank are summed in
    place with one all-reduce (after tally); export_tables()/stats() then return the job-wide tables on every rank."""
    dev = torch.device('cuda', torch.cuda.current_device())
    n = engine.tables_words()
    t = _scratch(engine, 'tables', n, torch.int64, dev)
    engine.pack_tables_device(t.data_ptr())
    dist.all_reduce(t[:n], op=dist.ReduceOp.SUM, group=group)
    engine.unpack_tables_device(t.data_ptr())
