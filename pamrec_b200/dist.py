"""Host-side arithmetic of the data-parallel / row-sharded layout (one process per GPU).

No reference counterpart: the reference runs one tf.Session on one device.  The rules:

* listwise groups of 5 rows (pamrec.py:73-75) are dealt round-robin: rank r takes groups g with g % W == r of every
  GLOBAL batch, so the union over ranks is exactly the batch a single process would have trained on;
* row ``i`` of an embedding table lives on rank ``i % W`` at local row ``i // W`` (round-robin spreads the hot head of a
  Zipf id distribution; every shard has ``ceil(rows / W)`` rows, the tail zero-padded);
* the device plan sorts lookups by ``owner * rows_per_shard + local_row`` (csrc/kernels_shard.cu), mirrored here by
  ``exchange_plan`` for the CPU tests.

Everything here is numpy on the host; the device path is csrc/api.cu:shard_exchange_fwd.
"""
import numpy as np

GROUP = 5
SEQ_FIELDS = ("item_history", "item_cate_history", "item_loop_times_history", "mask", "item_duration_history",
              "item_satisfied_history", "item_play_history")


def shard_rows(vocab_rows, world):
    return (int(vocab_rows) + world - 1) // world


def owner_of(ids, world):
    return np.asarray(ids) % world


def local_row(ids, world):
    return np.asarray(ids) // world


def shard_table(full, world, rank):
    """Rows ``rank, rank + W, ...`` of ``full`` as a [shard_rows, width] array (tail zero-padded)."""
    full = np.asarray(full)
    rps = shard_rows(full.shape[0], world)
    out = np.zeros((rps,) + full.shape[1:], full.dtype)
    mine = full[rank::world]
    out[:mine.shape[0]] = mine
    return out


def unshard_table(shards, vocab_rows):
    """Inverse of shard_table: ``shards[r]`` is rank r's array."""
    world = len(shards)
    first = np.asarray(shards[0])
    out = np.zeros((vocab_rows,) + first.shape[1:], first.dtype)
    for r, sh in enumerate(shards):
        n = len(range(r, vocab_rows, world))
        out[r::world] = np.asarray(sh)[:n]
    return out


def group_rows(n_rows, world, rank, group=GROUP):
    """Row indices of a global batch of ``n_rows`` rows (whole groups) that rank ``rank`` trains on."""
    if n_rows % group:
        raise ValueError(f"global batch {n_rows} is not a multiple of {group}")
    g = np.arange(rank, n_rows // group, world)
    return (g[:, None] * group + np.arange(group)[None, :]).reshape(-1)


def split_feed(feed, world, rank, grouped=True):
    """This rank's share of a global feed dict (io/sequential_iterator.py:1155-1175 layout, string keys).

    Training batches keep listwise groups intact (grouped=True); scoring batches are dealt row by row.
    Returns (local feed, global rows).  Every array whose first dimension is the batch is sliced."""
    n = int(np.asarray(feed["items"]).shape[0])
    rows = group_rows(n, world, rank) if grouped else np.arange(rank, n, world)
    out = {}
    for k, v in feed.items():
        a = np.asarray(v)
        out[k] = a[rows] if a.ndim >= 1 and a.shape[0] == n else v
    return out, n


def exchange_plan(ids, world, rps):
    """numpy mirror of the device plan: unique sharded-table addresses of ``ids`` grouped by owner.

    Returns (local_rows_by_owner: list of int arrays, inverse: position -> index into the concatenated unique list)."""
    ids = np.asarray(ids, np.int64).reshape(-1)
    keys = (ids % world) * rps + ids // world
    ukeys, inverse = np.unique(keys, return_inverse=True)
    owners = ukeys // rps
    per_owner = [(ukeys[owners == o] - o * rps).astype(np.int32) for o in range(world)]
    return per_owner, inverse.astype(np.int32)


# ---------------------------------------------------------------------------------------------- replicated tables
REPLICATE_BYTES = 64 * 2 ** 20     # tables="auto" on several GPUs: replicate while the four tables together stay this small


def choose_tables(tables, world, n_users, n_items, n_cates, limit_bytes=REPLICATE_BYTES):
    """Table placement of an engine (Engine.__init__): "local" on one GPU; on several, "replicated" (every rank holds whole
    tables, the merged row gradients ride the dense all-reduce) while item + category + 2 user tables fit `limit_bytes`, else
    "sharded" (row r on rank r % world, rows / row gradients exchanged)."""
    if tables in (None, "auto"):
        table_bytes = 4 * (int(n_items) * 16 + int(n_cates) * 4 + 2 * int(n_users) * 20)
        tables = "local" if world == 1 else ("replicated" if table_bytes <= limit_bytes else "sharded")
    if tables == "replicated" and world == 1:
        tables = "local"
    if tables not in ("local", "replicated", "sharded"):
        raise ValueError(f"unknown table placement {tables!r}")
    if world > 1 and tables == "local":
        raise ValueError("world_size > 1 needs tables='replicated' or 'sharded'")
    return tables


def replicated_contribution(ids, grads, vocab):
    """numpy mirror of what the run walk leaves in a rank's exchange buffer for one replicated table (kernels_sparse2.cu, mode
    SP2_DENSE): the rank's lookups merged into a DENSE [vocab, width] gradient table plus one touch mark per looked-up row."""
    ids = np.asarray(ids).reshape(-1)
    g = np.zeros((vocab, grads.shape[-1]), np.float64)
    np.add.at(g, ids, np.asarray(grads, np.float64).reshape(ids.size, -1))
    touch = np.zeros(vocab, np.float64)
    touch[np.unique(ids)] = 1.0
    return g, touch


def two_shot_slices(n, world):
    """[lo, hi) of the slice of an n-element contribution that rank r sums in the peer-memory all-reduce (kernels_p2p.cu:k_xr_*);
    n is padded to a multiple of 64 * world so that every slice is a whole number of 16-byte vectors."""
    q = 64 * world
    cap = (n + q - 1) // q * q
    step = cap // world
    return cap, [(r * step, (r + 1) * step) for r in range(world)]
