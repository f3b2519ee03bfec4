"""Multi-GPU sharding of a database (host-side logic; no data-path collective exists on this path).

Every (query, target) pair is independent, so N GPUs each search a shard and the
host merges by index.  Shards are balanced by residue count -- cells = Q x residues, so residue balance is
work balance -- by dealing the length-sorted sequences round-robin, which also gives every shard the same
length mix (each GPU's longest target is about as long as the others').
"""
from __future__ import annotations

import numpy as np

from .capi import SequenceDB


def deal_shards(lengths, world_size):
    """Index arrays (into the caller's database) of the `world_size` shards."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    return [np.sort(order[r::world_size]) for r in range(world_size)]


def shard_db(db: SequenceDB, index):
    """The sub-database holding `index` (kept in caller order)."""
    index = np.asarray(index, dtype=np.int64)
    lens = db.lengths[index].astype(np.int64)
    offsets = np.zeros(len(index) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    residues = np.empty(int(offsets[-1]), dtype=np.uint8)
    for k, i in enumerate(index):
        residues[offsets[k]:offsets[k + 1]] = db.sequence(int(i))
    return SequenceDB(residues, offsets)


def merge_results(n, shard_indices, shard_results):
    """Scatter per-shard result arrays (any dtype, including the OpalSearchResult record dtype) back
    into one array of n entries in caller order."""
    out = np.zeros(n, dtype=shard_results[0].dtype)
    for idx, res in zip(shard_indices, shard_results):
        out[np.asarray(idx, dtype=np.int64)] = res
    return out
