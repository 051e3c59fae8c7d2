"""Collaborative item indexing with the quadratic parts on the GPU.

Drop-in for ref src/src_t5/utils/indexing.py:149-256 `generate_collaborative_id`: same arguments, same item_map.  What moves
to the device (libp5b200.so, csrc/indexing.cu):
  * the item co-occurrence matrix over the training prefixes (ref :163-180: a Python loop over itertools.combinations,
    O(sum_u len_u^2) dict lookups and numpy scalar updates — minutes to hours on Yelp / Taobao);
  * the per-cluster sub-matrix extraction of the BFS (ref :220-231: a Python O(n^2) loop per cluster).
Both are integer counts, so the device results equal the reference's matrices bit for bit (tested).  The spectral
clustering itself (sklearn.cluster.SpectralClustering(assign_labels="cluster_qr", random_state=0, affinity="precomputed"))
is the reference's library call and stays on the host, consuming the device-built matrices.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict
from typing import Dict, List

import numpy as np
import torch

from . import _lib


def cooccurrence_matrix(user_sequence_dict: Dict[str, List[str]], item2id: Dict[str, int], float32: int = 0, device="cuda") -> torch.Tensor:
    """adjacency matrix of ref utils/indexing.py:163-180, built on the device; returns a CUDA tensor [n, n]"""
    lib = _lib.load()
    flat, offs = [], [0]
    for user in user_sequence_dict:
        flat.extend(item2id[i] for i in user_sequence_dict[user][:-2])
        offs.append(len(flat))
    n = len(item2id)
    items = torch.tensor(flat if flat else [0], dtype=torch.int32, device=device)
    offsets = torch.tensor(offs, dtype=torch.int64, device=device)
    adj = torch.empty((n, n), dtype=torch.float32 if float32 > 0 else torch.float64, device=device)
    _lib.check(lib.p5_cooccurrence(items.data_ptr(), offsets.data_ptr(), len(offs) - 1, n, 0 if float32 > 0 else 1, adj.data_ptr(),
                                   _lib.current_stream_ptr()))
    return adj


def submatrix(adj: torch.Tensor, idx: List[int]) -> torch.Tensor:
    """ref utils/indexing.py:220-231: adj restricted to the items `idx` (zero diagonal)"""
    lib = _lib.load()
    m = len(idx)
    ix = torch.tensor(idx, dtype=torch.int32, device=adj.device)
    out = torch.empty((m, m), dtype=adj.dtype, device=adj.device)
    _lib.check(lib.p5_submatrix(adj.data_ptr(), adj.shape[0], 1 if adj.dtype == torch.float64 else 0, ix.data_ptr(), m, out.data_ptr(),
                                _lib.current_stream_ptr()))
    return out


def generate_collaborative_id(user_sequence_dict, token_size, cluster_num, last_token, float32, ref_indexing=None, device="cuda"):
    """same contract as ref utils/indexing.py:149 (the token bookkeeping helpers `add_token_to_indexing`,
    `add_last_token_to_indexing_*` are the reference's own, passed in as `ref_indexing` = its utils.indexing module)"""
    from sklearn.cluster import SpectralClustering
    if ref_indexing is None:
        import importlib
        ref_indexing = importlib.import_module("utils.indexing")
    all_items, train_items = set(), set()
    for user in user_sequence_dict:
        all_items.update(set(user_sequence_dict[user]))
        train_items.update(set(user_sequence_dict[user][:-2]))
    item2id, id2item = dict(), dict()
    for item in train_items:
        item2id[item] = len(item2id)
        id2item[len(id2item)] = item
    adj = cooccurrence_matrix(user_sequence_dict, item2id, float32, device)

    def cluster(mat: torch.Tensor):
        return SpectralClustering(n_clusters=cluster_num, assign_labels="cluster_qr", random_state=0,
                                  affinity="precomputed").fit(mat.cpu().numpy()).labels_.tolist()
    labels = cluster(adj)
    grouping = defaultdict(list)
    for i in range(len(labels)):
        grouping[labels[i]].append((id2item[i], i))
    item_map, index_now = ref_indexing.add_token_to_indexing(dict(), grouping, 0, token_size)
    queue = [grouping[g] for g in grouping]
    while queue:
        group_items = queue.pop(0)
        if len(group_items) <= token_size:
            item_list = [it[0] for it in group_items]
            if last_token == "sequential":
                item_map = ref_indexing.add_last_token_to_indexing_sequential(item_map, item_list, token_size)
            elif last_token == "random":
                item_map = ref_indexing.add_last_token_to_indexing_random(item_map, item_list, token_size)
        else:
            labels = cluster(submatrix(adj, [it[1] for it in group_items]))
            grouping = defaultdict(list)
            for i in range(len(labels)):
                grouping[labels[i]].append(group_items[i])
            item_map, index_now = ref_indexing.add_token_to_indexing(item_map, grouping, index_now, token_size)
            for g in grouping:
                queue.append(grouping[g])
    remaining = list(all_items - train_items)
    if remaining:
        if last_token == "sequential":
            item_map = ref_indexing.add_last_token_to_indexing_sequential(item_map, remaining, token_size)
        elif last_token == "random":
            item_map = ref_indexing.add_last_token_to_indexing_random(item_map, remaining, token_size)
    return item_map
