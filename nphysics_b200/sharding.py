"""Island labelling and island/world sharding across ranks.

Bodies in different connected components of the (dynamic body)-(contact | joint) graph never
exchange data inside a step (rows touch only their 1-2 bodies, src/solver/sor_prox.rs:188-209), so
whole islands are assigned to ranks and every rank steps its own context with no data-path
collective; only per-rank stats are gathered.  The labelling is the union-find of
src/utils/union_find.rs:30-59 as ActivationManager uses it (src/detection/activation_manager.rs:
122-159): non-dynamic bodies are skipped so the ground does not glue islands together (:141-145).
"""
import numpy as np

from . import abi


def island_labels(status, pairs_a, pairs_b):
    """Connected-component label per body (dynamic bodies only; others get -1).

    pairs_a/pairs_b: body indices of every manifold and joint."""
    n = len(status)
    dyn = status == abi.BODY_DYNAMIC
    a = np.asarray(pairs_a, dtype=np.int64)
    b = np.asarray(pairs_b, dtype=np.int64)
    keep = dyn[a] & dyn[b]
    a, b = a[keep], b[keep]
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    g = coo_matrix((np.ones(len(a), dtype=np.int8), (a, b)), shape=(n, n))
    _, lab = connected_components(g, directed=False)
    lab = lab.astype(np.int64)
    lab[~dyn] = -1
    # renumber densely in order of first appearance
    used = lab[dyn]
    uniq, first = np.unique(used, return_index=True)
    order = uniq[np.argsort(first)]
    remap = -np.ones(lab.max() + 2, dtype=np.int64)
    remap[order] = np.arange(len(order))
    out = np.where(lab >= 0, remap[np.maximum(lab, 0)], -1)
    return out


def assign_islands(weights, n_ranks):
    """Greedy longest-processing-time bin packing of islands by weight (row count).
    Returns rank per island; deterministic."""
    weights = np.asarray(weights, dtype=np.int64)
    order = np.argsort(-weights, kind="stable")
    load = np.zeros(n_ranks, dtype=np.int64)
    rank_of = np.zeros(len(weights), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        rank_of[i] = r
        load[r] += weights[i]
    return rank_of, load


class Shard:
    """The sub-world one rank owns: its bodies (global indices), joints and a manifold filter."""

    def __init__(self, body_ids, bodies, joints, joint_ids, global_to_local):
        self.body_ids = body_ids
        self.bodies = bodies
        self.joints = joints
        self.joint_ids = joint_ids
        self.global_to_local = global_to_local

    def localize_colliders(self, colliders):
        """The colliders attached to this shard's bodies, re-indexed (for the device manifold producer)."""
        lb = self.global_to_local[colliders["body"]]
        c = colliders[lb >= 0].copy()
        c["body"] = lb[lb >= 0]
        return c

    def localize_manifolds(self, manifolds, contacts):
        """Keep the manifolds whose dynamic bodies live in this shard; re-index bodies and contacts."""
        g2l = self.global_to_local
        l1 = g2l[manifolds["body1"]]
        l2 = g2l[manifolds["body2"]]
        keep = (l1 >= 0) & (l2 >= 0)
        m = manifolds[keep].copy()
        m["body1"] = l1[keep]
        m["body2"] = l2[keep]
        nc = m["num_contacts"].astype(np.int64)
        starts = m["first_contact"].astype(np.int64)
        new_first = np.concatenate([[0], np.cumsum(nc)[:-1]]) if len(m) else np.zeros(0, dtype=np.int64)
        idx = np.repeat(starts - new_first, nc) + np.arange(int(nc.sum()))
        c = contacts[idx].copy() if len(idx) else contacts[:0].copy()
        m["first_contact"] = new_first
        return m, c, idx


def dense_labels(labels):
    """Representative labels (any non-negative ids, e.g. the device's "smallest body index of the island")
    renumbered 0, 1, ... in order of first appearance; -1 stays -1."""
    lab = np.asarray(labels, dtype=np.int64)
    out = -np.ones(len(lab), dtype=np.int64)
    ok = lab >= 0
    if ok.any():
        uniq, first, inv = np.unique(lab[ok], return_index=True, return_inverse=True)
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
        out[ok] = rank[inv]
    return out


def make_shards(bodies, joints, pairs_a=None, pairs_b=None, pair_weights=None, n_ranks=1, labels=None, body_rows=None):
    """Partition a world into n_ranks shards of whole islands.

    Two sources for the islands and their weights:
      * labels / body_rows: the DEVICE labelling (nb2_label_islands: csrc/activation.cu) -- island id per
        body (-1 for non-dynamic bodies) and the velocity rows booked on every body;
      * pairs_a / pairs_b / pair_weights: body pairs of every potential manifold + joint with their rows,
        labelled on the host (island_labels above; used by the CPU tests and as the device's cross-check).
    Non-dynamic bodies (ground, kinematic) are replicated into every shard that references them."""
    status = bodies["status"]
    if labels is not None:
        lab = dense_labels(labels)
        n_islands = int(lab.max()) + 1 if (lab >= 0).any() else 0
        w = np.zeros(max(n_islands, 1), dtype=np.int64)
        ok = lab >= 0
        if body_rows is not None:
            np.add.at(w, lab[ok], np.asarray(body_rows, dtype=np.int64)[ok])
    else:
        lab = island_labels(status, pairs_a, pairs_b)
        n_islands = int(lab.max()) + 1 if (lab >= 0).any() else 0
        pa = np.asarray(pairs_a, dtype=np.int64)
        pb = np.asarray(pairs_b, dtype=np.int64)
        pl = np.where(lab[pa] >= 0, lab[pa], lab[pb])
        w = np.zeros(max(n_islands, 1), dtype=np.int64)
        ok = pl >= 0
        np.add.at(w, pl[ok], np.asarray(pair_weights, dtype=np.int64)[ok])
    # islands without rows still need an owner: weight 1
    w = np.maximum(w, 1)[:max(n_islands, 0)]
    rank_of_island, load = assign_islands(w, n_ranks)
    shards = []
    non_dyn = np.nonzero(status != abi.BODY_DYNAMIC)[0]
    for r in range(n_ranks):
        mine = np.nonzero((lab >= 0) & (rank_of_island[np.maximum(lab, 0)] == r))[0]
        ids = np.sort(np.concatenate([non_dyn, mine]))
        g2l = -np.ones(len(bodies), dtype=np.int64)
        g2l[ids] = np.arange(len(ids))
        jkeep = np.zeros(len(joints), dtype=bool)
        if len(joints):
            j1, j2 = g2l[joints["body1"]], g2l[joints["body2"]]
            d1 = status[joints["body1"]] == abi.BODY_DYNAMIC
            d2 = status[joints["body2"]] == abi.BODY_DYNAMIC
            owner_lab = np.where(d1, lab[joints["body1"]], lab[joints["body2"]])
            jkeep = (owner_lab >= 0) & (rank_of_island[np.maximum(owner_lab, 0)] == r) & (j1 >= 0) & (j2 >= 0)
        js = joints[jkeep].copy()
        if len(js):
            js["body1"] = g2l[js["body1"]]
            js["body2"] = g2l[js["body2"]]
        shards.append(Shard(ids, bodies[ids].copy(), js, np.nonzero(jkeep)[0], g2l))
    return shards, lab, load


def gather_stats(stats, dist=None):
    """all_gather of the per-rank nb2_stats records (the only collective of the multi-GPU path).
    `dist` is torch.distributed (nccl on GPUs, gloo in the CPU tests) or None for a single rank."""
    rec = np.ascontiguousarray(stats, dtype=abi.stats_dtype).reshape(1)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return rec.copy()
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(rec.view(np.uint8).copy()).to(dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy().view(abi.stats_dtype) for o in out])


def combine_stats(all_stats):
    """Whole-job view of gathered per-rank stats: counters add, residual/penetration take the max."""
    out = np.zeros((), dtype=abi.stats_dtype)
    for name in abi.stats_dtype.names:
        col = all_stats[name]
        if name in ("residual_max", "max_penetration", "n_phases_velocity", "n_phases_position") or \
                name.startswith("t_"):
            out[name] = col.max()
        elif name == "residual_rms":
            out[name] = np.sqrt(np.mean(col.astype(np.float64) ** 2))
        elif name.startswith("pad"):
            continue
        else:
            out[name] = col.sum()
    return out
