"""gnn_torch.py -- a SECOND, independent CPU model of the reference's GNN oracle forward (torch, fp32).

TEST INFRASTRUCTURE ONLY.  Two jobs:
  * cross-check of oracle/gnn_oracle.py (the NumPy restatement the CUDA kernels are compared with): the two files share
    no code, no name formula and no array layout -- this one builds the graph from flat literal arrays with vectorised
    index arithmetic, derives haiku's module names by *replaying the construction order* of the reference's model code
    through small stand-in module classes, aggregates with ``index_add_`` (scatter-add, like jraph's segment_sum) and
    normalises with ``torch.nn.functional.layer_norm``; tests/test_oracle_golden.py requires both to agree to fp32 noise;
  * the CPU baseline of bench.py's oracle-forward leg ("torch-CPU fp32 restatement").

What it follows (file:line relative to the reference repository):
  python/src/model.py:11-18      get_embedding: GraphMapFeatures(embed_edge_fn = Linear(32), embed_node_fn = Linear(32))
  python/src/model.py:21-61      apply_interaction: per step InteractionNetwork(update_edge_fn = update_fn,
                                 update_node_fn = update_fn, include_sent_messages_in_node_update = True),
                                 update_fn = concatenated_args(LayerNorm(Sequential([Linear, relu] * len(mlp_layers))))
  python/src/model.py:127-144    network_definition_interaction_lcg: embedding, interaction, Linear(1) on the nodes
  python/src/sat_representations.py:607-666 + python/src/sat_instances.py:50-73   LCG graph and its one-hot features
  python/src/sat_representations.py:775-780   get_model_probabilities

PARITY STATUS: UNPINNED (same reason as gnn_oracle.py: jax / haiku / jraph and the trained parameter files are absent).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------- haiku stand-ins: names by construction order
class _Scope:
    """Hands out haiku-style module names: the first module of a class gets the snake-case class name, later ones
    ``name_1``, ``name_2``, ... in the order they are CONSTRUCTED (not called)."""

    def __init__(self, params):
        self.params = params
        self.count = {}

    def name(self, base: str) -> str:
        i = self.count.get(base, 0)
        self.count[base] = i + 1
        return base if i == 0 else f"{base}_{i}"

    def get(self, module: str):
        if module in self.params:
            return self.params[module]
        hits = [k for k in self.params if k.rsplit("/", 1)[-1] == module]
        if len(hits) != 1:
            raise KeyError(module)
        return self.params[hits[0]]


class Linear:
    def __init__(self, scope: _Scope, output_size: int):
        self.scope, self.output_size = scope, output_size
        self.module = scope.name("linear")

    def __call__(self, x):
        p = self.scope.get(self.module)
        w = torch.as_tensor(np.asarray(p["w"], dtype=np.float32))
        b = torch.as_tensor(np.asarray(p["b"], dtype=np.float32))
        assert w.shape == (x.shape[-1], self.output_size), (self.module, tuple(w.shape), x.shape[-1], self.output_size)
        return F.linear(x, w.t().contiguous(), b)


class LayerNorm:
    def __init__(self, scope: _Scope):
        self.scope = scope
        self.module = scope.name("layer_norm")

    def __call__(self, x):
        p = self.scope.get(self.module)
        scale = torch.as_tensor(np.asarray(p["scale"], dtype=np.float32))
        offset = torch.as_tensor(np.asarray(p["offset"], dtype=np.float32))
        return F.layer_norm(x, (x.shape[-1],), scale, offset, eps=1e-5)  # haiku: eps = 1e-5, biased variance


# ---------------------------------------------------------------- the graph, from flat literal arrays
def lcg_arrays(n_variables: int, row_ptr, signed_lits):
    """senders, receivers (int64 tensors) and the one-hot node / edge features of the LCG graph of a CNF given as a
    clause CSR (row_ptr[m+1], signed DIMACS literals).  Empty clauses are dropped (sat_instances.py:50)."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    lits = np.asarray(signed_lits, dtype=np.int64)
    lens = np.diff(row_ptr)
    keep = lens > 0
    m = int(keep.sum())
    n = int(n_variables)
    clause_of_lit = np.repeat(np.cumsum(keep) - 1, lens)  # empty clauses own no literal, so no index is wasted on them
    var = np.abs(lits) - 1
    lit_node = 2 * var + (lits > 0)                      # positive literal = odd node, negative = even node (:653-657)
    pair = np.arange(n, dtype=np.int64)
    senders = np.concatenate([lit_node, 2 * pair + 1])   # literal -> clause edges in file order, then x -> not-x (:661-664)
    receivers = np.concatenate([2 * n + clause_of_lit, 2 * pair])
    # np.eye(2)[label] with labels 1 / -1 (literals) and 0 (clauses): both literal polarities are [0, 1]
    nodes = np.zeros((2 * n + m, 2), dtype=np.float32)
    nodes[: 2 * n, 1] = 1.0
    nodes[2 * n:, 0] = 1.0
    edges = np.zeros((len(senders), 2), dtype=np.float32)
    edges[: len(lits), 0] = 1.0
    edges[len(lits):, 1] = 1.0
    return (torch.from_numpy(nodes), torch.from_numpy(edges), torch.from_numpy(senders), torch.from_numpy(receivers))


def graph_from_clauses(n_variables: int, clauses):
    lens = [len(c) for c in clauses]
    row_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = np.asarray([l for c in clauses for l in c], dtype=np.int64)
    return lcg_arrays(n_variables, row_ptr, flat)


# ---------------------------------------------------------------- the network
def forward(params, nodes, edges, senders, receivers, mlp_layers=(200, 200), num_message_passing_steps: int = 5):
    """network.apply(params, graph) -> decoded_nodes [N, 1] (float32 torch tensor)."""
    scope = _Scope(params)
    n_node = nodes.shape[0]
    with torch.no_grad():
        # get_embedding: the edge Linear is constructed before the node Linear (keyword order at model.py:13-16)
        embed_edge, embed_node = Linear(scope, 32), Linear(scope, 32)
        e, h = embed_edge(edges), embed_node(nodes)

        def update_fn(*args):
            feats = torch.cat(args, dim=-1)                 # jraph.concatenated_args
            layer_norm = LayerNorm(scope)                   # constructed before the MLP (model.py:48-51)
            net = [Linear(scope, d) for d in mlp_layers]    # mlp(): [Linear, relu] per layer
            x = feats
            for lin in net:
                x = torch.relu(lin(x))
            return layer_norm(x)

        for _ in range(num_message_passing_steps):
            # jraph.InteractionNetwork: edges first, from (edges, sender nodes, receiver nodes); then nodes, from
            # (nodes, sum of the NEW edges each node sends, sum of the new edges it receives)
            e = update_fn(e, h[senders], h[receivers])
            sent = torch.zeros((n_node, e.shape[1]), dtype=e.dtype).index_add_(0, senders, e)
            recv = torch.zeros((n_node, e.shape[1]), dtype=e.dtype).index_add_(0, receivers, e)
            h = update_fn(h, sent, recv)
        return Linear(scope, 1)(h)


def model_probabilities(decoded, n_variables: int):
    """Pair softmax, column 1 of the first n rows (sat_representations.py:775-780)."""
    d = decoded.reshape(-1)
    if d.shape[0] % 2:
        d = torch.cat([d, torch.zeros(1, dtype=d.dtype)])
    return torch.softmax(d.reshape(-1, 2), dim=1)[:n_variables, 1].numpy()


def probabilities(params, n_variables: int, row_ptr, signed_lits, **kw):
    nodes, edges, s, r = lcg_arrays(n_variables, row_ptr, signed_lits)
    return model_probabilities(forward(params, nodes, edges, s, r, **kw), n_variables)
