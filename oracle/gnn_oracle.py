"""gnn_oracle.py -- NumPy restatement of the reference's GNN oracle forward pass.

TEST INFRASTRUCTURE ONLY (checker for the CUDA oracle-forward kernels, and the CPU baseline of
that leg).  Nothing in the product package imports this file.

What it restates (file:line relative to the reference repository):
  python/src/sat_representations.py:607-666   LCG.get_graph          -> lcg_graph()
  python/src/sat_instances.py:50-73           get_problem_from_cnf   -> lcg_graph() (one-hot features)
  python/src/model.py:11-18                   get_embedding          -> forward() step 1
  python/src/model.py:21-61                   apply_interaction      -> forward() step 2
  python/src/model.py:127-144                 network_definition_interaction_lcg -> forward()
  python/src/sat_representations.py:756-780   LCG.get_model_probabilities -> model_probabilities()

Third-party behaviour restated from the published sources (not under /root/reference; pinned in
requirements.txt:43-49): jraph 0.0.6.dev0 ``InteractionNetwork`` (edge update sees
[edges, nodes[senders], nodes[receivers]]; node update sees [nodes, segment_sum(edges', senders),
segment_sum(edges', receivers)] with include_sent_messages_in_node_update=True), ``concatenated_args``
(concatenate along the last axis in argument order); dm-haiku ``Linear`` (x @ w + b), ``LayerNorm``
(axis=-1, eps=1e-5, biased variance, learned scale and offset), module naming by construction order
(``linear``, ``linear_1``, ... / ``layer_norm``, ``layer_norm_1``, ...).

PARITY STATUS: UNPINNED.  jax / haiku / jraph are not installable here, the pre-trained parameter
files are absent (/root/reference/.MISSING_LARGE_BLOBS) and the reference's tests store no
activations or logits, so this restatement cannot be checked against the real
reference in this environment.  It follows the cited lines literally.
"""
from __future__ import annotations

import numpy as np

EMBED = 32
LN_EPS = 1e-5
NUM_STEPS = 5


def lcg_graph(n_variables: int, clauses):
    """LCG.get_graph + the one-hot encoding of get_problem_from_cnf.

    Returns dict(nodes[N,2], edges[E,2], senders[E], receivers[E], n_node, n_edge).
    Node 2j is the negative literal of variable j, node 2j+1 the positive one, nodes >= 2n are the
    clauses in file order (sat_representations.py:636-657).  Raw node labels are 1 / -1 / 0 and the
    one-hot lookup ``np.eye(2)[label]`` maps BOTH 1 and -1 to [0, 1] (sat_instances.py:69).
    """
    clauses = [list(c) for c in clauses if len(c) > 0]  # sat_instances.py:50
    m = len(clauses)
    n = n_variables
    n_node = 2 * n + m
    labels = np.zeros(n_node, dtype=np.int64)
    labels[0:2 * n:2] = 1
    labels[1:2 * n:2] = -1
    senders, receivers, edge_labels = [], [], []
    for j, clause in enumerate(clauses):
        support = [abs(l) - 1 for l in clause]
        assert len(support) == len(set(support)), "Multiple occurrences of single variable in constraint"
        for l in clause:
            senders.append(2 * (abs(l) - 1) + (1 if l > 0 else 0))
            receivers.append(2 * n + j)
            edge_labels.append(0)
    for j in range(n):  # sat_representations.py:661-664
        senders.append(2 * j + 1)
        receivers.append(2 * j)
        edge_labels.append(1)
    eye = np.eye(2)
    return {
        "nodes": eye[labels],
        "edges": eye[np.asarray(edge_labels, dtype=np.int64)],
        "senders": np.asarray(senders, dtype=np.int64),
        "receivers": np.asarray(receivers, dtype=np.int64),
        "n_node": n_node,
        "n_edge": len(senders),
    }


def param_names(num_steps: int = NUM_STEPS):
    """haiku module names in construction order: (edge embed, node embed, per step (edge LN, edge
    linears, node LN, node linears), readout)."""
    names = {"edge_embed": "linear", "node_embed": "linear_1", "steps": [], "readout": None}
    lin = 2
    for t in range(num_steps):
        step = {}
        for kind, ln in (("edge", 2 * t), ("node", 2 * t + 1)):
            step[kind] = {
                "ln": "layer_norm" if ln == 0 else f"layer_norm_{ln}",
                "l1": f"linear_{lin}",
                "l2": f"linear_{lin + 1}",
            }
            lin += 2
        names["steps"].append(step)
    names["readout"] = f"linear_{lin}"
    return names


def init_params(seed: int = 42, mlp_layers=(200, 200), num_steps: int = NUM_STEPS, bias_scale: float = 0.0):
    """Random parameters with haiku's default initialisers: Linear weights ~ truncated normal
    (+-2 sigma) with sigma = 1/sqrt(fan_in), zero biases, LayerNorm scale 1 / offset 0.
    ``bias_scale`` > 0 additionally randomises biases / LN parameters so tests exercise them."""
    rng = np.random.default_rng(seed)
    h1, h2 = mlp_layers

    def lin(fan_in, fan_out):
        w = rng.standard_normal((fan_in, fan_out))
        while (np.abs(w) > 2).any():
            bad = np.abs(w) > 2
            w[bad] = rng.standard_normal(int(bad.sum()))
        p = {"w": (w / np.sqrt(fan_in)).astype(np.float32), "b": np.zeros(fan_out, dtype=np.float32)}
        if bias_scale:
            p["b"] = (bias_scale * rng.standard_normal(fan_out)).astype(np.float32)
        return p

    def ln(dim):
        p = {"scale": np.ones(dim, dtype=np.float32), "offset": np.zeros(dim, dtype=np.float32)}
        if bias_scale:
            p["scale"] = (1 + bias_scale * rng.standard_normal(dim)).astype(np.float32)
            p["offset"] = (bias_scale * rng.standard_normal(dim)).astype(np.float32)
        return p

    names = param_names(num_steps)
    params = {names["edge_embed"]: lin(2, EMBED), names["node_embed"]: lin(2, EMBED)}
    e_dim, h_dim = EMBED, EMBED
    for t in range(num_steps):
        s = names["steps"][t]
        params[s["edge"]["ln"]] = ln(h2)
        params[s["edge"]["l1"]] = lin(e_dim + 2 * h_dim, h1)
        params[s["edge"]["l2"]] = lin(h1, h2)
        params[s["node"]["ln"]] = ln(h2)
        params[s["node"]["l1"]] = lin(h_dim + 2 * h2, h1)
        params[s["node"]["l2"]] = lin(h1, h2)
        e_dim, h_dim = h2, h2
    params[names["readout"]] = lin(h_dim, 1)
    return params


def _segment_sum(data, ids, num):
    out = np.zeros((num, data.shape[1]), dtype=data.dtype)
    np.add.at(out, ids, data)
    return out


def forward(params, graph, dtype=np.float32, num_steps: int = NUM_STEPS):
    """network.apply(params, graph) of evaluate_with_given_params.py:116 -> decoded_nodes [N, 1]."""
    names = param_names(num_steps)
    P = {k: {kk: np.asarray(vv, dtype=dtype) for kk, vv in v.items()} for k, v in params.items()}
    relu = lambda x: np.maximum(x, 0)

    def linear(name, x):
        return x @ P[name]["w"] + P[name]["b"]

    def layer_norm(name, x):
        mean = x.mean(axis=-1, keepdims=True)
        var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
        return (x - mean) / np.sqrt(var + dtype(LN_EPS)) * P[name]["scale"] + P[name]["offset"]

    def update(s, feats):  # model.py:44-52: layer_norm(mlp(features)), ReLU after both linears
        return layer_norm(s["ln"], relu(linear(s["l2"], relu(linear(s["l1"], feats)))))

    senders, receivers = graph["senders"], graph["receivers"]
    N = graph["n_node"]
    e = linear(names["edge_embed"], graph["edges"].astype(dtype))  # model.py:13-16
    h = linear(names["node_embed"], graph["nodes"].astype(dtype))
    for t in range(num_steps):  # model.py:54-60
        s = names["steps"][t]
        e = update(s["edge"], np.concatenate([e, h[senders], h[receivers]], axis=1))
        sent = _segment_sum(e, senders, N)
        recv = _segment_sum(e, receivers, N)
        h = update(s["node"], np.concatenate([h, sent, recv], axis=1))
    return linear(names["readout"], h)  # model.py:144


def model_probabilities(decoded_nodes, n_variables: int):
    """LCG.get_model_probabilities (sat_representations.py:775-780)."""
    d = np.asarray(decoded_nodes)
    if d.shape[0] % 2 == 1:
        d = np.vstack([d, [[0]]]).astype(d.dtype)
    pairs = d.reshape(-1, 2)
    z = pairs - pairs.max(axis=1, keepdims=True)
    sm = np.exp(z) / np.exp(z).sum(axis=1, keepdims=True)
    return sm[:n_variables, 1]
