"""Host-side mirror of the reference's oracle interface on top of the C ABI (sls_gnn_*).

Reference interface mirrored here (relative to the reference repository):
  * ``network = hk.without_apply_rng(hk.transform(partial(get_network_definition("interaction", LCG),
    mlp_layers=mlp_layers)))`` and ``network.apply(params, graph)``
    (python/src/evaluate_with_given_params.py:97-101, :116; python/src/model.py:127-144)
    -> :class:`InteractionNetworkLCG` and its :meth:`apply`
  * ``LCG.get_model_probabilities(decoded_nodes, n)`` (python/src/sat_representations.py:756-780)
    -> :func:`get_model_probabilities`
  * the per-instance loop "build graph, apply, probabilities" of evaluate_with_given_params.py:108-121
    -> :meth:`InteractionNetworkLCG.probabilities` (one batched forward over many instances)

``params`` is the haiku parameter dict ``{module_name: {"w", "b"} | {"scale", "offset"}}`` (any
array-likes).  All dense layers run on the GPU (tcgen05 tensor cores); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .solver import Instance, _check

EMBED_DIM = 32  # model.py:14-15


def haiku_names(num_steps: int):
    """Module names in haiku construction order (model.py:13-16, :37-60, :144): edge embed ``linear``,
    node embed ``linear_1``; step t: edge ``layer_norm_{2t}``, ``linear_{2+4t}``, ``linear_{3+4t}``, node
    ``layer_norm_{2t+1}``, ``linear_{4+4t}``, ``linear_{5+4t}``; readout ``linear_{2+4*num_steps}``."""
    sfx = lambda base, i: base if i == 0 else f"{base}_{i}"
    names = {"edge_embed": "linear", "node_embed": "linear_1", "steps": [], "readout": sfx("linear", 2 + 4 * num_steps)}
    for t in range(num_steps):
        names["steps"].append({
            "edge": {"ln": sfx("layer_norm", 2 * t), "l1": sfx("linear", 2 + 4 * t), "l2": sfx("linear", 3 + 4 * t)},
            "node": {"ln": sfx("layer_norm", 2 * t + 1), "l1": sfx("linear", 4 + 4 * t), "l2": sfx("linear", 5 + 4 * t)},
        })
    return names


def _find(params, name):
    """haiku may prefix module names with a scope ("~/linear", "net/linear"): match on the last component."""
    if name in params:
        return params[name]
    hits = [k for k in params if k.split("/")[-1] == name]
    if len(hits) != 1:
        raise ValueError(f"parameter module {name!r} not found (have {sorted(params)[:6]}...)")
    return params[hits[0]]


def _natural_key(name: str):
    """'scope/linear_10' -> ('scope/linear', 10); a module without a numeric suffix is number 0 (haiku's first)."""
    head, _, tail = name.rpartition("_")
    return (head, int(tail)) if head and tail.isdigit() else (name, 0)


def resolve_modules(params, num_steps: int):
    """Map the network's modules to the keys of ``params``.  First by haiku's names (:func:`haiku_names`); if a file
    uses other names (another scope, another haiku version's numbering), by SHAPE AND CREATION ORDER: the modules
    holding ``w`` / ``b`` (Linear) and those holding ``scale`` / ``offset`` (LayerNorm) are taken in natural order of
    their numeric suffixes -- haiku numbers modules in construction order (model.py:13-16, :37-60, :144) and tree
    operations sort keys alphabetically, so dict order itself is not trusted -- and assigned to edge embed, node
    embed, per step edge l1 / l2, node l1 / l2, readout (LayerNorm: per step edge, node).  pack_params validates every
    shape afterwards, so a wrong guess is an error, never a silently permuted model."""
    names = haiku_names(num_steps)
    try:
        flat = [names["edge_embed"], names["node_embed"], names["readout"]]
        for st in names["steps"]:
            for kind in ("edge", "node"):
                flat += [st[kind]["l1"], st[kind]["l2"], st[kind]["ln"]]
        for nm in flat:
            _find(params, nm)
        return names
    except ValueError:
        pass
    lin = sorted((k for k in params if "w" in params[k]), key=_natural_key)
    lns = sorted((k for k in params if "scale" in params[k]), key=_natural_key)
    if len(lin) != 3 + 4 * num_steps or len(lns) != 2 * num_steps:
        raise ValueError(f"cannot map parameter modules: found {len(lin)} Linear and {len(lns)} LayerNorm modules, the "
                         f"interaction network with {num_steps} steps has {3 + 4 * num_steps} and {2 * num_steps}")
    out = {"edge_embed": lin[0], "node_embed": lin[1], "steps": [], "readout": lin[2 + 4 * num_steps]}
    for t in range(num_steps):
        out["steps"].append({
            "edge": {"ln": lns[2 * t], "l1": lin[2 + 4 * t], "l2": lin[3 + 4 * t]},
            "node": {"ln": lns[2 * t + 1], "l1": lin[4 + 4 * t], "l2": lin[5 + 4 * t]},
        })
    return out


def pack_params(params, num_steps: int):
    """Flatten a haiku parameter dict into the float32 blob sls_gnn_model_create expects; returns
    (blob, embed_dim, h1, h2).  Shapes are validated against the interaction-network definition."""
    names = resolve_modules(params, num_steps)
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    parts = []
    ee, ne = _find(params, names["edge_embed"]), _find(params, names["node_embed"])
    D = f32(ee["w"]).shape[1]
    for p in (ee, ne):
        w, b = f32(p["w"]), f32(p["b"])
        if w.shape != (2, D) or b.shape != (D,):
            raise ValueError(f"embedding must be Linear(2 -> {D}), got w{w.shape} b{b.shape}")
        parts += [w.ravel(), b]
    first = _find(params, names["steps"][0]["edge"]["l1"])
    h1 = f32(first["w"]).shape[1]
    h2 = f32(_find(params, names["steps"][0]["edge"]["l2"])["w"]).shape[1]
    de = dn = D
    for t in range(num_steps):
        for kind, k1 in (("edge", de + 2 * dn), ("node", dn + 2 * h2)):
            s = names["steps"][t][kind]
            l1, l2, ln = _find(params, s["l1"]), _find(params, s["l2"]), _find(params, s["ln"])
            w1, b1, w2, b2 = f32(l1["w"]), f32(l1["b"]), f32(l2["w"]), f32(l2["b"])
            sc, of = f32(ln["scale"]), f32(ln["offset"])
            if w1.shape != (k1, h1) or b1.shape != (h1,) or w2.shape != (h1, h2) or b2.shape != (h2,):
                raise ValueError(f"step {t} {kind} MLP has unexpected shapes w1{w1.shape} w2{w2.shape} (expected ({k1},{h1}), ({h1},{h2}))")
            if sc.shape != (h2,) or of.shape != (h2,):
                raise ValueError(f"step {t} {kind} LayerNorm has unexpected shapes {sc.shape} {of.shape}")
            parts += [w1.ravel(), b1, w2.ravel(), b2, sc, of]
        de = dn = h2
    ro = _find(params, names["readout"])
    w, b = f32(ro["w"]), f32(ro["b"])
    if w.shape != (h2, 1) or b.shape != (1,):
        raise ValueError(f"readout must be Linear({h2} -> 1) for LCG (model.py:144), got w{w.shape}")
    parts += [w.ravel(), b]
    return np.ascontiguousarray(np.concatenate(parts)), D, h1, h2


def get_model_probabilities(decoded_nodes, n_variables: int) -> np.ndarray:
    """LCG.get_model_probabilities (sat_representations.py:775-780) on an already computed decoded_nodes
    array: pair-softmax column 1 = sigmoid(decoded[2j+1] - decoded[2j]).  O(n) host arithmetic on the
    network's output, kept for call-site compatibility; the batched path does this on the GPU."""
    d = np.asarray(decoded_nodes, dtype=np.float32).reshape(-1)
    if d.shape[0] % 2 == 1:
        d = np.concatenate([d, np.zeros(1, dtype=np.float32)])
    pairs = d.reshape(-1, 2)
    z = pairs - pairs.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True))[:n_variables, 1]


def get_constraint_graph(n_variables: int, n_clauses: int, senders, receivers) -> dict:
    """LCG.get_constraint_graph (sat_representations.py:668-715) with the reference's own arguments: the LCG edge list
    (literal -> clause edges, then the n pair edges).  Two clauses are neighbours when they share a variable; the result is
    the jraph-style dict of this module (``senders`` / ``receivers`` in LCG node numbering, clause c = node 2n + c, zero
    ``edges`` / ``nodes``).  The reference takes its edge order from scipy's sparse product; here every clause lists its
    neighbours in increasing index -- the same edge SET, which is all the LLL loss (segment sums) depends on."""
    senders = np.asarray(senders, dtype=np.int64).reshape(-1)
    receivers = np.asarray(receivers, dtype=np.int64).reshape(-1)
    n_occ = len(senders) - int(n_variables)
    var_of = np.ascontiguousarray(senders[:n_occ] // 2, dtype=np.int32)
    clause_of = np.ascontiguousarray(receivers[:n_occ] - 2 * int(n_variables), dtype=np.int32)
    lib = _lib.load()
    n_edge = C.c_uint64(0)
    _check(lib.sls_constraint_graph(int(n_variables), int(n_clauses), var_of.ctypes.data, clause_of.ctypes.data, n_occ, None, None, 0,
                                    C.byref(n_edge)))
    out_s = np.zeros(max(n_edge.value, 1), dtype=np.int64)
    out_r = np.zeros(max(n_edge.value, 1), dtype=np.int64)
    _check(lib.sls_constraint_graph(int(n_variables), int(n_clauses), var_of.ctypes.data, clause_of.ctypes.data, n_occ,
                                    out_s.ctypes.data, out_r.ctypes.data, out_s.size, C.byref(n_edge)))
    e = int(n_edge.value)
    return {"n_node": int(n_clauses), "n_edge": e, "senders": out_s[:e], "receivers": out_r[:e],
            "edges": np.zeros(e), "nodes": np.zeros(int(n_clauses))}


class InteractionNetworkLCG:
    """The shipped oracle network: embedding + ``num_message_passing_steps`` interaction-network steps
    + Linear(1) readout on the literal-clause graph."""

    def __init__(self, params, num_message_passing_steps: int = 5):
        blob, D, h1, h2 = pack_params(params, num_message_passing_steps)
        self.embed_dim, self.h1, self.h2, self.num_steps = D, h1, h2, num_message_passing_steps
        self._h = C.c_void_p()
        _check(_lib.load().sls_gnn_model_create(blob.ctypes.data, blob.size, num_message_passing_steps, D, h1, h2, C.byref(self._h)))
        self.last_kernel_ms = 0.0

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().sls_gnn_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, graph) -> np.ndarray:
        """``network.apply(params, graph)``: graph has nodes[N,2], edges[E,2], senders[E], receivers[E]
        (a jraph.GraphsTuple, any namedtuple with these fields, or a dict).  Returns decoded_nodes [N,1]."""
        get = (lambda k: graph[k]) if isinstance(graph, dict) else (lambda k: getattr(graph, k))
        nodes = np.ascontiguousarray(np.asarray(get("nodes"), dtype=np.float32))
        edges = np.ascontiguousarray(np.asarray(get("edges"), dtype=np.float32))
        send = np.ascontiguousarray(np.asarray(get("senders"), dtype=np.int32))
        recv = np.ascontiguousarray(np.asarray(get("receivers"), dtype=np.int32))
        N, E = nodes.shape[0], edges.shape[0]
        if nodes.shape != (N, 2) or edges.shape != (E, 2) or send.shape != (E,) or recv.shape != (E,):
            raise ValueError("graph must have nodes[N,2], edges[E,2], senders[E], receivers[E]")
        out = np.zeros(max(N, 1), dtype=np.float32)
        ms = C.c_double(0)
        _check(_lib.load().sls_gnn_forward_graph(self._h, N, E, nodes.ctypes.data, edges.ctypes.data, send.ctypes.data,
                                                 recv.ctypes.data, out.ctypes.data, C.byref(ms)))
        self.last_kernel_ms = ms.value
        return out[:N].reshape(N, 1)

    def probabilities(self, instances, return_decoded: bool = False):
        """One batched forward over the LCG graphs of ``instances`` (list of :class:`Instance`); returns a
        list with one probability vector p[n_i] = P(x_j = True) per instance."""
        instances = list(instances)
        handles = (C.c_void_p * max(len(instances), 1))(*[i._h.value for i in instances])
        ns = [i.n for i in instances]
        total_n = int(sum(ns))
        prob = np.zeros(max(total_n, 1), dtype=np.float32)
        dec = None
        if return_decoded:
            total_nodes = 0
            for i in instances:
                row, _ = i.csr()
                total_nodes += 2 * i.n + int(np.count_nonzero(np.diff(row.astype(np.int64)) > 0))
            dec = np.zeros(max(total_nodes, 1), dtype=np.float32)
        ms = C.c_double(0)
        _check(_lib.load().sls_gnn_forward_lcg(self._h, handles, len(instances), prob.ctypes.data,
                                               dec.ctypes.data if dec is not None else None, C.byref(ms)))
        self.last_kernel_ms = ms.value
        out = np.split(prob[:total_n], np.cumsum(ns)[:-1]) if instances else []
        return (out, dec) if return_decoded else out

    def launch_count(self) -> int:
        return _lib.load().sls_gnn_launch_count(self._h)

    def set_profile(self, on: bool) -> None:
        """Measurement only: bracket every kernel class of the following forwards with CUDA events."""
        _check(_lib.load().sls_gnn_set_profile(self._h, int(bool(on))))

    def profile(self) -> dict:
        """Device milliseconds of the last forward by kernel class (needs :meth:`set_profile`)."""
        ms = (C.c_double * 4)()
        _check(_lib.load().sls_gnn_get_profile(self._h, ms))
        return {"edge_update_ms": ms[0], "node_update_ms": ms[1], "segment_sum_ms": ms[2], "other_ms": ms[3]}


def network_definition_interaction_lcg(graph, mlp_layers, num_message_passing_steps: int = 5):
    """Marker with the reference's signature (model.py:127-144).  It is never traced: :func:`transform` recognises
    it and runs the CUDA forward instead."""
    raise RuntimeError("use transform(network_definition).apply(params, graph)")


def get_network_definition(network_type, graph_representation="LCG"):
    """Mirror of model.get_network_definition (model.py:183-206) for the shipped combination only."""
    rep = graph_representation if isinstance(graph_representation, str) else getattr(graph_representation, "__name__", "")
    if network_type != "interaction" or rep != "LCG":
        raise ValueError("Invalid network_type or graph_representation (only 'interaction' + LCG is built, see DESIGN.md)")
    return network_definition_interaction_lcg


class _Transformed:
    """What ``hk.without_apply_rng(hk.transform(network_definition))`` gives the reference's call sites
    (evaluate_with_given_params.py:97-101): an object whose ``apply(params, graph)`` returns decoded_nodes [N,1].
    The packed device copy of the last ``params`` is cached together with a reference to the object and a fingerprint of
    its arrays: the cache hits only for the very same object (``is``) with unchanged contents, so neither a recycled
    address nor an in-place update can make the forward run with stale weights."""

    def __init__(self, fn):
        import functools

        self.steps = 5
        f = fn
        while isinstance(f, functools.partial):  # partial(network_definition, mlp_layers=...) as in the reference
            self.steps = f.keywords.get("num_message_passing_steps", self.steps)
            f = f.func
        if f is not network_definition_interaction_lcg:
            raise ValueError("only get_network_definition('interaction', LCG) is supported")
        self._params = None   # the cached object itself (keeps it alive: its address cannot be reused)
        self._print = None
        self._net = None

    @staticmethod
    def _fingerprint(params):
        """Cheap content check of a haiku parameter dict: shapes plus a strided sample and the sum of every array
        (1.5 M floats: ~1 ms, against ~10 ms for re-packing and re-uploading the model)."""
        out = []
        for mod in sorted(params):
            for name in sorted(params[mod]):
                a = np.asarray(params[mod][name])
                flat = a.reshape(-1)
                out.append((mod, name, a.shape, float(flat.sum(dtype=np.float64)), flat[:: max(1, flat.size // 16)].tobytes()))
        return out

    def apply(self, params, graph):
        fp = self._fingerprint(params)
        if self._net is None or self._params is not params or self._print != fp:
            if self._net is not None:
                self._net.close()
            self._net = InteractionNetworkLCG(params, self.steps)
            self._params, self._print = params, fp
        return self._net.apply(graph)


def transform(network_definition):
    """Stand-in for ``hk.transform`` at the reference's call site."""
    return _Transformed(network_definition)


def without_apply_rng(transformed):
    """Stand-in for ``hk.without_apply_rng``."""
    return transformed
