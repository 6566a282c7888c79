"""Reading the reference's trained-oracle files without jax / haiku / the reference package.

The reference writes ``jnp.save(model_path, np.asarray([params, [inv_temp, alpha, beta, gamma, mlp_layers,
graph_representation, network_type, return_candidates]], dtype=object))`` (python/src/train.py:221-239) and
reads it back with ``np.load(model_path, allow_pickle=True)`` (python/src/evaluate_with_given_params.py:67).
The payload is a pickle that references ``jax`` array reconstructors, possibly haiku's mapping type, and the
class object ``python.src.sat_representations.LCG`` -- none of which exist on a box that only has this engine.
:func:`load_model` unpickles it with a restricted ``find_class`` that maps those references to NumPy arrays,
plain dicts and small stand-in classes, and returns ``(params, details)`` ready for
:class:`sls_sat_solving_with_deep_learning_b200.gnn.InteractionNetworkLCG`.

STATUS: the nine shipped ``Data/models/**/*.npy`` files are absent from the reference checkout this was built
against (``.MISSING_LARGE_BLOBS``), so this loader is exercised only on files produced by the test-suite with
stand-in ``jax`` modules that pickle the way jax 0.4.23 arrays do (``_reconstruct_array(fun, args, arr_state,
aval_state)``).  It must be re-checked against a real model file.
"""
from __future__ import annotations

import io
import pickle
from dataclasses import dataclass

import numpy as np


class _Representation:
    """Stand-in for the pickled class objects LCG / VCG (only ``__name__`` is used downstream,
    evaluate_with_given_params.py:78,85)."""


def _make_representation(name: str):
    return type(name, (_Representation,), {})


def _reconstruct_jax_array(fun, args, arr_state, aval_state=None):
    """jax._src.array._reconstruct_array: rebuild the host ndarray, ignore the abstract-value state."""
    arr = fun(*args)
    arr.__setstate__(arr_state)
    return np.asarray(arr)


class _DictLike(dict):
    """Stand-in for haiku's FlatMap / FlatMapping: accepts whatever its reduce passes and keeps dict contents."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for a in args:
            try:
                self.update(dict(a))
            except (TypeError, ValueError):
                pass
        self.update(kwargs)

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.update(state.get("_mapping", state))


class _Stub:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs

    def __setstate__(self, state):
        self.state = state


class _Unpickler(pickle.Unpickler):
    # exact (module, name) pairs an ndarray / scalar / dtype pickle refers to -- nothing else of numpy is reachable
    # (numpy.testing, f2py and distutils hold callables that execute code)
    SAFE_NUMPY = {
        (mod, name)
        for mod in ("numpy.core.multiarray", "numpy._core.multiarray")
        for name in ("_reconstruct", "scalar", "_frombuffer")
    } | {("numpy", "ndarray"), ("numpy", "dtype"),
         ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer")}

    def find_class(self, module, name):
        if (module, name) in self.SAFE_NUMPY:
            return super().find_class(module, name)
        if module == "numpy.dtypes" and name.endswith("DType") and name[:-5].isalnum():
            return super().find_class(module, name)  # dtype classes (numpy >= 1.25 pickles e.g. Float32DType)
        if module in ("builtins", "collections", "functools") and name in (
            "dict", "list", "tuple", "set", "frozenset", "float", "int", "complex", "bool", "str", "bytes", "slice",
            "OrderedDict", "partial", "object", "range",
        ):
            return super().find_class(module, name)
        if module.startswith("jax") or module.startswith("jaxlib"):
            if name == "_reconstruct_array":
                return _reconstruct_jax_array
            return _Stub
        if module.startswith("haiku"):
            return _DictLike
        if name in ("LCG", "VCG", "SATRepresentation"):
            return _make_representation(name)
        return _Stub


@dataclass
class ModelDetails:
    inv_temp: float
    alpha: float
    beta: float
    gamma: float
    mlp_layers: list
    graph_representation: str  # "LCG" / "VCG"
    network_type: str
    return_candidates: bool


def _to_numpy_tree(x):
    if isinstance(x, dict):
        return {str(k): _to_numpy_tree(v) for k, v in x.items()}
    return np.asarray(x)


def load_model(path: str):
    """Returns ``(params, ModelDetails)``; ``params`` is ``{module_name: {"w","b"} | {"scale","offset"}}`` of ndarrays."""
    with open(path, "rb") as f:
        version = np.lib.format.read_magic(f)
        if version == (1, 0):
            shape, _, dtype = np.lib.format.read_array_header_1_0(f)
        else:
            shape, _, dtype = np.lib.format.read_array_header_2_0(f)
        if dtype != np.dtype(object):
            raise ValueError(f"{path}: expected a pickled object array (train.py:221-239), found dtype {dtype}")
        payload = _Unpickler(io.BytesIO(f.read())).load()
    arr = np.asarray(payload, dtype=object)
    if arr.shape[0] != 2:
        raise ValueError(f"{path}: expected [params, details], found an array of shape {arr.shape}")
    params, details = arr[0], list(arr[1])
    if not isinstance(params, dict):
        raise ValueError(f"{path}: parameters are a {type(params).__name__}, expected a mapping of haiku modules")
    rep = details[5]
    rep_name = rep if isinstance(rep, str) else getattr(rep, "__name__", type(rep).__name__)
    md = ModelDetails(float(details[0]), float(details[1]), float(details[2]), float(details[3]),
                      [int(v) for v in details[4]], str(rep_name), str(details[6]), bool(details[7]))
    return _to_numpy_tree(params), md


def load_oracle(path: str, num_message_passing_steps: int = 5):
    """Model file -> ready network (the first half of evaluate_with_given_params.py:66-101)."""
    from .gnn import InteractionNetworkLCG, get_network_definition

    params, md = load_model(path)
    get_network_definition(md.network_type, md.graph_representation)  # raises for anything but interaction + LCG
    return InteractionNetworkLCG(params, num_message_passing_steps), md
