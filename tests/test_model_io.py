"""Model-file reader.  The shipped .npy files are absent, so a file is produced here
with stand-in `jax` / reference modules that pickle like the real ones (jax arrays reduce to
`jax._src.array._reconstruct_array(fun, args, arr_state, aval_state)`; the graph representation is a
class pickled by reference), then the stand-ins are removed and the file is read back."""
import sys
import types

import numpy as np

from oracle import gnn_oracle as go
from sls_sat_solving_with_deep_learning_b200 import model_io


def _install_fakes():
    jax_array = types.ModuleType("jax._src.array")

    def _reconstruct_array(fun, args, arr_state, aval_state):  # pragma: no cover (only pickled by reference)
        raise RuntimeError("the fake must never be called by the loader")

    class ArrayImpl:
        def __init__(self, a):
            self.a = np.asarray(a)

        def __reduce__(self):
            fun, args, arr_state = self.a.__reduce__()
            return (_reconstruct_array, (fun, args, arr_state, {"weak_type": False}))

    _reconstruct_array.__module__ = "jax._src.array"
    _reconstruct_array.__qualname__ = "_reconstruct_array"
    jax_array._reconstruct_array = _reconstruct_array
    jax_array.ArrayImpl = ArrayImpl
    rep = types.ModuleType("python.src.sat_representations")
    LCG = type("LCG", (), {"__module__": "python.src.sat_representations"})
    rep.LCG = LCG
    mods = {"jax": types.ModuleType("jax"), "jax._src": types.ModuleType("jax._src"), "jax._src.array": jax_array,
            "python": types.ModuleType("python"), "python.src": types.ModuleType("python.src"),
            "python.src.sat_representations": rep}
    sys.modules.update(mods)
    return ArrayImpl, LCG, list(mods)


def test_load_model_without_jax(tmp_path):
    params = go.init_params(seed=1, mlp_layers=(32, 32), num_steps=2, bias_scale=0.1)
    ArrayImpl, LCG, names = _install_fakes()
    try:
        jparams = {k: {kk: ArrayImpl(vv) for kk, vv in v.items()} for k, v in params.items()}
        details = [2.0, 1.0, 0.5, 0.0, [32, 32], LCG, "interaction", False]  # train.py:221-239
        path = str(tmp_path / "model.npy")
        np.save(path, np.asarray([jparams, details], dtype=object))
    finally:
        for n in names:
            sys.modules.pop(n, None)
    assert "jax" not in sys.modules
    got, md = model_io.load_model(path)
    assert md.graph_representation == "LCG" and md.network_type == "interaction" and md.mlp_layers == [32, 32]
    assert (md.inv_temp, md.alpha, md.beta, md.gamma, md.return_candidates) == (2.0, 1.0, 0.5, 0.0, False)
    assert sorted(got) == sorted(params)
    for k in params:
        for kk in params[k]:
            assert isinstance(got[k][kk], np.ndarray) and (got[k][kk] == params[k][kk]).all()
    # the blob packer accepts it
    from sls_sat_solving_with_deep_learning_b200 import gnn

    blob, D, h1, h2 = gnn.pack_params(got, 2)
    assert (D, h1, h2) == (32, 32, 32) and blob.dtype == np.float32


def test_random_oracle_params_have_the_reference_layout():
    # synth.random_oracle_params is bench.py's stand-in for the absent Data/models files: haiku names in construction order,
    # shapes of the interaction network with mlp_layers=[200, 200] (evaluate_with_given_params.py:97-101)
    from sls_sat_solving_with_deep_learning_b200 import gnn, synth

    params = synth.random_oracle_params(seed=1)
    blob, embed, h1, h2 = gnn.pack_params(params, 5)
    assert (embed, h1, h2) == (32, 200, 200)
    assert blob.dtype == np.float32 and blob.size == 1_473_993  # 1.47 M parameters
    assert set(params) == {"linear"} | {f"linear_{i}" for i in range(1, 23)} | {"layer_norm"} | {f"layer_norm_{i}" for i in range(1, 10)}
    assert np.abs(params["linear_2"]["w"]).max() <= 2.0 / np.sqrt(96) + 1e-6


def test_pack_params_maps_modules_by_shape_and_order_when_names_differ():
    # the haiku names are from memory -- a file with another scope or numbering must still load, by creation
    # order and shape, and a wrong guess must be an error
    from oracle import gnn_oracle as go
    from sls_sat_solving_with_deep_learning_b200 import gnn

    params = go.init_params(seed=3, mlp_layers=(32, 32), num_steps=2, bias_scale=0.1)
    blob, D, h1, h2 = gnn.pack_params(params, 2)
    # (1) scoped names, keys sorted alphabetically (what a jax tree operation leaves behind)
    scoped = {f"net/~/{k}": v for k, v in sorted(params.items())}
    blob2, *_ = gnn.pack_params(scoped, 2)
    assert np.array_equal(blob, blob2)
    # (2) another numbering scheme: creation order survives as the numeric suffix
    lin = sorted((k for k in params if "w" in params[k]), key=gnn._natural_key)
    lns = sorted((k for k in params if "scale" in params[k]), key=gnn._natural_key)
    renamed = {f"dense_{i + 3}": params[k] for i, k in enumerate(lin)}
    renamed.update({f"norm_{i}": params[k] for i, k in enumerate(lns)})
    blob3, *_ = gnn.pack_params(dict(sorted(renamed.items())), 2)
    assert np.array_equal(blob, blob3)
    # (3) a module too few: refused
    import pytest

    broken = dict(renamed)
    broken.pop("norm_1")
    with pytest.raises(ValueError):
        gnn.pack_params(broken, 2)
    # (4) right count, wrong order (two layers swapped): the shape check catches it
    swapped = dict(renamed)
    swapped["dense_5"], swapped["dense_6"] = swapped["dense_6"], swapped["dense_5"]
    with pytest.raises(ValueError):
        gnn.pack_params(swapped, 2)
