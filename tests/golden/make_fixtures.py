"""Regenerates tests/golden/ from the reference checkout (run in the build container only).

Copies a small, fixed subset of the reference's shipped instances together with their known
satisfying assignments (``*_sol.pkl``) and sampled candidate assignments (``*_samples_sol.npy``),
and records the expected per-sample energies.  The reference stores no expected energies, so they
are computed twice here -- by the C oracle (oracle/sls_oracle.c) and by an independent NumPy
evaluation written in this file -- and must agree.  What the reference itself pins is
``energy(solution) == 0`` (python/src/data_utils.py:130) and the three known-answer tests of
python/tests/test_sat_instances.py:69-135 (restated in tests/test_oracle_golden.py).

usage: python tests/golden/make_fixtures.py [/root/reference]
"""
import glob
import json
import os
import pickle
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import sls_oracle  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
MAX_SAMPLES = 64


def load_solution(path):
    s = pickle.load(open(path, "rb"))
    if isinstance(s, dict):  # nnf/kissat dict format (python/src/data_utils.py:108-111)
        s = [x for _, x in sorted(s.items())]
        return np.asarray(s, dtype=bool).astype(np.uint8)
    s = np.asarray(s)
    if (np.abs(s) >= 2).any() or (s < 0).any():  # signed literal list (data_utils.py:112-116)
        return ((np.sign(s) + 1) // 2).astype(np.uint8)
    return s.astype(np.uint8)


def numpy_energy(n, clauses, a):
    """Independent evaluation: a clause is violated iff no literal is satisfied."""
    e = 0
    viol = []
    for cl in clauses:
        sat = any((a[abs(l) - 1] == 1) == (l > 0) for l in cl)
        viol.append(0 if sat else 1)
        e += 0 if sat else 1
    return e, viol


def read_dimacs(path):
    n = 0
    clauses = []
    cur = []
    for line in open(path):
        if line.startswith("c"):
            continue
        if line.startswith("p"):
            n = int(line.split()[2])
            continue
        for tok in line.split():
            v = int(tok)
            if v == 0:
                clauses.append(cur)
                cur = []
            else:
                cur.append(v)
    return n, clauses


def pick():
    out = []
    t = os.path.join(REF, "python/tests/test_instances")
    out.append(("ref_tests_medium", os.path.join(t, "single_instance/medium")))
    out.append(("ref_tests_huge", os.path.join(t, "multiple_instances/huge")))
    test = sorted(glob.glob(os.path.join(REF, "Data/random_sat_data/test/*.cnf")))
    for i in np.linspace(0, len(test) - 1, 10).astype(int):
        out.append(("r3sat_test_" + os.path.basename(test[i])[:-4], test[i][:-4]))
    train = sorted(glob.glob(os.path.join(REF, "Data/random_sat_data/training/*.cnf")))
    for i in (0, len(train) // 2):
        out.append(("r3sat_train_" + os.path.basename(train[i])[:-4], train[i][:-4]))
    for d in ("easy", "medium", "hard"):
        for k in ("ca", "ps"):
            files = sorted(glob.glob(os.path.join(REF, f"Data/G4SAT/{d}/{k}/test/sat/*.cnf")), key=os.path.getsize)
            out.append((f"g4sat_{d}_{k}_" + os.path.basename(files[0])[:-4], files[0][:-4]))
    return out


def main():
    inst_dir = os.path.join(HERE, "instances")
    shutil.rmtree(inst_dir, ignore_errors=True)
    os.makedirs(inst_dir)
    expected = {}
    for name, stem in pick():
        shutil.copyfile(stem + ".cnf", os.path.join(inst_dir, name + ".cnf"))
        os.chmod(os.path.join(inst_dir, name + ".cnf"), 0o644)
        n, clauses = read_dimacs(stem + ".cnf")
        oi = sls_oracle.Instance.from_dimacs(stem + ".cnf")
        assert (oi.n, oi.m) == (n, len(clauses))
        sol = load_solution(stem + "_sol.pkl")
        assert len(sol) == n
        rows = [sol]
        if os.path.exists(stem + "_samples_sol.npy"):
            samples = np.load(stem + "_samples_sol.npy").astype(np.uint8)
            assert (samples[0] == sol).all()  # row 0 is the solution
            rows += list(samples[1:MAX_SAMPLES])
        rows = np.asarray(rows, dtype=np.uint8)
        energies = []
        for a in rows:
            e_np, viol_np = numpy_energy(n, clauses, a)
            assert e_np == oi.energy(a) and viol_np == oi.eval_clauses(a).tolist()
            energies.append(int(e_np))
        assert energies[0] == 0  # data_utils.py:130
        np.save(os.path.join(inst_dir, name + "_assignments.npy"), np.packbits(rows, axis=1, bitorder="little"))
        expected[name] = {"n": n, "m": len(clauses), "num_lits": int(sum(len(c) for c in clauses)),
                          "rows": int(rows.shape[0]), "energies": energies,
                          "source": os.path.relpath(stem + ".cnf", REF)}
        print(name, n, len(clauses), rows.shape[0], energies[:5])
    json.dump(expected, open(os.path.join(HERE, "expected.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
