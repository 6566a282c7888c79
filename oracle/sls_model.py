"""A second, deliberately naive model of the reference solver -- TEST INFRASTRUCTURE ONLY.

``oracle/sls_oracle.c`` is the restatement the CUDA kernels are compared with bit for bit; it replaces the
reference's data structures by ones a GPU can share (bitmap over canonical clause ids, Philox counter streams,
rank select in clause order).  This module keeps the reference's own shapes instead -- a hash set of violated
clause *contents*, a list of clauses per variable, Python's Mersenne Twister -- and shares no code, no random
stream and no element order with the C oracle.  Nothing can therefore be compared run by run; what the tests
compare (tests/test_oracle_golden.py) are the *distributions* the two produce on the same instance and weights:
solve fraction, number of steps, best energy.  Agreement says that the C oracle's substitutions (which the CUDA
path inherits exactly) do not change the stochastic process of rust/src/lib.rs.

Followed, function by function: rust/src/lib.rs:13-18 (violation), :20-49 (Moser-Tardos resampling), :51-55
(Schoening), :57-75 (probSAT), :77-101 (single-flip Moser-Tardos), :165-199 (setup), :201-317 (one chain),
:320-334 (reduction over chains).  Pure Python, small cases only.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field


def violated(clause, assignment) -> bool:
    """lib.rs:13-18: every literal is false (an empty clause is violated)."""
    return all((lit > 0) != assignment[abs(lit) - 1] for lit in clause)


def _flip_probability(lit, assignment, weights) -> float:
    """lib.rs:28-30: w if the variable is False, 1 - w if it is True."""
    v = abs(lit) - 1
    return 1.0 - weights[v] if assignment[v] else weights[v]


def _marked(clause, assignment, rng, weights):
    """lib.rs:24-35 / :85-93: one uniform draw per literal, in clause order."""
    return [abs(lit) - 1 for lit in clause if rng.random() < _flip_probability(lit, assignment, weights)]


def rule_moser(clause, assignment, rng, weights):
    flipped = _marked(clause, assignment, rng, weights)
    for v in flipped:
        assignment[v] = not assignment[v]
    return flipped


def rule_moser_single(clause, assignment, rng, weights):
    flipped = _marked(clause, assignment, rng, weights)
    if flipped:
        v = flipped[rng.randrange(len(flipped))]
        assignment[v] = not assignment[v]
        return [v]
    return []


def rule_schoening(clause, assignment, rng, weights):
    v = abs(rng.choice(clause)) - 1
    assignment[v] = not assignment[v]
    return [v]


def rule_probsat(clause, assignment, rng, weights):
    w = [_flip_probability(lit, assignment, weights) for lit in clause]
    if any(x < 0 for x in w) or sum(w) == 0:
        raise ValueError("WeightedIndex: invalid weights")  # lib.rs:71 unwrap()
    v = abs(rng.choices(clause, weights=w)[0]) - 1
    assignment[v] = not assignment[v]
    return [v]


RULES = {"moser": rule_moser, "moser_single": rule_moser_single, "schoening": rule_schoening, "probsat": rule_probsat}


@dataclass
class ModelResult:
    best_assignment: list
    found: list
    best_energy: int
    steps: list
    trajectories: list = field(default_factory=list)
    best_energy_run: list = field(default_factory=list)


def run_sls(algo, n_vars, clauses, weights_initialize, weights_resample, nsteps, nruns, return_trajectories=False,
            pre_compute_mapping=True, prob_flip_best=0.0, seed=0) -> ModelResult:
    """lib.rs:165-335 with clauses as tuples of DIMACS literals."""
    rule = RULES[algo]  # KeyError = the reference's panic on an unknown algorithm (lib.rs:125)
    clauses = [tuple(c) for c in clauses]

    def find_violated(assignment):
        return {c for c in clauses if violated(c, assignment)}  # a set of clause contents, lib.rs:176-181

    var_to_clauses = [[] for _ in range(n_vars)]
    if pre_compute_mapping:
        for c in clauses:
            for lit in c:
                var_to_clauses[abs(lit) - 1].append(c)

    out = ModelResult([], [], len(clauses), [])
    for run in range(nruns):
        rng = random.Random(seed * 1_000_003 + run)
        assignment = [rng.random() < weights_initialize[i] for i in range(n_vars)]
        violated_set = find_violated(assignment)
        best_run, best_assignment_run = len(clauses), []
        trajectory = []
        numfalse = len(violated_set)
        if numfalse < best_run:
            best_run, best_assignment_run = numfalse, list(assignment)
        if return_trajectories:
            trajectory.append(numfalse)
        step = 0
        while numfalse > 0 and step < nsteps:
            step += 1
            clause = rng.choice(sorted(violated_set))  # uniform over the set; the reference's order is its hash order
            if rng.random() < prob_flip_best:
                best_v, fewest = 0, None
                for lit in clause:
                    v = abs(lit) - 1
                    assignment[v] = not assignment[v]
                    if pre_compute_mapping:
                        score = sum(violated(c, assignment) for c in var_to_clauses[v])
                    else:
                        score = len(find_violated(assignment))
                    assignment[v] = not assignment[v]
                    if fewest is None or score < fewest:
                        fewest, best_v = score, v
                assignment[best_v] = not assignment[best_v]
                flipped = [best_v]
            else:
                flipped = rule(clause, assignment, rng, weights_resample)
            if pre_compute_mapping:
                for v in flipped:
                    for c in var_to_clauses[v]:
                        if violated(c, assignment):
                            violated_set.add(c)
                        else:
                            violated_set.discard(c)
            else:
                violated_set = find_violated(assignment)
            numfalse = len(violated_set)
            if numfalse < best_run:
                best_run, best_assignment_run = numfalse, list(assignment)
            if return_trajectories:
                trajectory.append(numfalse)
        if best_run < out.best_energy:
            out.best_energy, out.best_assignment = best_run, best_assignment_run
        out.found.append(numfalse == 0)
        out.steps.append(step)
        out.best_energy_run.append(best_run)
        if return_trajectories:
            out.trajectories.append(trajectory)
    return out
