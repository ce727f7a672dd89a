"""Stoichiometric matrix of a reaction list (reference: catalax/model/stoich.py:9-59).

The mechanistic kernels never read this matrix -- ``tools/codegen.FieldSpec.total_exprs`` folds
the f32-rounded coefficients into the generated right-hand side -- but it is the reference's
public layout contract (rows = states in the given order, columns = reactions sorted by symbol)
and the input of ``RateFlowODE``."""
from __future__ import annotations

from typing import List

import numpy as np


def derive_stoich_matrix(reactions: List, state_order: List[str]) -> np.ndarray:
    """S[i, j] = net coefficient of ``state_order[i]`` in the j-th reaction (by sorted symbol):
    negative for reactants, positive for products, f32 like the reference's matrix."""
    row = {s: i for i, s in enumerate(state_order)}
    ordered = sorted(reactions, key=lambda r: r.symbol)
    S = np.zeros((len(state_order), len(ordered)), dtype=np.float32)
    for j, r in enumerate(ordered):
        for e in r.reactants:
            if e.state in row:
                S[row[e.state], j] -= e.stoichiometry
        for e in r.products:
            if e.state in row:
                S[row[e.state], j] += e.stoichiometry
    return S
