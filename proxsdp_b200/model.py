"""A small MathOptInterface-shaped front end: `Optimizer`.

The reference's user surface is `ProxSDP.Optimizer` (src/MOI_wrapper.jl:54-74), which
assembles `AffineSets` + `ConicSets` in `_optimize!` (src/MOI_wrapper.jl:220-342) and
calls `chambolle_pock` (line 310).  Julia / MOI are not available in this image, so this
module mirrors just enough of that surface for the parity tests to read like the
reference's own (`test/moi_proxsdp_unit.jl`): add variables, declare variable cones
(`VectorOfVariables`-in-PSDTriangle / -in-SOC), add `VectorAffineFunction`-in-`Zeros` /
-in-`Nonpositives` rows, set a linear objective with a sense, optimise, query.

Conventions restated from `_optimize!`:
  * `b = -constants` of the Zeros block, `h = -constants` of the Nonpositives block
    (MOI_wrapper.jl:235,242);
  * `c` is sign-flipped for MAX_SENSE and the objective values flipped back afterwards
    (MOI_wrapper.jl:247-255,336-337);
  * constraint duals are returned negated (MOI_wrapper.jl:502,511).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

from .options import Options
from .structs import AffineSets, ConicSets, Result, SDPSet, SOCSet, sympackeddim

MIN_SENSE = "MIN_SENSE"
MAX_SENSE = "MAX_SENSE"

# MOI_wrapper.jl:381-399
_TERMINATION = {
    0: "OPTIMIZE_NOT_CALLED", 1: "OPTIMAL", 2: "TIME_LIMIT", 3: "ITERATION_LIMIT",
    4: "INFEASIBLE_OR_UNBOUNDED", 5: "DUAL_INFEASIBLE", 6: "INFEASIBLE",
}

Term = Tuple[float, int]  # (coefficient, variable)


def _default_backend():
    from .solver import chambolle_pock
    return chambolle_pock


class Optimizer:
    """`ProxSDP.Optimizer(; kwargs...)` (MOI_wrapper.jl:62-74)."""

    def __init__(self, backend: Optional[Callable] = None, **kwargs):
        self.options = Options(**kwargs)
        self._backend = backend
        self.empty()

    # ---- model lifetime -------------------------------------------------
    def empty(self):
        self.nvars = 0
        self._eq_rows: List[Tuple[List[Term], float]] = []   # (terms, constant): terms + constant == 0
        self._in_rows: List[Tuple[List[Term], float]] = []   # terms + constant <= 0
        self._psd: List[np.ndarray] = []
        self._soc: List[np.ndarray] = []
        self._obj_terms: List[Term] = []
        self._obj_const = 0.0
        self._sense = MIN_SENSE
        self.sol = Result()
        self.aff: Optional[AffineSets] = None
        self.con: Optional[ConicSets] = None

    def is_empty(self) -> bool:
        return self.nvars == 0 and not self._eq_rows and not self._in_rows

    # ---- options (MOI_wrapper.jl:84-139) --------------------------------
    def set_attribute(self, name: str, value):
        self.options.set(name, value)

    def get_attribute(self, name: str):
        return self.options.get(name)

    def set_silent(self, flag: bool):
        self.options.log_verbose = not flag

    def set_time_limit_sec(self, value: Optional[float]):
        self.options.time_limit = 3600_00.0 if value is None else float(value)

    def get_time_limit_sec(self) -> Optional[float]:
        return None if self.options.time_limit == 3600_00.0 else self.options.time_limit

    # ---- variables and cones --------------------------------------------
    def add_variables(self, k: int) -> List[int]:
        out = list(range(self.nvars, self.nvars + k))
        self.nvars += k
        return out

    def add_variable(self) -> int:
        return self.add_variables(1)[0]

    def add_psd_cone(self, variables: Sequence[int]) -> int:
        """VectorOfVariables-in-PositiveSemidefiniteConeTriangle (column-major upper triangle)."""
        sympackeddim(len(variables))
        self._psd.append(np.asarray(variables, dtype=np.int64))
        return len(self._psd) - 1

    def add_soc_cone(self, variables: Sequence[int]) -> int:
        """VectorOfVariables-in-SecondOrderCone; variables[0] is t."""
        self._soc.append(np.asarray(variables, dtype=np.int64))
        return len(self._soc) - 1

    def add_psd_variable(self, side: int) -> Tuple[np.ndarray, int]:
        """`@variable(model, X[1:n,1:n], PSD)`: returns an n x n index matrix and the cone id."""
        tri = self.add_variables(side * (side + 1) // 2)
        idx = np.zeros((side, side), dtype=np.int64)
        c = 0
        for j in range(side):
            for i in range(j + 1):
                idx[i, j] = idx[j, i] = tri[c]
                c += 1
        return idx, self.add_psd_cone(tri)

    # ---- affine constraints ---------------------------------------------
    def add_zeros(self, terms: Sequence[Term], constant: float) -> int:
        """One row of VectorAffineFunction-in-Zeros: sum(coef*x) + constant == 0."""
        self._eq_rows.append((list(terms), float(constant)))
        return len(self._eq_rows) - 1

    def add_nonpositives(self, terms: Sequence[Term], constant: float) -> int:
        """One row of VectorAffineFunction-in-Nonpositives: sum(coef*x) + constant <= 0."""
        self._in_rows.append((list(terms), float(constant)))
        return len(self._in_rows) - 1

    # bridged scalar forms
    def add_equal_to(self, terms, rhs):
        return self.add_zeros(terms, -rhs)

    def add_less_than(self, terms, rhs):
        return self.add_nonpositives(terms, -rhs)

    def add_greater_than(self, terms, rhs):
        return self.add_nonpositives([(-c, v) for c, v in terms], rhs)

    # ---- objective ------------------------------------------------------
    def set_objective(self, sense: str, terms: Sequence[Term], constant: float = 0.0):
        assert sense in (MIN_SENSE, MAX_SENSE)
        self._sense = sense
        self._obj_terms = list(terms)
        self._obj_const = float(constant)

    # ---- _optimize! (MOI_wrapper.jl:220-342) ----------------------------
    def _rows_to_csc(self, rows, n):
        ri, ci, vals, consts = [], [], [], []
        for r, (terms, const) in enumerate(rows):
            for coef, var in terms:
                ri.append(r)
                ci.append(var)
                vals.append(coef)
            consts.append(const)
        M = sp.coo_matrix((vals, (ri, ci)), shape=(len(rows), n)).tocsc()
        M.sum_duplicates()
        return M, -np.asarray(consts, dtype=np.float64)

    def build(self) -> Tuple[AffineSets, ConicSets]:
        n = self.nvars
        A, b = self._rows_to_csc(self._eq_rows, n)
        G, h = self._rows_to_csc(self._in_rows, n)
        obj_sign = -1.0 if self._sense == MAX_SENSE else 1.0
        c = np.zeros(n)
        for coef, var in self._obj_terms:
            c[var] += obj_sign * coef
        aff = AffineSets(n, A.shape[0], G.shape[0], 0, A, G, b, h, c)
        con = ConicSets()
        for soc in self._soc:
            con.socone.append(SOCSet(soc, len(soc)))
        for psc in self._psd:
            con.sdpcone.append(SDPSet(psc, len(psc), sympackeddim(len(psc))))
        return aff, con

    def optimize(self):
        aff, con = self.build()
        self.aff, self.con = aff, con
        backend = self._backend or _default_backend()
        sol = backend(aff, con, self.options)
        obj_sign = -1.0 if self._sense == MAX_SENSE else 1.0
        sol.objval = obj_sign * sol.objval + self._obj_const
        sol.dual_objval = obj_sign * sol.dual_objval + self._obj_const
        self.sol = sol
        return sol

    # ---- attributes (MOI_wrapper.jl:361-530) ----------------------------
    def termination_status(self) -> str:
        return _TERMINATION[self.sol.status]

    def raw_status_string(self) -> str:
        return self.sol.status_string

    def primal_status(self) -> str:
        s = self.sol.status
        if s == 0:
            return "NO_SOLUTION"
        if s == 5 and self.sol.certificate_found:
            return "INFEASIBILITY_CERTIFICATE"
        return "FEASIBLE_POINT" if self.sol.primal_feasible_user_tol else "INFEASIBLE_POINT"

    def dual_status(self) -> str:
        s = self.sol.status
        if s == 0:
            return "NO_SOLUTION"
        if s == 6 and self.sol.certificate_found:
            return "INFEASIBILITY_CERTIFICATE"
        return "FEASIBLE_POINT" if self.sol.dual_feasible_user_tol else "INFEASIBLE_POINT"

    def objective_value(self) -> float:
        return self.sol.objval

    def dual_objective_value(self) -> float:
        return self.sol.dual_objval

    def solve_time_sec(self) -> float:
        return self.sol.time

    def pdhg_iterations(self) -> int:
        return int(self.sol.iter)

    def result_count(self) -> int:
        return self.sol.result_count

    def variable_primal(self, variables):
        return self.sol.primal[np.asarray(variables, dtype=np.int64)]

    def constraint_primal_zeros(self, row):
        return self.sol.slack_eq[row]

    def constraint_primal_nonpositives(self, row):
        return self.sol.slack_in[row]

    def constraint_dual_zeros(self, row):
        return -self.sol.dual_eq[row]

    def constraint_dual_nonpositives(self, row):
        return -self.sol.dual_in[row]

    def constraint_dual_psd(self, cone):
        return self.sol.dual_cone[self._psd[cone]]

    def constraint_dual_soc(self, cone):
        return self.sol.dual_cone[self._soc[cone]]
