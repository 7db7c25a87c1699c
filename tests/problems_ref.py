"""The reference's own known-answer problems, restated on the Optimizer mirror.

Each function follows a function of reference test/moi_proxsdp_unit.jl (line ranges in
the docstrings) and returns the checks that file asserts.  Variable indices are 0-based.
"""
import numpy as np

from proxsdp_b200 import MAX_SENSE, MIN_SENSE, Optimizer


def simple_lp(opt: Optimizer):
    """test/moi_proxsdp_unit.jl:1-49 (bridged GreaterThan bounds -> Nonpositives rows)."""
    opt.empty()
    X = opt.add_variables(2)
    opt.add_equal_to([(2.0, X[0]), (1.0, X[1])], 4.0)
    opt.add_equal_to([(1.0, X[0]), (2.0, X[1])], 4.0)
    opt.add_greater_than([(1.0, X[0])], 0.0)
    opt.add_greater_than([(1.0, X[1])], 0.0)
    opt.set_objective(MIN_SENSE, [(-4.0, X[0]), (-3.0, X[1])])
    opt.optimize()
    return dict(obj=(opt.objective_value(), -9.33333, 1e-2), x=(opt.variable_primal(X), [1.3333, 1.3333], 1e-2))


def simple_lp_2_1d_sdp(opt: Optimizer):
    """test/moi_proxsdp_unit.jl:51-95: the two bounds as 1x1 PSD cones."""
    opt.empty()
    X = opt.add_variables(2)
    opt.add_equal_to([(2.0, X[0]), (1.0, X[1])], 4.0)
    opt.add_equal_to([(1.0, X[0]), (2.0, X[1])], 4.0)
    opt.add_psd_cone([X[0]])
    opt.add_psd_cone([X[1]])
    opt.set_objective(MIN_SENSE, [(-4.0, X[0]), (-3.0, X[1])])
    opt.optimize()
    return dict(obj=(opt.objective_value(), -9.33333, 1e-2), x=(opt.variable_primal(X), [1.3333, 1.3333], 1e-2))


def lp_in_SDP_equality_form(opt: Optimizer):
    """test/moi_proxsdp_unit.jl:97-138: LP on the diagonal of a 4x4 PSD cone."""
    opt.empty()
    X = opt.add_variables(10)
    opt.add_psd_cone(X)
    opt.add_equal_to([(2.0, X[0]), (1.0, X[2]), (1.0, X[5])], 4.0)
    opt.add_equal_to([(1.0, X[0]), (2.0, X[2]), (1.0, X[9])], 4.0)
    opt.set_objective(MIN_SENSE, [(-4.0, X[0]), (-3.0, X[2])])
    opt.optimize()
    return dict(obj=(opt.objective_value(), -9.33333, 1e-2),
                x=(opt.variable_primal(X), [1.3333, 0, 1.3333, 0, 0, 0, 0, 0, 0, 0], 1e-2))


def lp_in_SDP_inequality_form(opt: Optimizer):
    """test/moi_proxsdp_unit.jl:140-182: Nonpositives rows, MAX sense."""
    opt.empty()
    X = opt.add_variables(3)
    opt.add_psd_cone(X)
    opt.add_nonpositives([(2.0, X[0]), (1.0, X[2])], -4.0)
    opt.add_nonpositives([(1.0, X[0]), (2.0, X[2])], -4.0)
    opt.set_objective(MAX_SENSE, [(4.0, X[0]), (3.0, X[2])])
    opt.optimize()
    return dict(obj=(opt.objective_value(), 9.33333, 1e-2), x=(opt.variable_primal(X), [1.3333, 0.0, 1.3333], 1e-2))


def sdp_from_moi(opt: Optimizer):
    """test/moi_proxsdp_unit.jl:184-223: min X11 + X22 s.t. X21 = 1 -> X = ones, obj 2."""
    opt.empty()
    X = opt.add_variables(3)
    opt.add_psd_cone(X)
    opt.add_zeros([(1.0, X[1])], -1.0)
    opt.set_objective(MIN_SENSE, [(1.0, X[0]), (1.0, X[2])])
    opt.optimize()
    return dict(term=(opt.termination_status(), "OPTIMAL"), pstat=(opt.primal_status(), "FEASIBLE_POINT"),
                dstat=(opt.dual_status(), "FEASIBLE_POINT"), obj=(opt.objective_value(), 2.0, 1e-2),
                x=(opt.variable_primal(X), np.ones(3), 1e-2))


def double_sdp_from_moi(opt: Optimizer):
    """test/moi_proxsdp_unit.jl:225-271: two independent copies -> obj 4."""
    opt.empty()
    X = opt.add_variables(3)
    Y = opt.add_variables(3)
    opt.add_psd_cone(X)
    opt.add_psd_cone(Y)
    opt.add_zeros([(1.0, X[1])], -1.0)
    opt.add_zeros([(1.0, Y[1])], -1.0)
    opt.set_objective(MIN_SENSE, [(1.0, X[0]), (1.0, X[2]), (1.0, Y[0]), (1.0, Y[2])])
    opt.optimize()
    return dict(term=(opt.termination_status(), "OPTIMAL"), pstat=(opt.primal_status(), "FEASIBLE_POINT"),
                dstat=(opt.dual_status(), "FEASIBLE_POINT"), obj=(opt.objective_value(), 4.0, 1e-2),
                x=(opt.variable_primal(X), np.ones(3), 1e-2), y=(opt.variable_primal(Y), np.ones(3), 1e-2))


def sdp_wiki(opt: Optimizer, sense=MIN_SENSE):
    """test/moi_proxsdp_unit.jl:302-338: the Wikipedia SDP example, -0.978 (Min) / 0.872 (Max)."""
    opt.empty()
    X = opt.add_variables(6)
    opt.add_psd_cone(X)
    opt.add_zeros([(1.0, X[0])], -1.0)
    opt.add_zeros([(1.0, X[2])], -1.0)
    opt.add_zeros([(1.0, X[5])], -1.0)
    opt.add_nonpositives([(1.0, X[1])], 0.1)
    opt.add_nonpositives([(-1.0, X[1])], -0.2)
    opt.add_nonpositives([(1.0, X[4])], -0.5)
    opt.add_nonpositives([(-1.0, X[4])], 0.4)
    opt.set_objective(sense, [(1.0, X[3])])
    opt.optimize()
    return dict(obj=(opt.objective_value(), -0.978 if sense == MIN_SENSE else 0.872, 1e-2))


ALL = [simple_lp, simple_lp_2_1d_sdp, lp_in_SDP_equality_form, lp_in_SDP_inequality_form, sdp_from_moi,
       double_sdp_from_moi]


def check(results):
    for key, tup in results.items():
        if len(tup) == 2:
            assert tup[0] == tup[1], (key, tup)
        else:
            got, want, atol = tup
            assert np.allclose(got, want, atol=atol), (key, got, want)
