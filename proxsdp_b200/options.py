"""Options — Python mirror of the reference's `Options` keyword struct.

Same field names, order and defaults as reference src/options.jl:1-132.  The
flat C image is `proxsdp_options_t` (include/proxsdp_b200_types.h); every field
travels as an 8-byte int64/double, Julia Bools as 0/1.  Unknown option names raise,
like `MOI.set(::Optimizer, ::RawOptimizerAttribute, …)` does in the reference
(src/MOI_wrapper.jl:84-103).
"""
from __future__ import annotations

import ctypes

# (name, kind, default)   kind: 'b' Bool, 'i' Int, 'f' Float64
OPTION_FIELDS = [
    ("log_verbose", "b", False),
    ("log_freq", "i", 1000),
    ("timer_verbose", "b", False),
    ("timer_file", "b", False),
    ("disable_julia_logger", "b", True),
    ("time_limit", "f", 3600_00.0),
    ("warn_on_limit", "b", False),
    ("extended_log", "b", False),
    ("extended_log2", "b", False),
    ("log_repeat_header", "b", False),
    ("tol_gap", "f", 1e-4),
    ("tol_feasibility", "f", 1e-4),
    ("tol_feasibility_dual", "f", 1e-4),
    ("tol_primal", "f", 1e-4),
    ("tol_dual", "f", 1e-4),
    ("tol_psd", "f", 1e-7),
    ("tol_soc", "f", 1e-7),
    ("check_dual_feas", "b", False),
    ("check_dual_feas_freq", "i", 1000),
    ("max_obj", "f", 1e20),
    ("min_iter_max_obj", "i", 10),
    ("min_iter_time_infeas", "i", 1000),
    ("infeas_gap_tol", "f", 1e-4),
    ("infeas_limit_gap_tol", "f", 1e-1),
    ("infeas_stable_gap_tol", "f", 1e-4),
    ("infeas_feasibility_tol", "f", 1e-4),
    ("infeas_stable_feasibility_tol", "f", 1e-8),
    ("certificate_search", "b", True),
    ("certificate_obj_tol", "f", 1e-1),
    ("certificate_fail_tol", "f", 1e-8),
    ("min_beta", "f", 1e-5),
    ("max_beta", "f", 1e5),
    ("initial_beta", "f", 1.0),
    ("initial_adapt_level", "f", 0.9),
    ("adapt_decay", "f", 0.8),
    ("adapt_window", "i", 50),
    ("convergence_window", "i", 200),
    ("convergence_check", "i", 50),
    ("max_iter", "i", 0),
    ("min_iter", "i", 40),
    ("divergence_min_update", "i", 50),
    ("max_iter_lp", "i", 10_000_000),
    ("max_iter_conic", "i", 1_000_000),
    ("max_iter_local", "i", 0),
    ("advanced_initialization", "b", True),
    ("line_search_flag", "b", True),
    ("max_linsearch_steps", "i", 5000),
    ("delta", "f", 0.9999),
    ("initial_theta", "f", 1.0),
    ("linsearch_decay", "f", 0.75),
    ("full_eig_decomp", "b", False),
    ("max_target_rank_krylov_eigs", "i", 16),
    ("min_size_krylov_eigs", "i", 100),
    ("warm_start_eig", "b", True),
    ("rank_increment", "i", 1),
    ("rank_increment_factor", "i", 1),
    ("eigsolver", "i", 2),
    ("eigsolver_min_lanczos", "i", 25),
    ("eigsolver_resid_seed", "i", 1234),
    ("arpack_tol", "f", 1e-10),
    ("arpack_resid_init", "i", 3),
    ("arpack_reset_resid", "b", True),
    ("arpack_max_iter", "i", 10_000),
    ("krylovkit_reset_resid", "b", False),
    ("krylovkit_resid_init", "i", 3),
    ("krylovkit_tol", "f", 1e-12),
    ("krylovkit_max_iter", "i", 100),
    ("krylovkit_eager", "b", False),
    ("krylovkit_verbose", "i", 0),
    ("reduce_rank", "b", False),
    ("rank_slack", "i", 3),
    ("full_eig_freq", "i", 10_000_000),
    ("full_eig_len", "i", 0),
    ("equilibration", "b", False),
    ("equilibration_iters", "i", 1000),
    ("equilibration_lb", "f", -10.0),
    ("equilibration_ub", "f", 10.0),
    ("equilibration_limit", "f", 0.9),
    ("equilibration_force", "b", False),
    ("approx_norm", "b", True),
    # ---- extensions (not in options.jl; zero = reference behaviour) ----
    ("initial_target_rank", "i", 0),
    ("freeze_target_rank", "b", False),
    ("device_id", "i", 0),
    ("trace_cap", "i", 0),
    ("implicit_psd_operator", "b", False),
]

N_REFERENCE_FIELDS = 80  # options.jl has 80 fields; the rest are extensions


class OptionsPOD(ctypes.Structure):
    _fields_ = [
        (name, ctypes.c_double if kind == "f" else ctypes.c_int64) for name, kind, _ in OPTION_FIELDS
    ]


class Options:
    """Keyword struct: `Options(tol_gap=1e-5, max_iter=100)`."""

    __slots__ = [name for name, _, _ in OPTION_FIELDS]

    def __init__(self, **kwargs):
        for name, _, default in OPTION_FIELDS:
            object.__setattr__(self, name, default)
        for k, v in kwargs.items():
            self.set(k, v)

    def set(self, name: str, value) -> None:
        if name not in self.__slots__:
            # MOI_wrapper.jl:90,101
            raise ValueError(f"Option {name} is not valid.")
        object.__setattr__(self, name, value)

    def get(self, name: str):
        if name not in self.__slots__:
            raise ValueError(f"Option {name} is not valid.")
        return getattr(self, name)

    def __setattr__(self, name, value):
        self.set(name, value)

    def copy(self) -> "Options":
        return Options(**{n: getattr(self, n) for n in self.__slots__})

    def to_pod(self) -> OptionsPOD:
        pod = OptionsPOD()
        for name, kind, _ in OPTION_FIELDS:
            v = getattr(self, name)
            setattr(pod, name, float(v) if kind == "f" else int(v))
        return pod

    def __repr__(self):
        changed = {n: getattr(self, n) for n, _, d in OPTION_FIELDS if getattr(self, n) != d}
        return f"Options({', '.join(f'{k}={v!r}' for k, v in changed.items())})"
