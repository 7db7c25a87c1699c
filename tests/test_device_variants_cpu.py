"""CPU checks of the arithmetic variants the device kernels use where they deviate from the oracle's literal
restatement (both scripts restate the device arithmetic in numpy / Python floats):

* lanczos_cl3.cuh re-orthogonalises with the local three-term step + ONE classical Gram-Schmidt pass (KrylovKit and
  the oracle: local step + two modified Gram-Schmidt passes) -> scripts/lz_variant_check.py compares mat-vec counts,
  converged counts, Ritz values and basis orthogonality with the C oracle on 13 matrix families.
* ritz_bi.cuh counts eigenvalues with a pre-scaled, division-free Sturm recurrence rescaled every 8 rows
  -> scripts/sturm_check.py compares the counts with LAPACK on random / graded / clustered tridiagonals.
* ritz_bi.cuh finds Ritz values by secant-steered clustered multisection, and lanczos_cl3.cuh turns the bordered Rayleigh
  quotient of a thick restart back into a tridiagonal by a small Lanczos run with a DGKS loop and a relative breakdown test
  (KrylovKit: dense `tridiageigh!` and a Householder reduction) -> scripts/ritz_restart_check.py checks both against LAPACK,
  the restart on repeated values and on zero / negligible / graded couplings.
* runtime.cu `k_equilibrate` walks equilibrate! as R scalar recurrences coupled by one sum (the reference's column
  scaling is a multiple of the identity by construction) -> scripts/equilibrate_check.py compares E and D with the literal
  restatement (oracle/oracle_np.py) on sparse matrices with badly scaled rows.
"""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_one_pass_reorthogonalisation_matches_oracle_counts():
    assert _load("lz_variant_check").main() == 0


def test_scaled_sturm_count_matches_lapack():
    assert _load("sturm_check").main() == 0


def test_ritz_values_and_retridiagonalising_restart_match_lapack():
    assert _load("ritz_restart_check").main() == 0


def test_collapsed_equilibration_matches_literal_restatement():
    assert _load("equilibrate_check").main() == 0
