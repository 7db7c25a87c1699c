"""CPU test of the log `chambolle_pock` writes under `log_verbose` (reference src/printing.jl): the host-only header
proxsdp_b200/csrc/printing.hpp is compiled alone with g++ and its lines are compared with what the Julia source produces
for the same numbers (Julia's `show(::Float64)`, `round(x; digits)`, `Printf.@sprintf` and the column padding)."""
import os
import subprocess
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
#include "printing.hpp"
using namespace pb::plog;
int main() {
    const double xs[] = {1e-4, 1e-5, 1e-6, 1e-7, 1e-2, 0.001, 360000.0, 1e6, 1e5, 123456.7, 1234567.0, 0.5, 2.0, -14210.75823, 0.1 + 0.2, 1e-10, 12.0, 3.6e6};
    for (double x : xs) printf("F %s\n", julia_float(x).c_str());
    printf("R %s %s %s %s\n", julia_round(0.0123456, 2).c_str(), julia_round(-14210.758234567, 5).c_str(), julia_round(100 * 4.2e-7, 2).c_str(), julia_round(3.0, 2).c_str());
    emit(parameters(1e-4, 1e-4, 1e-7, 1e-7, 1e-6, 1e-7, false, true, 1000000, 360000.0));
    emit(parameters(1e-4, 1e-4, 1e-7, 1e-7, 1e-6, 1e-7, true, true, 110, 39600.0));
    emit(constraints(4, 0));
    emit(constraints(1, 20));
    emit(prob_data({3, 3, 5}, {4}));
    emit(header_2(false, false, true));
    emit(header_2(true, true, false));
    emit(progress(1000, -17.99560812, 3.2e-5, 9.1e-6, 1.234e-4, 5.6e-8, 2, 0.0312, -17.9967, -1.0, false, false, false));
    emit(progress(72, 18.0, 1e-4, 0.0, 1e-7, 1e-9, 12, 1.5, -3.25, 0.000123, true, true, false));
    emit(result("Optimal solution found", 0.0312, -17.99560812, -17.9967123, 3.2e-5, 9.1e-6, 0.0, 1));
    emit(note("Dual ray found"));
    return 0;
}
'''

EXPECTED = textwrap.dedent('''\
    F 0.0001
    F 1.0e-5
    F 1.0e-6
    F 1.0e-7
    F 0.01
    F 0.001
    F 360000.0
    F 1.0e6
    F 100000.0
    F 123456.7
    F 1.234567e6
    F 0.5
    F 2.0
    F -14210.75823
    F 0.30000000000000004
    F 1.0e-10
    F 12.0
    F 3.6e6
    R 0.01 -14210.75823 0.0 3.0
        Solver parameters:
           tol_gap = 0.0001 tol_feasibility = 0.0001
           tol_primal = 1.0e-7 tol_dual = 1.0e-7 tol_psd = 1.0e-7
           max_iter = 1000000 time_limit = 360000.0s
        Solver parameters:
           tol_gap = 0.0001 tol_feasibility = 0.0001
           tol_primal = 1.0e-7 tol_dual = 1.0e-7 tol_soc = 1.0e-6 tol_psd = 1.0e-7
           max_iter = 110 time_limit = 39600.0s
        Constraints:
           4 linear equalities and 
        Constraints:
           1 linear equality and 20 linear equalities
        Cones:
           2 second order cones of size 3
           1 second order cone of size 5
           1 psd cone of size 4
    ---------------------------------------------------------------------------------------
        Initializing Primal-Dual Hybrid Gradient method
    ---------------------------------------------------------------------------------------
    |  iter  | prim obj | rel. gap |  feasb.  | prim res | dual res | tg. rank |  time(s) |
    ---------------------------------------------------------------------------------------
    |  iter  | prim obj | rel. gap |  feasb.  | prim res | dual res | tg. rank |  time(s) | dual obj | d feasb. |
    |   1000 |-1.80e+01 | 3.20e-05 | 9.10e-06 | 1.23e-04 | 5.60e-08 |        2 |   0.0312 |
    |     72 | 1.80e+01 | 1.00e-04 | 0.00e+00 | 1.00e-07 | 1.00e-09 |       12 |      1.5 |   -3.250 |  0.00012 |
    ---------------------------------------------------------------------------------------
        Solver status:
           Optimal solution found
           Time elapsed     = 0.03 seconds
           Primal objective = -17.99561
           Dual objective   = -17.99671
           Duality gap      = 0.0 %
    ---------------------------------------------------------------------------------------
        Primal feasibility:
           ||A(X) - b|| / (1 + ||b||) = 9.0e-6    [linear equalities] 
           ||max(G(X) - h, 0)|| / (1 + ||h||) = 0.0    [linear inequalities]
        Rank of p.s.d. variable is 1.
    =======================================================================================
    ---------------------------------------------------------------------------------------
        Dual ray found
    ---------------------------------------------------------------------------------------
    ''')


def test_log_lines_follow_the_reference_format(tmp_path):
    src = tmp_path / "driver.cpp"
    src.write_text(DRIVER)
    exe = tmp_path / "driver"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "proxsdp_b200", "csrc"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.splitlines() == EXPECTED.splitlines()
