// printing.hpp — the log `chambolle_pock` writes when opt.log_verbose is set (reference src/printing.jl:1-169 and the
// println calls of src/pdhg.jl:44-52, 176-178, 195-259, 486-505, 642-660).  Host-only C++ (no CUDA): the lines are built
// as strings so that tests/test_printing.py can compile this header alone and compare them with the reference's format.
//
// Julia prints a Float64 through `show`: the shortest digit string that round-trips, in fixed notation when the decimal
// exponent lies in [-4, 5] and as d.ddde±x otherwise, always with a digit after the point (1.0e-5, 0.0001, 360000.0).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace pb {
namespace plog {

inline std::string julia_float(double v) {
    if (v != v) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Inf" : "-Inf";
    if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[64];
    int prec = 0;
    for (prec = 0; prec < 17; ++prec) {
        snprintf(buf, sizeof(buf), "%.*e", prec, v);
        if (strtod(buf, nullptr) == v) break;
    }
    // buf = [-]d[.ddd]e[+-]xx
    std::string s(buf);
    const size_t epos = s.find('e');
    const int e10 = atoi(s.c_str() + epos + 1);
    std::string mant = s.substr(0, epos);
    std::string sign;
    if (mant[0] == '-') { sign = "-"; mant = mant.substr(1); }
    std::string digits;
    for (char c : mant) if (c != '.') digits += c;
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    std::string out;
    if (e10 >= -4 && e10 <= 5) {
        if (e10 < 0) {
            out = "0." + std::string((size_t)(-e10 - 1), '0') + digits;
        } else {
            if ((int)digits.size() <= e10 + 1) out = digits + std::string((size_t)(e10 + 1 - (int)digits.size()), '0') + ".0";
            else out = digits.substr(0, (size_t)e10 + 1) + "." + digits.substr((size_t)e10 + 1);
        }
    } else {
        out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : std::string("0")) + "e" + std::to_string(e10);
    }
    return sign + out;
}

// round(x; digits = d) as Julia prints it afterwards
inline std::string julia_round(double x, int d) {
    if (x != x || std::isinf(x)) return julia_float(x);
    const double s = std::pow(10.0, d);
    double r = std::nearbyint(x * s) / s;
    if (r == 0.0) r = std::signbit(x) ? -0.0 : 0.0;
    return julia_float(r);
}

inline std::string fmt(const char* f, double v) { char b[64]; snprintf(b, sizeof(b), f, v); return b; }
inline std::string pad(const std::string& s, size_t w) { return (s.size() < w ? std::string(w - s.size(), ' ') : std::string()) + s; }

const char* const BAR87 = "---------------------------------------------------------------------------------------";
const char* const EQ87 = "=======================================================================================";

// print_header_1 (printing.jl:1-9)
inline std::string header_1() {
    std::string s;
    s += std::string(BAR87) + "\n" + EQ87 + "\n";
    s += "                  ProxSDP : Proximal Semidefinite Programming Solver                   \n";
    s += "                         (c) Mario Souto and Joaquim D. Garcia, 2020                   \n";
    s += "                                                              v1.8.4                   \n";
    s += std::string(BAR87) + "\n";
    return s;
}

// print_parameters (printing.jl:11-26)
inline std::string parameters(double tol_gap, double tol_feasibility, double tol_primal, double tol_dual, double tol_soc,
                              double tol_psd, bool has_soc, bool has_psd, long long max_iter_local, double time_limit) {
    std::string s = "    Solver parameters:\n";
    s += "       tol_gap = " + julia_float(tol_gap) + " tol_feasibility = " + julia_float(tol_feasibility) + "\n";
    s += "       tol_primal = " + julia_float(tol_primal) + " tol_dual = " + julia_float(tol_dual);
    if (has_soc) s += " tol_soc = " + julia_float(tol_soc);
    if (has_psd) s += " tol_psd = " + julia_float(tol_psd);
    s += "\n";
    s += "       max_iter = " + std::to_string(max_iter_local) + " time_limit = " + julia_float(time_limit) + "s\n";
    return s;
}

// print_constraints (printing.jl:28-35; the reference words the inequalities with `eqs` as well)
inline std::string eqs(long long v) { return v > 0 ? std::to_string(v) + " linear equalit" + (v != 1 ? "ies" : "y") : std::string(); }
inline std::string constraints(long long p, long long m) {
    return "    Constraints:\n       " + eqs(p) + " and " + eqs(m) + "\n";
}

// print_prob_data (printing.jl:37-69).  The reference walks a Dict (hash order); here the sizes come in ascending order.
inline std::string prob_data(const std::vector<long long>& soc_lens, const std::vector<long long>& psd_sides) {
    std::map<long long, long long> soc, psd;
    for (long long l : soc_lens) soc[l]++;
    for (long long l : psd_sides) psd[l]++;
    std::string s = "    Cones:\n";
    for (auto& kv : soc) s += "       " + std::to_string(kv.second) + " second order cone" + (kv.second != 1 ? "s" : "") + " of size " + std::to_string(kv.first) + "\n";
    for (auto& kv : psd) s += "       " + std::to_string(kv.second) + " psd cone" + (kv.second != 1 ? "s" : "") + " of size " + std::to_string(kv.first) + "\n";
    return s;
}

// print_header_2 (printing.jl:71-97)
inline std::string header_2(bool extended_log, bool extended_log2, bool beg) {
    std::string bar = BAR87;
    std::string cols = "|  iter  | prim obj | rel. gap |  feasb.  | prim res | dual res | tg. rank |  time(s) |";
    if (extended_log || extended_log2) { bar += "-----------"; cols += " dual obj |"; }
    if (extended_log2) { bar += "-----------"; cols += " d feasb. |"; }
    std::string s;
    if (beg) s += bar + "\n" + "    Initializing Primal-Dual Hybrid Gradient method\n" + bar + "\n";
    s += cols + "\n";
    if (beg) s += bar + "\n";
    return s;
}

// print_progress (printing.jl:99-151)
inline std::string progress(long long iter, double prim_obj, double gap, double feas, double primal_res, double dual_res,
                            long long sum_target_rank, double elapsed, double dual_obj, double dual_feas_val,
                            bool extended_log, bool extended_log2, bool repeat_header) {
    std::string a = "|";
    a += pad(std::to_string(iter) + " |", 9);
    a += pad(fmt("%.2e", prim_obj) + " |", 11);
    a += pad(fmt("%.2e", gap) + " |", 11);
    a += pad(fmt("%.2e", feas) + " |", 11);
    a += pad(fmt("%.2e", primal_res) + " |", 11);
    a += pad(fmt("%.2e", dual_res) + " |", 11);
    a += pad(fmt("%g", (double)sum_target_rank) + " |", 11);
    a += pad(fmt("%g", elapsed) + " |", 11);
    if (extended_log || extended_log2) a += pad(fmt("%.3f", dual_obj) + " |", 11);
    if (extended_log2) a += pad(fmt("%.5f", dual_feas_val) + " |", 11);
    std::string s;
    if (repeat_header) s += header_2(extended_log, extended_log2, false);
    return s + a + "\n";
}

// print_result (printing.jl:153-169)
inline std::string result(const std::string& stop_reason_string, double time_, double prim_obj, double dual_obj, double gap,
                          double equa_feasibility, double ineq_feasibility, long long max_rank) {
    std::string s = std::string(BAR87) + "\n";
    s += "    Solver status:\n";
    s += "       " + stop_reason_string + "\n";
    s += "       Time elapsed     = " + julia_round(time_, 2) + " seconds\n";
    s += "       Primal objective = " + julia_round(prim_obj, 5) + "\n";
    s += "       Dual objective   = " + julia_round(dual_obj, 5) + "\n";
    s += "       Duality gap      = " + julia_round(100 * gap, 2) + " %\n";
    s += std::string(BAR87) + "\n";
    s += "    Primal feasibility:\n";
    s += "       ||A(X) - b|| / (1 + ||b||) = " + julia_round(equa_feasibility, 6) + "    [linear equalities] \n";
    s += "       ||max(G(X) - h, 0)|| / (1 + ||h||) = " + julia_round(ineq_feasibility, 6) + "    [linear inequalities]\n";
    s += "    Rank of p.s.d. variable is " + std::to_string(max_rank) + ".\n";
    s += std::string(EQ87) + "\n";
    return s;
}

// the three-line notes of the certificate search (pdhg.jl:195-259, 642-660)
inline std::string note(const char* text) { return std::string(BAR87) + "\n    " + text + "\n" + BAR87 + "\n"; }

inline void emit(const std::string& s) { fputs(s.c_str(), stdout); fflush(stdout); }

}  // namespace plog
}  // namespace pb
