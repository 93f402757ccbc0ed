// CPU build of tardis_b200/csrc/source_function.cuh (the very functions the kernels call) for tests/test_source_function.py.
#include "../tardis_b200/csrc/source_function.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

extern "C" {
// Same loops as the kernels in engine.cu, on [L,S] / [T,S] C-order host tables.  Returns the number of Jacobi sweeps (-1: not converged).
int shim_source_function(int L, int S, int T, int n_levels, int macroatom, const long long *lower, const long long *upper, const long long *ttype,
                         const long long *tline, const double *wave, const double *tau, const double *tp, const double *jblue_est,
                         const double *edotlu_est, const double *volume, double time_explosion, double time_of_simulation, double c,
                         double tol, int max_it, double *att, double *jred, double *jblue, double *e_dot_u /* [n_levels,S] */) {
    std::vector<std::vector<int>> lvl_lines(n_levels), in_rows(n_levels), in_src(n_levels);
    std::vector<int> em_row(L, -1);
    for (int l = 0; l < L; l++) lvl_lines[upper[l]].push_back(l);
    for (int t = 0; t < T; t++) {
        const int l = (int)tline[t];
        if (ttype[t] == -1) { em_row[l] = t; continue; }
        const int src = ttype[t] == 1 ? (int)lower[l] : (int)upper[l], dst = ttype[t] == 1 ? (int)upper[l] : (int)lower[l];
        in_rows[dst].push_back(t); in_src[dst].push_back(src);
    }
    int sweeps = 0;
    std::vector<double> e(n_levels), c0(n_levels), c1(n_levels), p_shell(T);
    std::vector<std::vector<double>> C(S, std::vector<double>(n_levels));
    for (int s = 0; s < S; s++) {
        const double nrm = tbsf::e_dot_lu_norm(time_of_simulation, volume[s]);
        std::vector<double> tau_s(L), est_s(L);
        for (int l = 0; l < L; l++) { tau_s[l] = tau[(size_t)l * S + s]; est_s[l] = edotlu_est[(size_t)l * S + s]; }
        for (int u = 0; u < n_levels; u++) {
            double part[tbsf::SUM_LANES];
            for (int lane = 0; lane < tbsf::SUM_LANES; lane++)
                part[lane] = tbsf::e_dot_u_partial(lane, lvl_lines[u].data(), (int)lvl_lines[u].size(), nrm, tau_s.data(), est_s.data());
            e[u] = tbsf::butterfly32(part);
        }
        C[s] = e;
    }
    if (macroatom) {
        bool converged = false;
        std::vector<std::vector<double>> E = C, P(S, std::vector<double>(T));
        for (int s = 0; s < S; s++) for (int t = 0; t < T; t++) P[s][t] = tp[(size_t)t * S + s];
        while (!converged && sweeps < max_it) {
            std::vector<double> dmax(S, 0.0), cmax(S, 0.0);
            for (int k = 0; k < 8 && sweeps < max_it; k++, sweeps++) {
                for (int s = 0; s < S; s++) {
                    dmax[s] = cmax[s] = 0.0;
                    for (int j = 0; j < n_levels; j++) {
                        double part[tbsf::SUM_LANES];
                        for (int lane = 0; lane < tbsf::SUM_LANES; lane++)
                            part[lane] = tbsf::jacobi_partial(lane, in_rows[j].data(), in_src[j].data(), (int)in_rows[j].size(), P[s].data(), C[s].data());
                        const double v = E[s][j] + tbsf::butterfly32(part);
                        c1[j] = v;
                        dmax[s] = std::max(dmax[s], std::fabs(v - C[s][j])); cmax[s] = std::max(cmax[s], std::fabs(v));
                    }
                    std::copy(c1.begin(), c1.end(), C[s].begin());
                }
            }
            converged = true;
            for (int s = 0; s < S; s++) if (!(dmax[s] <= tol * cmax[s])) converged = false;
        }
        if (!converged) return -1;
    }
    for (int s = 0; s < S; s++) {
        const double jn = tbsf::j_blue_lu_norm(c, time_explosion, time_of_simulation, volume[s]);
        for (int l = 0; l < L; l++) {
            const size_t i = (size_t)l * S + s;
            att[i] = tbsf::att_s_ul(wave[l], tp[(size_t)em_row[l] * S + s], C[s][upper[l]], time_explosion);
            jblue[i] = jblue_est[i] * jn;
            jred[i] = tbsf::j_red_lu(jblue[i], tau[i], att[i]);
        }
        for (int u = 0; u < n_levels; u++) e_dot_u[(size_t)u * S + s] = C[s][u];
    }
    return sweeps;
}
}
