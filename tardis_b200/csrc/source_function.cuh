// Formal-integral source function on the estimators the transport kernels left in HBM (second half).
//
// Replaces SourceFunctionSolver.solve, /root/reference/tardis/spectrum/formal_integral/source_function.py:27-358, which the
// reference runs on the host between the last Monte Carlo iteration and the formal integral: a pandas group-by over the
// [L,S] line estimator, one sparse direct solve (scipy spsolve) of an n_levels x n_levels system PER SHELL, and three [L,S]
// table expressions.  Here every element function is plain __host__ __device__ arithmetic in the reference's operation order
// (this header is also compiled for the CPU by tests/source_function_shim.cpp); the kernels that call them are in engine.cu.
//
//   e_dot_lu[l,s]  = (1 / (t_sim V_s)) (1 - exp(-tau)) Edotlu                                   :189-192
//   e_dot_u[u,s]   = sum of e_dot_lu over the lines whose upper level is u (fixed order, see below)  :194-198
//   macroatom:       (I - Q_s)^T C_s = e_dot_u[:,s],  Q_s[src,dst] = sum of the internal rows     :200-221
//                    solved as the fixed point C <- e + Q^T C (Q >= 0, row sums < 1: the series of an absorbing chain,
//                    i.e. the same unique solution the reference's LU finds), iterated to a relative change < tolerance
//   att_S_ul[l,s]  = ((lambda_l (q_ul e_dot_u[upper(l)])) t_exp) / (4 pi)                          :276-291
//   Jblue_lu[l,s]  = J_blue (c t_exp / (4 pi t_sim V_s))                                         :320-334
//   Jred_lu[l,s]   = Jblue_lu exp(-tau) + att_S_ul                                               :358
#pragma once
#include <cmath>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace tbsf {

constexpr double FOUR_PI = 12.566370614359172;  // 4 * np.pi

// 1 / (time_of_simulation * volume), :189
__host__ __device__ inline double e_dot_lu_norm(double time_of_simulation, double volume) { return 1.0 / (time_of_simulation * volume); }

// e_dot_lu_norm_factor * exptau * e_dot_lu_estimator, :190-192 (left to right)
__host__ __device__ inline double e_dot_lu(double norm, double tau, double e_dot_lu_estimator) {
    const double exptau = 1.0 - exp(-tau);
    return norm * exptau * e_dot_lu_estimator;
}

// wave * (q_ul * e_dot_u) * time_explosion / (4 * np.pi), :280
__host__ __device__ inline double att_s_ul(double wavelength_cm, double q_ul, double e_dot_u, double time_explosion) {
    return wavelength_cm * (q_ul * e_dot_u) * time_explosion / FOUR_PI;
}

// const.c.cgs * time_explosion / (4 * np.pi * time_of_simulation * volume), :320-328
__host__ __device__ inline double j_blue_lu_norm(double c, double time_explosion, double time_of_simulation, double volume) {
    return c * time_explosion / (FOUR_PI * time_of_simulation * volume);
}

// Jblue_lu * np.exp(-tau_sobolevs) + att_S_ul, :358
__host__ __device__ inline double j_red_lu(double j_blue_lu, double tau, double att) { return j_blue_lu * exp(-tau) + att; }

// Sums over a level's row list are formed by 32 interleaved partial sums (entry k goes to partial k mod 32, ascending k) that are
// then combined by an xor butterfly (offsets 16, 8, 4, 2, 1).  On the device a warp owns one (level, shell): a lane per partial,
// __shfl_xor for the butterfly -- a level of the bench model ends 18 000 internal rows, one thread per level took 5 ms per
// sweep.  `butterfly32` is the same combination on an array, for the host build of this header.
constexpr int SUM_LANES = 32;
__host__ __device__ inline double butterfly32(double *v /* [32], destroyed */) {
    for (int o = SUM_LANES / 2; o >= 1; o >>= 1) {
        double t[SUM_LANES];
        for (int i = 0; i < SUM_LANES; i++) t[i] = v[i] + v[i ^ o];
        for (int i = 0; i < SUM_LANES; i++) v[i] = t[i];
    }
    return v[0];
}
// partial `lane` of e_dot_u of one level: its lines lane, lane + 32, ... in ascending order
__host__ __device__ inline double e_dot_u_partial(int lane, const int *lines, int n, double norm, const double *tau_shell, const double *est_shell) {
    double acc = 0.0;
    for (int k = lane; k < n; k += SUM_LANES) acc += e_dot_lu(norm, tau_shell[lines[k]], est_shell[lines[k]]);
    return acc;
}
// partial `lane` of (Q^T C_old)_j: the internal rows that END in level j (CSR by destination, ascending row order)
__host__ __device__ inline double jacobi_partial(int lane, const int *in_rows, const int *in_src, int n_in, const double *p_shell,
                                                 const double *c_old_shell) {
    double acc = 0.0;
    for (int k = lane; k < n_in; k += SUM_LANES) acc += p_shell[in_rows[k]] * c_old_shell[in_src[k]];
    return acc;
}

}  // namespace tbsf
