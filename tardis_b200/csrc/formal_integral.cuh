// Formal integral on the tables the source-function solve left in HBM ("the reference's Numba-CUDA formal
// integral is the only existing GPU code and could then be beaten in place").
//
// Replaces, for the path FormalIntegralSolver.solve takes after the source function
// (/root/reference/tardis/spectrum/formal_integral/formal_integral_solver.py:208-285):
//   interpolate_integrator_quantities   formal_integral_solver.py:305-430  (scipy interp1d over the shell mid-points)
//   numba_formal_integral               formal_integral_numba.py:377-567   (and its CUDA twin, formal_integral_cuda.py:272-621:
//                                                                           one thread per (frequency, impact parameter))
//   calculate_intersection_point / populate_intersection_points / line_search   formal_integral_numba.py:20-184
//   calculate_impact_parameters / intensity_black_body                          base.py:86-120
//
// Every element function is __host__ __device__ arithmetic in the reference's operation order (the library is built with
// -fmad=false); tests/formal_integral_shim.cpp compiles this header for the CPU and drives it with the kernels' loops.
//
// How the work is laid out (engine.cu holds the kernels):
//  * one table of 32-byte CELLS [shell][line]: { exp(-tau), att_S_ul, Jblue_lu, Jred_lu of the PREVIOUS line } -- everything a
//    ray needs when it passes line l in shell s is one sector.  The interpolation kernel writes it straight from the S-shell
//    tables of the source function (linear weights / nearest shell per new shell), so the 4 x [L, S2] host tables of the
//    reference (1.3 GB at L = 5e5, 79 shells) never exist.
//  * a warp takes 32 neighbouring impact parameters of ONE frequency and sweeps the line list once, all lanes at the same line:
//    the line frequency is a uniform load, the lanes of a warp sit in at most a few different shells, so one sweep step touches a
//    few sectors instead of 32 scattered ones (the reference's thread-per-ray walk reads five [L*S] arrays at 32 unrelated
//    offsets per step).  Each lane still performs exactly the reference's sequence of operations for its ray.
//  * the recurrence along a ray is sequential; nothing is re-associated.
//
// Reads behind the line list: when a ray's window reaches beyond the reddest line the reference addresses element
// shell * n_lines + n_lines of its flat tables: the first line of the next shell, or (last shell) memory behind the arrays.
// Rows here carry two extra cells that reproduce the flat addressing, with 0.0 behind the last shell.
#pragma once
#include <cmath>
#include <cstdint>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace tbfi {

constexpr double C_INV = 3.33564e-11;   // base.py:12
constexpr double KB_CGS = 1.3806488e-16;
constexpr double H_CGS = 6.62606957e-27;
constexpr double PI = 3.141592653589793;  // np.pi

struct alignas(32) Cell {
    double exp_tau;    // exp(-tau_sobolev[line, shell])                      formal_integral_numba.py:240
    double att;        // att_S_ul[line, shell]
    double jblue;      // Jblue_lu[line, shell]
    double jred_prev;  // Jred_lu flat[shell * n_lines + line - 1]: what `line_Jred_lu_idx` addresses once the first line is behind
};

// one cell = one 256-bit load (sm_100: ld.global.nc.v4.f64 -> LDG.E.ENL2.256)
__host__ __device__ inline Cell load_cell(const Cell *p) {
#ifdef __CUDA_ARCH__
    Cell c;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(c.exp_tau), "=d"(c.att), "=d"(c.jblue), "=d"(c.jred_prev) : "l"(p));
    return c;
#else
    return *p;
#endif
}

// cells per shell row: n_lines + 2 (see the header comment)
__host__ __device__ inline long long row_cells(int n_lines) { return (long long)n_lines + 2; }

// ---------------------------------------------------------------------------------------------------------------------------
// interpolation (formal_integral_solver.py:305-430; scipy.interpolate.interp1d, _call_linear / _call_nearest)
// ---------------------------------------------------------------------------------------------------------------------------
struct ShellWeights {  // one per integrator shell
    int lo, hi;        // linear: the two model shells
    double w_lo, w_hi; // (x_hi - x_new) / (x_hi - x_lo), (x_new - x_lo) / (x_hi - x_lo)
    int nearest;       // nearest model shell (half-way points go to the LEFT neighbour: interp1d kind="nearest", side="left")
};

// x: mid-points of the n model shells (ascending), x_new: mid-point of the integrator shell
inline ShellWeights shell_weights(const double *x, int n, double x_new) {
    ShellWeights w{};
    int idx = 0;  // np.searchsorted(x, x_new): number of entries < x_new
    while (idx < n && x[idx] < x_new) idx++;
    if (idx < 1) idx = 1;
    if (idx > n - 1) idx = n - 1;
    w.lo = idx - 1; w.hi = idx;
    w.w_hi = (x_new - x[w.lo]) / (x[w.hi] - x[w.lo]);
    w.w_lo = (x[w.hi] - x_new) / (x[w.hi] - x[w.lo]);
    int k = 0;  // np.searchsorted(x_bds, x_new, side="left") with x_bds = x[1:] / 2 + x[:-1] / 2
    while (k < n - 1 && (x[k + 1] / 2.0 + x[k] / 2.0) < x_new) k++;
    w.nearest = k;  // already inside [0, n - 1]
    return w;
}

// y_new = w_hi * y_hi + w_lo * y_lo, then .clip(0.0)  (NaN stays NaN, as in numpy)
__host__ __device__ inline double interp_clip(double w_lo, double w_hi, double y_lo, double y_hi) {
    const double y = w_hi * y_hi + w_lo * y_lo;
    return y < 0.0 ? 0.0 : y;
}

// The S-shell tables as they lie in HBM: shell-major rows of `lpad` doubles (tau_t of the model, att / Jred / Jblue of
// tb200_solve_source_function)
struct Tables {
    const double *tau_t, *att_t, *jred_t, *jblue_t;
    int n_lines, lpad;
};

__host__ __device__ inline double interp_table(const double *t, int lpad, const ShellWeights &w, int line) {
    return interp_clip(w.w_lo, w.w_hi, t[(size_t)w.lo * lpad + line], t[(size_t)w.hi * lpad + line]);
}

// element `flat` of the reference's Fortran-flattened [L, S2] table (index = shell * n_lines + line), 0.0 outside
__host__ __device__ inline double flat_at(const double *t, const Tables &T, const ShellWeights *w, int n_shells, long long flat) {
    if (flat < 0 || flat >= (long long)n_shells * T.n_lines) return 0.0;
    return interp_table(t, T.lpad, w[flat / T.n_lines], (int)(flat % T.n_lines));
}

// cell `line` (0 ... n_lines + 1) of integrator shell `shell`
__host__ __device__ inline Cell build_cell(const Tables &T, const ShellWeights *w, int n_shells, int shell, int line) {
    Cell c;
    const long long flat = (long long)shell * T.n_lines + line;
    if (line < T.n_lines) {
        c.exp_tau = exp(-T.tau_t[(size_t)w[shell].nearest * T.lpad + line]);
        c.att = interp_table(T.att_t, T.lpad, w[shell], line);
        c.jblue = interp_table(T.jblue_t, T.lpad, w[shell], line);
    } else {
        c.exp_tau = 1.0; c.att = 0.0;
        c.jblue = line == T.n_lines ? flat_at(T.jblue_t, T, w, n_shells, flat) : 0.0;
    }
    c.jred_prev = flat_at(T.jred_t, T, w, n_shells, flat - 1);
    return c;
}

// np.linspace(start, stop, num)[i] (num >= 2): arange * step + start, the last element set to stop
inline double linspace_at(double start, double stop, int num, int i) {
    if (i == num - 1) return stop;
    const double step = (stop - start) / (double)(num - 1);
    return (double)i * step + start;
}

// ---------------------------------------------------------------------------------------------------------------------------
// geometry of one ray
// ---------------------------------------------------------------------------------------------------------------------------
// formal_integral_numba.py:20-51
__host__ __device__ inline double intersection_point(double radius, double p, double inv_t) {
    if (radius > p) return sqrt(radius * radius - p * p) * C_INV * inv_t;
    return 0.0;
}

// base.py:104-120
__host__ __device__ inline double intensity_black_body(double frequency, double temperature) {
    if (frequency == 0) return NAN;
    const double beta_rad = 1 / (KB_CGS * temperature);
    const double coefficient = 2 * H_CGS * C_INV * C_INV;
    return coefficient * frequency * frequency * frequency / (exp(H_CGS * frequency * beta_rad) - 1);
}

// number of entries of the descending list that are > x: formal_integral_numba.py:152-184 (line_search; its two range tests
// return the same count) and :482-485 (n_lines - searchsorted(nu[::-1], x, side="right"))
__host__ __device__ inline int count_greater(const double *nu, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
        if (nu[mid] > x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct Shells {
    const double *r_inner, *r_outer;  // integrator shells
    const double *kappa;              // electron_densities * SIGMA_THOMSON per shell (formal_integral_numba.py:500-502)
    int n;
    double inv_t;                     // 1 / time_explosion
    double t_over_cinv;               // time_explosion / C_INV
};

// One ray = one (frequency, impact parameter) pair of numba_formal_integral's loops (:462-558).
struct Ray {
    double I, esc, z_start, nu_end, kappa, nu, p;
    int line_idx;   // next line of the list this ray will pass
    int seg, n_seg; // current segment, number of segments (n_intersections - 1)
    int offset;     // first shell the ray crosses; -1: the ray starts at the photosphere
    int shell;      // shell_ids[seg]
    bool first, done;

    // intersection point k and its shell (populate_intersection_points, :54-118), recomputed instead of stored
    __host__ __device__ inline double point(const Shells &g, int k, int *shell_id) const {
        if (offset < 0) { *shell_id = k; return 1 - intersection_point(g.r_outer[k], p, g.inv_t); }
        const int half = g.n - offset;
        if (k < half) { const int i = g.n - 1 - k; *shell_id = i; return 1 + intersection_point(g.r_outer[i], p, g.inv_t); }
        const int i = k - g.n + 2 * offset; *shell_id = i; return 1 - intersection_point(g.r_outer[i], p, g.inv_t);
    }

    // :262-290, :472-493.  A ray with nothing to integrate comes back `done` with its final intensity.
    __host__ __device__ inline void init(const Shells &g, const double *line_list_nu, int n_lines, double nu_, int p_idx, int n_p,
                                         double inner_temperature) {
        nu = nu_;
        p = (double)p_idx * g.r_outer[g.n - 1] / (double)(n_p - 1);  // base.py:101
        esc = 0; first = true; done = false; seg = 0; I = 0; line_idx = n_lines; shell = 0; z_start = 0; nu_end = 0; kappa = 0;
        int n_int;
        if (p <= g.r_inner[0]) { offset = -1; n_int = g.n; }
        else {
            offset = 0;
            while (offset < g.n && intersection_point(g.r_outer[offset], p, g.inv_t) == 0) offset++;
            n_int = 2 * (g.n - offset);
        }
        n_seg = n_int - 1;
        if (n_int == 0) { done = true; return; }  // p == r_max: no intersection, intensity 0
        const double z0 = point(g, 0, &shell);
        if (offset < 0) I = intensity_black_body(nu * z0, inner_temperature);
        if (n_seg <= 0) { I *= p; done = true; return; }
        const double nu_start = nu * z0;
        z_start = g.t_over_cinv * (1.0 - z0);
        line_idx = count_greater(line_list_nu, n_lines, nu_start);
        kappa = g.kappa[shell];
        int unused;
        nu_end = nu * point(g, 1, &unused);
    }

    // :538-557: the electron-scattering term up to the shell boundary, then the next segment
    __host__ __device__ inline void boundary_step(const Shells &g, const Cell *cells, long long row) {
        const Cell *c = cells + (long long)shell * row + line_idx;
        const double jred = first ? c[1].jred_prev : c[0].jred_prev;
        const double avg = 0.5 * (jred + c[0].jblue);
        const double z_end = g.t_over_cinv * (1.0 - nu_end / nu);
        esc += (z_end - z_start) * kappa * (avg - I);
        z_start = z_end;
        seg++;
        if (seg == n_seg) { I *= p; done = true; return; }
        (void)point(g, seg, &shell);
        kappa = g.kappa[shell];
        int unused;
        nu_end = nu * point(g, seg + 1, &unused);
    }

    // :503-536: one resonance point
    __host__ __device__ inline void line_step(const Shells &g, const Cell &c, double nu_line) {
        const double z_end = g.t_over_cinv * (1.0 - nu_line / nu);
        if (first) {
            esc += (z_end - z_start) * kappa * (c.jblue - I);
            first = false;
        } else {
            const double avg = 0.5 * (c.jred_prev + c.jblue);
            esc += (z_end - z_start) * kappa * (avg - I);
        }
        I += esc;
        I *= c.exp_tau;
        I += c.att;
        esc = 0;
        z_start = z_end;
        line_idx++;
    }

    // the ray at line l == line_idx < n_lines: the boundaries it crosses before that line, then the line
    __host__ __device__ inline void pass_line(const Shells &g, const Cell *cells, long long row, double nu_line) {
        while (!(nu_line > nu_end)) {  // lines of a segment: nu_line > nu_end (:482-485)
            boundary_step(g, cells, row);
            if (done) return;
        }
        line_step(g, load_cell(cells + ((long long)shell * row + line_idx)), nu_line);
    }

    // behind the last line: only boundaries are left
    __host__ __device__ inline void finish(const Shells &g, const Cell *cells, long long row) {
        while (!done) boundary_step(g, cells, row);
    }
};

// ---------------------------------------------------------------------------------------------------------------------------
// 8 pi^2 np.trapezoid(I, dx = r_max / n_p)  (:559-563).  numpy: (d * (y[1:] + y[:-1]) / 2.0).sum(); the sum here runs over
// TRAPZ_LANES interleaved partial sums combined by a halving tree (the CPU build uses the same order)
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int TRAPZ_LANES = 256;
__host__ __device__ inline double trapz_partial(int lane, const double *y, int n, double d) {
    double s = 0.0;
    for (int i = lane; i + 1 < n; i += TRAPZ_LANES) s = s + d * (y[i + 1] + y[i]) / 2.0;
    return s;
}
__host__ __device__ inline double luminosity_density(double trapz) { return 8 * PI * PI * trapz; }

}  // namespace tbfi
