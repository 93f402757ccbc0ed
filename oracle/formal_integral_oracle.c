/* CPU parity oracle of the formal integral -- TEST INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline leg).
 *
 * A statement-by-statement restatement, in plain IEEE double arithmetic, of
 *   /root/reference/tardis/spectrum/formal_integral/formal_integral_numba.py
 *     calculate_intersection_point      :20-51
 *     populate_intersection_points      :54-118
 *     reverse_binary_search / line_search :121-184
 *     initialize_formal_integral_inputs :194-301   (impact parameters, intersection points, black-body start intensity)
 *     get_electron_scattering_optical_depth :304-374
 *     numba_formal_integral             :377-567   (the loops; the trapezoid of :559-563 is left to numpy in the Python wrapper,
 *                                                   which calls np.trapezoid exactly as the reference does)
 *   /root/reference/tardis/spectrum/formal_integral/base.py
 *     calculate_impact_parameters       :86-101
 *     intensity_black_body              :104-120
 * Pinned against goldens produced by the unmodified reference functions (tests/golden/formal_integral_*.npz,
 * oracle/reference_runner.py::run_reference_formal_integral).  The product never links this file.
 *
 * Out-of-range reads of the reference: when a ray's frequency window reaches beyond the reddest line, `line_idx` becomes n_lines
 * and `mean_intensity_blue_lu[line_idx_offset]` / `mean_intensity_red_lu[...]` address the first line of the next shell -- or,
 * in the last shell, memory behind the arrays (Numba does not check bounds).  The tables given here carry ONE extra element
 * behind the last shell (the wrapper appends 0.0), which makes that case deterministic; everywhere else the flat-array
 * semantics of the reference are kept. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define C_INV 3.33564e-11   /* base.py:12 */
#define KB_CGS 1.3806488e-16
#define H_CGS 6.62606957e-27

static double intersection_point(double radius, double p, double inv_t) { /* :20-51 */
    if (radius > p) return sqrt(radius * radius - p * p) * C_INV * inv_t;
    return 0.0;
}

/* :54-118 */
static int64_t populate(const double *r_inner, const double *r_outer, int64_t N, double time_explosion, double p, double *points, int64_t *shell_ids) {
    const double inv_t = 1 / time_explosion;
    int64_t offset = N;
    if (p <= r_inner[0]) {
        for (int64_t i = 0; i < N; i++) {
            points[i] = 1 - intersection_point(r_outer[i], p, inv_t);
            shell_ids[i] = i;
        }
        return N;
    }
    for (int64_t i = 0; i < N; i++) {
        const double x = intersection_point(r_outer[i], p, inv_t);
        if (x == 0) continue;
        if (offset == N) offset = i;
        const int64_t i_low = N - i - 1, i_up = N + i - 2 * offset;
        points[i_low] = 1 + x; shell_ids[i_low] = i;
        points[i_up] = 1 - x; shell_ids[i_up] = i;
    }
    return 2 * (N - offset);
}

/* number of entries of the descending array nu that are > x  ==  n - searchsorted(nu[::-1], x, side="right") */
static int64_t count_greater(const double *nu, int64_t n, double x) {
    int64_t lo = 0, hi = n;  /* first index with nu[i] <= x */
    while (lo < hi) {
        const int64_t mid = (lo + hi) / 2;
        if (nu[mid] > x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* :152-184 (reverse_binary_search + 1 == the same count inside the list's range) */
static int64_t line_search(const double *nu, double nu_insert, int64_t n) {
    if (nu_insert > nu[0]) return 0;
    if (nu_insert < nu[n - 1]) return n;
    return count_greater(nu, n, nu_insert);
}

static double intensity_black_body(double frequency, double temperature) { /* base.py:104-120 */
    if (frequency == 0) return NAN;
    const double beta_rad = 1 / (KB_CGS * temperature);
    const double coefficient = 2 * H_CGS * C_INV * C_INV;
    return coefficient * frequency * frequency * frequency / (exp(H_CGS * frequency * beta_rad) - 1);
}

/* intensities_nu_p [n_frequencies][n_p] (row 0 of every frequency stays 0; every entry already multiplied by its impact parameter,
 * :558).  att_S_ul / Jred_lu / Jblue_lu: flat [n_shells * n_lines + 1] (shell-major: index = shell * n_lines + line), exp_tau likewise
 * [n_shells * n_lines].  Returns 0, or -1 when an allocation failed. */
int tb_oracle_formal_integral(int64_t n_shells, const double *r_inner, const double *r_outer, double time_explosion, int64_t n_lines,
                              const double *line_list_nu, double inner_temperature, int64_t n_frequencies, const double *frequencies,
                              const double *att_S_ul, const double *Jred_lu, const double *Jblue_lu, const double *exp_tau,
                              const double *electron_densities, double sigma_thomson, int64_t n_p, double *intensities_nu_p) {
    const int64_t N = n_shells;
    double *points = (double *)malloc((size_t)(2 * N) * sizeof(double));
    int64_t *shell_ids = (int64_t *)malloc((size_t)(2 * N) * sizeof(int64_t));
    if (!points || !shell_ids) { free(points); free(shell_ids); return -1; }
    const double radius_max = r_outer[N - 1], radius_photosphere = r_inner[0];
    for (int64_t f = 0; f < n_frequencies; f++) {
        const double nu = frequencies[f];
        double *I_nu = intensities_nu_p + f * n_p;
        I_nu[0] = 0.0;
        for (int64_t pi = 1; pi < n_p; pi++) {
            const double p = (double)pi * radius_max / (double)(n_p - 1);  /* base.py:101 */
            for (int64_t k = 0; k < 2 * N; k++) { points[k] = 0.0; shell_ids[k] = 0; }
            const int64_t n_int = populate(r_inner, r_outer, N, time_explosion, p, points, shell_ids);
            double I = (p <= radius_photosphere) ? intensity_black_body(nu * points[0], inner_temperature) : 0.0;  /* :283-290 */
            /* :476-493 */
            const double nu_start = nu * points[0];
            double z_start = time_explosion / C_INV * (1.0 - points[0]);
            const int64_t idx_nu_start = line_search(line_list_nu, nu_start, n_lines);
            int64_t line_idx = idx_nu_start;
            int64_t line_idx_offset = idx_nu_start + shell_ids[0] * n_lines;
            int64_t jred_idx = line_idx_offset;
            int first = 1;
            double esc = 0;
            for (int64_t i = 0; i < n_int - 1; i++) {
                const double escat_opacity = electron_densities[shell_ids[i]] * sigma_thomson;
                const double nu_end = nu * points[i + 1];
                const int64_t nu_end_idx = count_greater(line_list_nu, n_lines, nu_end);  /* n_lines - searchsorted(nu[::-1], nu_end, "right") */
                const int64_t steps = nu_end_idx - line_idx > 0 ? nu_end_idx - line_idx : 0;
                for (int64_t s = 0; s < steps; s++) {
                    const double z_end = time_explosion / C_INV * (1.0 - line_list_nu[line_idx] / nu);
                    if (first == 1) {  /* :348-358 */
                        esc += (z_end - z_start) * escat_opacity * (Jblue_lu[line_idx_offset] - I);
                        first = 0;
                    } else {           /* :359-371 */
                        const double avg = 0.5 * (Jred_lu[jred_idx] + Jblue_lu[line_idx_offset]);
                        esc += (z_end - z_start) * escat_opacity * (avg - I);
                        jred_idx += 1;
                    }
                    I += esc;                        /* :522 */
                    I *= exp_tau[line_idx_offset];   /* :524-526 */
                    I += att_S_ul[line_idx_offset];  /* :527-529 */
                    esc = 0;
                    z_start = z_end;
                    line_idx += 1;
                    line_idx_offset += 1;
                }
                /* :538-552 */
                const double avg = 0.5 * (Jred_lu[jred_idx] + Jblue_lu[line_idx_offset]);
                const double z_end = time_explosion / C_INV * (1.0 - nu_end / nu);
                esc += (z_end - z_start) * escat_opacity * (avg - I);
                z_start = z_end;
                const int64_t direction = (shell_ids[i + 1] - shell_ids[i]) * n_lines;
                line_idx_offset += direction;
                jred_idx += direction;
            }
            I *= p;  /* :558 */
            I_nu[pi] = I;
        }
    }
    free(points); free(shell_ids);
    return 0;
}

/* the helpers alone, for the known-answer tests ported from the reference's test_numba_formal_integral.py */
double tb_oracle_fi_intersection_point(double radius, double p, double inv_t) { return intersection_point(radius, p, inv_t); }
int64_t tb_oracle_fi_populate(const double *r_inner, const double *r_outer, int64_t N, double time_explosion, double p, double *points, int64_t *shell_ids) {
    return populate(r_inner, r_outer, N, time_explosion, p, points, shell_ids);
}
int64_t tb_oracle_fi_line_search(const double *nu, double nu_insert, int64_t n) { return line_search(nu, nu_insert, n); }
double tb_oracle_fi_black_body(double frequency, double temperature) { return intensity_black_body(frequency, temperature); }
