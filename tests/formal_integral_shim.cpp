// CPU build of the PRODUCT's formal-integral header (tardis_b200/csrc/formal_integral.cuh) with the kernels' loops, for
// tests/test_formal_integral.py.  g++ -O2 -ffp-contract=off: the same plain IEEE arithmetic the library gets from -fmad=false.
#include <cstdint>
#include <vector>

#include "../tardis_b200/csrc/formal_integral.cuh"

extern "C" {

// interpolation set-up of one integrator shell (what tb200_formal_integral computes on the host)
void fi_shim_shell_weights(const double *x, int n, double x_new, int *lo, int *hi, double *w_lo, double *w_hi, int *nearest) {
    const tbfi::ShellWeights w = tbfi::shell_weights(x, n, x_new);
    *lo = w.lo; *hi = w.hi; *w_lo = w.w_lo; *w_hi = w.w_hi; *nearest = w.nearest;
}
double fi_shim_linspace_at(double start, double stop, int num, int i) { return tbfi::linspace_at(start, stop, num, i); }
double fi_shim_intersection_point(double radius, double p, double inv_t) { return tbfi::intersection_point(radius, p, inv_t); }
double fi_shim_black_body(double frequency, double temperature) { return tbfi::intensity_black_body(frequency, temperature); }
int fi_shim_count_greater(const double *nu, int n, double x) { return tbfi::count_greater(nu, n, x); }

// intersection points / shell ids of one ray as Ray::point yields them; returns n_intersections
int fi_shim_ray_points(int n_shells, const double *r_inner, const double *r_outer, double time_explosion, int p_idx, int n_p, double *points,
                       int *shell_ids) {
    std::vector<double> kappa((size_t)n_shells, 0.0);
    tbfi::Shells g{r_inner, r_outer, kappa.data(), n_shells, 1 / time_explosion, time_explosion / tbfi::C_INV};
    const double nu_line[1] = {1.0};
    tbfi::Ray r;
    r.init(g, nu_line, 1, 1.0, p_idx, n_p, 1e4);
    const int n_int = r.n_seg + 1;
    for (int k = 0; k < n_int; k++) points[k] = r.point(g, k, &shell_ids[k]);
    return n_int;
}

// the whole pipeline on S-shell tables in the device layout ([S][lpad]): cells -> rays -> trapezoid
//   weights: per integrator shell lo, hi, nearest (int) and w_lo, w_hi (double), as fi_shim_shell_weights returned them
int fi_shim_formal_integral(int n_model_shells, int n_shells, const double *r_inner_i, const double *r_outer_i, double time_explosion, int n_lines,
                            int lpad, const double *line_list_nu, const double *tau_t, const double *att_t, const double *jred_t,
                            const double *jblue_t, const int *w_lo_idx, const int *w_hi_idx, const int *w_nearest, const double *w_lo,
                            const double *w_hi, const double *electron_densities /* [n_model_shells] */, double sigma_thomson,
                            double inner_temperature, int n_frequencies, const double *frequencies, int n_p, double *intensities_nu_p,
                            double *luminosity_densities, double *cells_out /* [n_shells][n_lines + 2][4] or NULL */, int warp_sweep) {
    (void)n_model_shells;
    std::vector<tbfi::ShellWeights> w((size_t)n_shells);
    std::vector<double> kappa((size_t)n_shells);
    for (int s = 0; s < n_shells; s++) {
        w[(size_t)s].lo = w_lo_idx[s]; w[(size_t)s].hi = w_hi_idx[s]; w[(size_t)s].nearest = w_nearest[s];
        w[(size_t)s].w_lo = w_lo[s]; w[(size_t)s].w_hi = w_hi[s];
        kappa[(size_t)s] = electron_densities[w_nearest[s]] * sigma_thomson;
    }
    const tbfi::Tables T{tau_t, att_t, jred_t, jblue_t, n_lines, lpad};
    const long long row = tbfi::row_cells(n_lines);
    std::vector<tbfi::Cell> cells((size_t)(row * n_shells));
    for (int s = 0; s < n_shells; s++)
        for (int l = 0; l < (int)row; l++) cells[(size_t)(s * row + l)] = tbfi::build_cell(T, w.data(), n_shells, s, l);
    if (cells_out)
        for (size_t k = 0; k < cells.size(); k++) {
            cells_out[4 * k] = cells[k].exp_tau; cells_out[4 * k + 1] = cells[k].att; cells_out[4 * k + 2] = cells[k].jblue; cells_out[4 * k + 3] = cells[k].jred_prev;
        }
    tbfi::Shells g{r_inner_i, r_outer_i, kappa.data(), n_shells, 1 / time_explosion, time_explosion / tbfi::C_INV};
    for (int f = 0; f < n_frequencies; f++) {
        double *I = intensities_nu_p + (size_t)f * n_p;
        I[0] = 0.0;
        if (!warp_sweep) {
            for (int p = 1; p < n_p; p++) {  // the sweep of one lane: lines in ascending index, then what is left behind the list
                tbfi::Ray r;
                r.init(g, line_list_nu, n_lines, frequencies[f], p, n_p, inner_temperature);
                while (!r.done && r.line_idx < n_lines) r.pass_line(g, cells.data(), row, line_list_nu[r.line_idx]);
                if (!r.done) r.finish(g, cells.data(), row);
                I[p] = r.I;
            }
        } else {
            // fi_rays_kernel's loop, statement by statement, with the 32 lanes of a warp run one after the other
            const int n_blocks = (n_p - 1 + 31) / 32;
            for (int b = 0; b < n_blocks; b++) {
                tbfi::Ray r[32];
                for (int lane = 0; lane < 32; lane++) {
                    const int p_idx = 1 + b * 32 + lane;
                    if (p_idx < n_p) r[lane].init(g, line_list_nu, n_lines, frequencies[f], p_idx, n_p, inner_temperature);
                    else { r[lane].done = true; r[lane].I = 0.0; r[lane].line_idx = n_lines; }
                }
                for (;;) {
                    int l = 0x7fffffff;  // __reduce_min_sync
                    for (int lane = 0; lane < 32; lane++) { const int v = r[lane].done ? 0x7fffffff : r[lane].line_idx; if (v < l) l = v; }
                    if (l == 0x7fffffff) break;
                    if (l >= n_lines) {
                        for (int lane = 0; lane < 32; lane++) if (!r[lane].done) r[lane].finish(g, cells.data(), row);
                        break;
                    }
                    const double nl = line_list_nu[l];
                    for (int lane = 0; lane < 32; lane++)
                        if (!r[lane].done && r[lane].line_idx == l) r[lane].pass_line(g, cells.data(), row, nl);
                }
                for (int lane = 0; lane < 32; lane++) {
                    const int p_idx = 1 + b * 32 + lane;
                    if (p_idx < n_p) I[p_idx] = r[lane].I;
                }
            }
        }
        double part[tbfi::TRAPZ_LANES];
        const double d = r_outer_i[n_shells - 1] / (double)n_p;
        for (int lane = 0; lane < tbfi::TRAPZ_LANES; lane++) part[lane] = tbfi::trapz_partial(lane, I, n_p, d);
        for (int o = tbfi::TRAPZ_LANES / 2; o >= 1; o >>= 1)
            for (int lane = 0; lane < o; lane++) part[lane] = part[lane] + part[lane + o];
        luminosity_densities[f] = tbfi::luminosity_density(part[0]);
    }
    return 0;
}
}
