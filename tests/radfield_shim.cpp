// CPU build of tardis_b200/csrc/radfield.cuh for the unit tests (tests/test_radfield.py): the SAME functions the CUDA
// kernels run, driven the way tb200_solve_radiation_field drives them.  Test harness only.
#include "../tardis_b200/csrc/radfield.cuh"

extern "C" void shim_radfield(int n_shells, int n_lines, const double *j, const double *nu_bar, const double *j_blue /* [L,S] */,
                              const double *volume, const double *nu, double time_explosion, double time_of_simulation, double w_epsilon,
                              int window, const double *constants /* t_const, sigma_sb, c, h, k_b */, double *t_rad, double *w,
                              double *j_blues /* [L,S] */) {
    tbr::Constants K{constants[0], constants[1], constants[2], constants[3], constants[4]};
    for (int s = 0; s < n_shells; s++) {
        tbr::dilute_planck(K, j[s], nu_bar[s], time_of_simulation, volume[s], &t_rad[s], &w[s]);
        const double norm = K.c * time_explosion / (4 * M_PI * time_of_simulation * volume[s]);
        for (int l = 0; l < n_lines; l++)
            j_blues[(size_t)l * n_shells + s] = tbr::j_blue_cell(K, j_blue[(size_t)l * n_shells + s], norm, nu[l], t_rad[s], w[s], w_epsilon, window != 0);
    }
}
