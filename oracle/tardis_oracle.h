/* tardis_oracle.h -- C interface of the CPU parity oracle.  TEST INFRASTRUCTURE ONLY
 * (see the header of tardis_oracle.c).  Layouts are the reference's own:
 * tau_sobolev / transition_probabilities / j_blue / edotlu are row-major
 * (line or transition, shell) as in tardis/opacities/opacity_state.py:183. */
#ifndef TARDIS_ORACLE_H
#define TARDIS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TARDIS_ORACLE_ERR_NU_DIFF 1      /* MonteCarloException, calculate_distances.py:106 */
#define TARDIS_ORACLE_ERR_MACRO_ATOM 2   /* MacroAtomError, macro_atom.py:95 */
#define TARDIS_ORACLE_ERR_VPACKET_LOOP 3 /* reference would loop forever (virtual_packet.py:191-243) */
#define TARDIS_ORACLE_ERR_CONTINUUM 4    /* continuum tables inconsistent (index out of range) */

typedef struct {
    int64_t n_shells, n_lines;
    const double *r_inner, *r_outer; /* [S] */
    double time_explosion;
    const double *electron_density; /* [S] */
    const double *line_list_nu;     /* [L] descending */
    const double *tau_sobolev;      /* [L,S] */
    int64_t n_transitions, n_blocks;
    const double *transition_probabilities; /* [T,S] */
    const int64_t *line2macro_level_upper;  /* [L] */
    const int64_t *macro_block_edge_index;  /* [n_blocks+1] */
    const int64_t *transition_type, *destination_level_id, *transition_line_id; /* [T] */
    /* continuum (IIP mode) tables, OpacityStateNumbaIIP, opacities/opacity_state_numba_iip.py:8-125; unused unless
     * config.continuum_processes_enabled */
    const double *t_electrons;                  /* [S] */
    int64_t n_continua, n_phot;                 /* number of bound-free continua; total cross-section points */
    const double *bf_threshold_list_nu;         /* [n_continua] */
    const double *photo_ion_nu_threshold_mins, *photo_ion_nu_threshold_maxs; /* [n_continua] */
    const int64_t *photo_ion_block_references;  /* [n_continua + 1] */
    const double *chi_bf;                       /* [n_phot, S] */
    const double *x_sect, *phot_nus;            /* [n_phot] */
    const double *ff_opacity_factor;            /* [S] */
    const double *emissivities;                 /* [n_phot, S] */
    const int64_t *photo_ion_activation_idx;    /* [n_activation] */
    int64_t n_activation, k_packet_idx;
    int64_t n_markov;                           /* absorbing_markov_probabilities is [S, n_markov, n_markov] */
    const double *absorbing_markov_probabilities;
} tardis_oracle_model;

/* MonteCarloConfiguration, transport/montecarlo/configuration/base.py:11-49 */
typedef struct {
    int enable_full_relativity;
    int line_interaction_type; /* 0 scatter, 1 downbranch, 2 macroatom */
    int disable_line_scattering;
    double sigma_thomson;      /* constants.SIGMA_THOMSON (1e-200 when e-scattering disabled) */
    int64_t number_of_vpackets;
    double survival_probability, vpacket_tau_russian;
    double vpacket_spawn_start_frequency, vpacket_spawn_end_frequency;
    const double *spectrum_frequency_grid; /* [n_grid] */
    int64_t n_grid;
    int continuum_processes_enabled; /* IIP mode (modes/iip/...): full relativity forced, no virtual packets */
} tardis_oracle_config;

typedef struct {
    int64_t n_packets;
    const double *initial_radii, *initial_nus, *initial_mus, *initial_energies;
    const int64_t *packet_seeds;
} tardis_oracle_packets;

typedef struct {
    int64_t n_line_steps, n_boundary_events, n_line_events, n_escat_events, n_rng_draws;
    int64_t n_macro_jumps, n_macro_scanned, n_vpackets, n_vpacket_line_steps;
    int64_t n_continuum_events, n_bf_estimator_updates;
} tardis_oracle_counters;

/* one TrackerFull row (packets/trackers/tracker_full.py:19-110) */
typedef struct {
    int64_t packet_id, interaction_type, status, before_shell_id, after_shell_id, line_absorb_id, line_emit_id;
    double radius, before_nu, before_mu, before_energy, after_nu, after_mu, after_energy;
} tardis_oracle_event;

typedef struct {
    double *output_nus, *output_energies; /* [N] */
    double *j, *nu_bar;                   /* [S] */
    double *j_blue, *edotlu;              /* [L,S] */
    double *vhist;                        /* [n_grid] or NULL */
    /* last-interaction tracker SoA, all [N] or all NULL */
    int64_t *last_interaction_type, *last_event_id, *last_shell_id, *last_line_absorb_id, *last_line_emit_id;
    double *last_radius, *last_before_nu, *last_before_mu, *last_before_energy;
    double *last_after_nu, *last_after_mu, *last_after_energy;
    /* full event log for the first n_tracked_packets packets, or NULL */
    tardis_oracle_event *events; /* [n_tracked_packets * max_events_per_packet] */
    int64_t *event_counts;       /* [n_tracked_packets] */
    int64_t n_tracked_packets, max_events_per_packet;
    /* virtual packet log (virtual_packet_logging), or NULL */
    double *vlog_nus, *vlog_energies, *vlog_initial_mus, *vlog_initial_rs;
    int64_t *vlog_packet_index;
    int64_t vlog_capacity, vlog_count;
    /* EstimatorsContinuum (estimators/estimators_continuum.py:15-175), [n_continua, S] / [S]; NULL unless IIP mode */
    double *photo_ion_estimator, *stim_recomb_estimator, *bf_heating_estimator, *stim_recomb_cooling_estimator;
    double *ff_heating_estimator;
    int64_t *photo_ion_estimator_statistics;
    tardis_oracle_counters counters;
} tardis_oracle_outputs;

int tardis_oracle_run(const tardis_oracle_model *m, const tardis_oracle_config *c,
                      const tardis_oracle_packets *pk, tardis_oracle_outputs *out, int nthreads);

double tardis_oracle_rng_double(uint32_t seed, int64_t skip);
double tardis_oracle_distance_boundary(double r, double mu, double r_inner, double r_outer, int64_t *delta_shell);
double tardis_oracle_distance_line(double r, double mu, double nu, double comov_nu, int is_last_line, double nu_line,
                                   double time_explosion, int full_rel, int *error);
double tardis_oracle_doppler_factor(double velocity, double mu, int full_rel);
double tardis_oracle_inverse_doppler_factor(double velocity, double mu, int full_rel);

#ifdef __cplusplus
}
#endif
#endif
