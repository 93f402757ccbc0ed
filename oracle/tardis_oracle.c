/*
 * tardis_oracle.c -- CPU restatement of TARDIS's Monte Carlo packet-propagation path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * engine in tardis_b200/csrc.  It may be compiled, linked or executed only by
 * tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs
 * of bench.py.  Nothing under tardis_b200/ may call it and it is never a
 * runtime fallback.
 *
 * It follows the reference's algorithm statement by statement (plain IEEE
 * double arithmetic, no fast-math, no FMA contraction: build with
 * -O2 -ffp-contract=off).  Every function cites the reference file:line
 * (paths relative to /root/reference/tardis/) it restates.  The reference
 * itself is Numba with fastmath=True, so it is NOT IEEE-reproducible; this
 * oracle is pinned against it through tests/golden (generated from the real
 * reference by tests/golden/make_golden.py): integer trajectories identical,
 * floats to ~1e-12.
 *
 * Third-party arithmetic on the path: Numba's MT19937 (numba 0.65,
 * numba/_random.c:37-75 init/shuffle, numba/cpython/randomimpl.py:109-147
 * tempering and double generation), seeded per packet by
 * np.random.seed(seed) (transport/montecarlo/modes/montecarlo_transport.py:65).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

#include "tardis_oracle.h"

/* transport/montecarlo/configuration/constants.py:3-8 */
#define C_SPEED_OF_LIGHT 2.99792458e10
#define CLOSE_LINE_THRESHOLD 1e-14
#define MISS_DISTANCE 1e99

/* CODATA-2010 cgs (tardis/constants.py:1); KB, H of configuration/constants.py:7-8 and interaction_events.py:15-16 */
#define K_BOLTZMANN 1.3806488e-16
#define H_PLANCK 6.62606957e-27
#define M_ELECTRON 9.10938291e-28
#define E_ESU 4.80320425e-10

/* packets/radiative_packet.py:12-43 */
enum { IT_BOUNDARY = 1, IT_LINE = 2, IT_ESCATTERING = 4, IT_CONTINUUM_PROCESS = 8 };
enum { ST_IN_PROCESS = 0, ST_EMITTED = 1, ST_REABSORBED = 2, ST_ADIABATIC_COOLING = 4 };

/* ------------------------------------------------------------------ RNG */
/* numba/_random.c:37-75 */
#define MT_N 624
#define MT_M 397
typedef struct {
    uint32_t mt[MT_N];
    int index;
    int64_t draws; /* number of doubles drawn (counter, not in the reference) */
} mt_state;

static void mt_init(mt_state *s, uint32_t seed)
{
    for (int pos = 0; pos < MT_N; pos++) {
        s->mt[pos] = seed;
        seed = 1812433253U * (seed ^ (seed >> 30)) + (uint32_t)pos + 1U;
    }
    s->index = MT_N;
    s->draws = 0;
}

static void mt_shuffle(mt_state *s)
{
    int i;
    uint32_t y;
    uint32_t *mt = s->mt;
    for (i = 0; i < MT_N - MT_M; i++) {
        y = (mt[i] & 0x80000000U) | (mt[i + 1] & 0x7fffffffU);
        mt[i] = mt[i + MT_M] ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 1)) & 0x9908b0dfU);
    }
    for (; i < MT_N - 1; i++) {
        y = (mt[i] & 0x80000000U) | (mt[i + 1] & 0x7fffffffU);
        mt[i] = mt[i + (MT_M - MT_N)] ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 1)) & 0x9908b0dfU);
    }
    y = (mt[MT_N - 1] & 0x80000000U) | (mt[0] & 0x7fffffffU);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 1)) & 0x9908b0dfU);
}

/* numba/cpython/randomimpl.py:109-131 */
static uint32_t mt_next_u32(mt_state *s)
{
    if (s->index >= MT_N) {
        mt_shuffle(s);
        s->index = 0;
    }
    uint32_t y = s->mt[s->index++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680U;
    y ^= (y << 15) & 0xefc60000U;
    y ^= y >> 18;
    return y;
}

/* numba/cpython/randomimpl.py:134-147 */
static double mt_next_double(mt_state *s)
{
    uint32_t a = mt_next_u32(s) >> 5;
    uint32_t b = mt_next_u32(s) >> 6;
    s->draws++;
    return ((double)b + (double)a * 67108864.0) / 9007199254740992.0;
}

double tardis_oracle_rng_double(uint32_t seed, int64_t skip)
{
    mt_state s;
    mt_init(&s, seed);
    for (int64_t i = 0; i < skip; i++) mt_next_double(&s);
    return mt_next_double(&s);
}

/* ------------------------------------------------------------ state types */
typedef struct {
    double r, mu, nu, energy;
    int64_t next_line_id, current_shell_id, status;
    int64_t index;
} rpacket_t;

typedef struct {
    double r, mu, nu, energy;
    int64_t next_line_id, current_shell_id, status;
} vpacket_t;

typedef struct {
    /* per-thread accumulators */
    double *j, *nu_bar;       /* [S] */
    double *j_blue, *edotlu;  /* [L*S] row-major (line, shell) */
    double *vhist;            /* [n_grid] */
    double *photo_ion, *stim_recomb, *bf_heating, *stim_recomb_cooling; /* [n_continua * S] */
    double *ff_heating;       /* [S] */
    int64_t *photo_ion_stats; /* [n_continua * S] */
    tardis_oracle_counters cnt;
    int error;
    int shared_line_estimators; /* j_blue / edotlu are one table shared by all threads (atomic adds) */
} accum_t;

/* Used only when the line-estimator table is shared between threads (nthreads > 8): the reference keeps a
 * full 2 x L x S copy per thread (modes/montecarlo_transport.py:309-314), which at 128 threads is 20 GB of
 * page faults and a long serial reduction; the shared table is the faster CPU implementation, so the CPU
 * baseline is not handicapped by it. */
static inline void atomic_add_double(double *p, double v)
{
    uint64_t *u = (uint64_t *)p;
    uint64_t old = __atomic_load_n(u, __ATOMIC_RELAXED), neu;
    double d;
    do {
        memcpy(&d, &old, 8);
        d += v;
        memcpy(&neu, &d, 8);
    } while (!__atomic_compare_exchange_n(u, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

/* last-interaction tracker, packets/trackers/tracker_last_interaction.py:7-254 */
typedef struct {
    double radius, before_nu, before_mu, before_energy, after_nu, after_mu, after_energy;
    int64_t shell_id, interaction_type, line_absorb_id, line_emit_id, interactions_count, boundary_buffer;
} tracker_t;

typedef struct {
    const tardis_oracle_model *m;
    const tardis_oracle_config *c;
    accum_t *acc;
    mt_state rng;
    tracker_t trk;
    /* event log (TrackerFull equivalent), optional */
    tardis_oracle_event *ev;
    int64_t ev_cap, ev_n;
    /* vpacket log, optional */
    tardis_oracle_outputs *out;
    /* vpackets of the current packet (VPacketCollection), grown as needed */
    double *vp_nu, *vp_energy, *vp_mu, *vp_r;
    int64_t vp_n, vp_cap;
    /* chi_continuum_calculator results of the current step (opacities/opacities.py:208-246) */
    double chi_bf_tot, chi_ff;
    double *chi_bf_contributions, *x_sect_bfs; /* [n_continua] */
    int64_t *current_continua;                 /* [n_continua] */
    int64_t n_current;
} ctx_t;

/* ------------------------------------------------------------ frame transforms */
/* transport/frame_transformations.py:12-56 */
static double get_doppler_factor(double velocity, double mu, int full_rel)
{
    double inv_c = 1 / C_SPEED_OF_LIGHT;
    double beta = velocity * inv_c;
    if (!full_rel) return 1.0 - mu * beta;
    return (1.0 - mu * beta) / sqrt(1 - beta * beta);
}

/* transport/frame_transformations.py:44-72 */
static double get_inverse_doppler_factor(double velocity, double mu, int full_rel)
{
    double inv_c = 1 / C_SPEED_OF_LIGHT;
    double beta = velocity * inv_c;
    if (!full_rel) return 1.0 / (1.0 - mu * beta);
    return (1.0 + mu * beta) / sqrt(1 - beta * beta);
}

/* transport/frame_transformations.py:86-109 */
static double angle_aberration_CMF_to_LF(double r, double time_explosion, double mu)
{
    double ct = C_SPEED_OF_LIGHT * time_explosion;
    double beta = r / ct;
    return (mu + beta) / (1.0 + beta * mu);
}
static double angle_aberration_LF_to_CMF(double r, double time_explosion, double mu)
{
    double ct = C_SPEED_OF_LIGHT * time_explosion;
    double beta = r / ct;
    return (mu - beta) / (1.0 - beta * mu);
}

/* ------------------------------------------------------------ distances */
/* transport/geometry/calculate_distances.py:25-62 */
static double calculate_distance_boundary(double r, double mu, double r_inner, double r_outer, int64_t *delta_shell)
{
    double distance;
    if (mu > 0.0) {
        distance = sqrt(r_outer * r_outer + ((mu * mu - 1.0) * r * r)) - (r * mu);
        *delta_shell = 1;
    } else {
        double check = r_inner * r_inner + (r * r * (mu * mu - 1.0));
        if (check >= 0.0) {
            distance = -r * mu - sqrt(check);
            *delta_shell = -1;
        } else {
            distance = sqrt(r_outer * r_outer + ((mu * mu - 1.0) * r * r)) - (r * mu);
            *delta_shell = 1;
        }
    }
    return distance;
}

/* transport/geometry/calculate_distances.py:198-219 */
static double calculate_distance_line_full_relativity(double nu_line, double nu, double time_explosion, double r, double mu)
{
    double nu_r = nu_line / nu;
    double ct = C_SPEED_OF_LIGHT * time_explosion;
    return -mu * r + (ct - nu_r * nu_r * sqrt(ct * ct - (1 + r * r * (1 - mu * mu) * (1 + 1.0 / (nu_r * nu_r))))) / (1 + nu_r * nu_r);
}

/* transport/geometry/calculate_distances.py:66-112; *error set on nu_diff < 0 */
static double calculate_distance_line(double r, double mu, double nu, double comov_nu, int is_last_line,
                                      double nu_line, double time_explosion, int full_rel, int *error)
{
    if (is_last_line) return MISS_DISTANCE;
    double nu_diff = comov_nu - nu_line;
    if (fabs(nu_diff / nu) < CLOSE_LINE_THRESHOLD) return 0.0;
    double distance;
    if (nu_diff >= 0) {
        distance = (nu_diff / nu) * C_SPEED_OF_LIGHT * time_explosion;
    } else {
        *error = TARDIS_ORACLE_ERR_NU_DIFF; /* MonteCarloException("nu difference is less than 0.0") */
        return 0.0;
    }
    if (full_rel) return calculate_distance_line_full_relativity(nu_line, nu, time_explosion, r, mu);
    return distance;
}

/* ------------------------------------------------------------ trackers */
static void log_event(ctx_t *x, const rpacket_t *p, int type, int64_t from_shell, int64_t to_shell, int before)
{
    /* TrackerFull, packets/trackers/tracker_full.py:153-306 */
    if (!x->ev) return;
    if (x->ev_n >= x->ev_cap) return;
    tardis_oracle_event *e = &x->ev[x->ev_n];
    if (before) {
        e->packet_id = p->index;
        e->interaction_type = type;
        e->status = p->status;
        e->radius = p->r;
        e->before_shell_id = from_shell;
        e->after_shell_id = to_shell;
        e->before_nu = p->nu;
        e->before_mu = p->mu;
        e->before_energy = p->energy;
        e->line_absorb_id = (type == IT_LINE) ? p->next_line_id : -1;
        e->line_emit_id = -1;
        if (type == IT_BOUNDARY) {
            e->after_nu = p->nu;
            e->after_mu = p->mu;
            e->after_energy = p->energy;
            x->ev_n++;
        }
    } else {
        e->after_nu = p->nu;
        e->after_mu = p->mu;
        e->after_energy = p->energy;
        if (type == IT_LINE) e->line_emit_id = p->next_line_id - 1;
        x->ev_n++;
    }
}

static void track_boundary_event(ctx_t *x, const rpacket_t *p, int64_t from_shell, int64_t to_shell)
{
    x->trk.boundary_buffer += 1; /* tracker_last_interaction.py:209-231 */
    x->acc->cnt.n_boundary_events++;
    log_event(x, p, IT_BOUNDARY, from_shell, to_shell, 1);
}

static void track_interaction_before(ctx_t *x, const rpacket_t *p, int type)
{
    tracker_t *t = &x->trk;
    t->before_nu = p->nu;
    t->before_mu = p->mu;
    t->before_energy = p->energy;
    if (type == IT_LINE) {
        t->line_absorb_id = p->next_line_id; /* :84-99 */
    } else {
        t->line_absorb_id = -1; /* :127-143 */
        t->line_emit_id = -1;
    }
    log_event(x, p, type, p->current_shell_id, p->current_shell_id, 1);
}

static void track_interaction_after(ctx_t *x, const rpacket_t *p, int type)
{
    tracker_t *t = &x->trk;
    t->after_nu = p->nu;
    t->after_mu = p->mu;
    t->after_energy = p->energy;
    if (type == IT_LINE) t->line_emit_id = p->next_line_id - 1;
    t->interactions_count += 1 + t->boundary_buffer;
    t->boundary_buffer = 0;
    t->radius = p->r;
    t->shell_id = p->current_shell_id;
    t->interaction_type = type;
    log_event(x, p, type, 0, 0, 0);
}

/* ------------------------------------------------------------ estimators */
/* estimators/radfield_estimator_calcs.py:128-164, frame_transformations.py:74-83 */
static void update_estimators_line(ctx_t *x, const rpacket_t *p, int64_t cur_line_id, double distance_trace)
{
    const tardis_oracle_model *m = x->m;
    double energy;
    if (!x->c->enable_full_relativity) {
        double doppler_factor = 1.0 - ((distance_trace + p->mu * p->r) / (m->time_explosion * C_SPEED_OF_LIGHT));
        energy = p->energy * doppler_factor;
    } else {
        energy = p->energy;
    }
    int64_t k = cur_line_id * m->n_shells + p->current_shell_id;
    if (x->acc->shared_line_estimators) {
        atomic_add_double(&x->acc->j_blue[k], energy / p->nu);
        atomic_add_double(&x->acc->edotlu[k], energy);
    } else {
        x->acc->j_blue[k] += energy / p->nu;
        x->acc->edotlu[k] += energy;
    }
    x->acc->cnt.n_line_steps++;
}

/* ------------------------------------------------------------ trace_packet */
/* modes/homologous_rad_packet_transport.py:30-174 (classic: escat_prob = 1.0,
 * continuum_process_enabled = False, modes/classic/packet_propagation.py:142-153) */
static double trace_packet(ctx_t *x, rpacket_t *p, double continuous_opacity, double escat_prob, int continuum_process_enabled,
                           int *interaction_type, int64_t *delta_shell)
{
    const tardis_oracle_model *m = x->m;
    const tardis_oracle_config *c = x->c;
    double r_inner = m->r_inner[p->current_shell_id];
    double r_outer = m->r_outer[p->current_shell_id];
    double distance_boundary = calculate_distance_boundary(p->r, p->mu, r_inner, r_outer, delta_shell);

    int64_t start_line_id = p->next_line_id;
    double tau_event = -log(mt_next_double(&x->rng));
    double tau_trace_line_combined = 0.0;

    double velocity = p->r / m->time_explosion;
    double doppler_factor = get_doppler_factor(velocity, p->mu, c->enable_full_relativity);
    double comov_nu = p->nu * doppler_factor;

    double distance_continuous = tau_event / continuous_opacity;
    int64_t cur_line_id = start_line_id;
    int64_t n_lines = m->n_lines;
    int64_t last_line_id = n_lines - 1;
    double distance = 0.0;
    int broke = 0;
    for (cur_line_id = start_line_id; cur_line_id < n_lines; cur_line_id++) {
        double nu_line = m->line_list_nu[cur_line_id];
        double tau_trace_line = m->tau_sobolev[cur_line_id * m->n_shells + p->current_shell_id];
        tau_trace_line_combined += tau_trace_line;

        int is_last_line = cur_line_id == last_line_id;
        double distance_trace = calculate_distance_line(p->r, p->mu, p->nu, comov_nu, is_last_line, nu_line,
                                                        m->time_explosion, c->enable_full_relativity, &x->acc->error);
        double tau_trace_continuous = continuous_opacity * distance_trace;
        double tau_trace_combined = tau_trace_line_combined + tau_trace_continuous;
        distance = fmin(fmin(distance_trace, distance_boundary), distance_continuous);

        if (distance_trace != 0) {
            if (distance == distance_boundary) {
                *interaction_type = IT_BOUNDARY;
                p->next_line_id = cur_line_id;
                broke = 1;
                break;
            }
            if (distance == distance_continuous) {
                if (continuum_process_enabled) {
                    double zrand = mt_next_double(&x->rng);
                    *interaction_type = (zrand < escat_prob) ? IT_ESCATTERING : IT_CONTINUUM_PROCESS;
                } else {
                    *interaction_type = IT_ESCATTERING;
                }
                p->next_line_id = cur_line_id;
                broke = 1;
                break;
            }
        }
        update_estimators_line(x, p, cur_line_id, distance_trace);

        if (tau_trace_combined > tau_event && !c->disable_line_scattering) {
            *interaction_type = IT_LINE;
            p->next_line_id = cur_line_id;
            distance = distance_trace;
            broke = 1;
            break;
        }
        distance_continuous = (tau_event - tau_trace_line_combined) / continuous_opacity;
    }
    if (!broke) {
        /* for-else: ran off the end of the list (:157-172); next_line_id unchanged */
        if (distance_continuous < distance_boundary) {
            distance = distance_continuous;
            if (continuum_process_enabled) {
                double zrand = mt_next_double(&x->rng);
                *interaction_type = (zrand < escat_prob) ? IT_ESCATTERING : IT_CONTINUUM_PROCESS;
            } else {
                *interaction_type = IT_ESCATTERING;
            }
        } else {
            distance = distance_boundary;
            *interaction_type = IT_BOUNDARY;
        }
    }
    return distance;
}

/* ------------------------------------------------------------ movement */
/* packets/movement.py:31-76, estimators/radfield_estimator_calcs.py:25-53 */
static void move_r_packet(ctx_t *x, rpacket_t *p, double distance)
{
    const tardis_oracle_model *m = x->m;
    int full_rel = x->c->enable_full_relativity;
    double velocity = p->r / m->time_explosion;
    double doppler_factor = get_doppler_factor(velocity, p->mu, full_rel);
    double r = p->r;
    if (distance > 0.0) {
        double new_r = sqrt(r * r + distance * distance + 2.0 * r * distance * p->mu);
        p->mu = (p->mu * r + distance) / new_r;
        p->r = new_r;
        double comov_nu = p->nu * doppler_factor;
        double comov_energy = p->energy * doppler_factor;
        if (full_rel) distance *= doppler_factor;
        x->acc->j[p->current_shell_id] += comov_energy * distance;
        x->acc->nu_bar[p->current_shell_id] += comov_energy * distance * comov_nu;
    }
}

/* packets/movement.py:80-102 */
static void move_packet_across_shell_boundary(int64_t *current_shell_id, int64_t *status, int64_t delta_shell, int64_t no_of_shells)
{
    int64_t next_shell_id = *current_shell_id + delta_shell;
    if (next_shell_id >= no_of_shells)
        *status = ST_EMITTED;
    else if (next_shell_id < 0)
        *status = ST_REABSORBED;
    else
        *current_shell_id = next_shell_id;
}

/* ------------------------------------------------------------ interactions */
/* interaction_events.py:227-258 */
static void line_emission(ctx_t *x, rpacket_t *p, int64_t emission_line_id)
{
    const tardis_oracle_model *m = x->m;
    int full_rel = x->c->enable_full_relativity;
    double velocity = p->r / m->time_explosion;
    double inverse_doppler_factor = get_inverse_doppler_factor(velocity, p->mu, full_rel);
    p->nu = m->line_list_nu[emission_line_id] * inverse_doppler_factor;
    p->next_line_id = emission_line_id + 1;
    if (full_rel) p->mu = angle_aberration_CMF_to_LF(p->r, m->time_explosion, p->mu);
}

/* macro_atom.py:52-104 ; returns transition_line_id, *ttype = transition type */
static int64_t macro_atom_interaction(ctx_t *x, int64_t activation_level_id, int64_t shell, int64_t *ttype)
{
    const tardis_oracle_model *m = x->m;
    int64_t current_transition_type = 0;
    int64_t transition_id = 0;
    while (current_transition_type >= 0) {
        double probability = 0.0;
        double probability_event = mt_next_double(&x->rng);
        x->acc->cnt.n_macro_jumps++;
        int64_t block_start = m->macro_block_edge_index[activation_level_id];
        int64_t block_end = m->macro_block_edge_index[activation_level_id + 1];
        int found = 0;
        for (transition_id = block_start; transition_id < block_end; transition_id++) {
            double tp = m->transition_probabilities[transition_id * m->n_shells + shell];
            probability += tp;
            x->acc->cnt.n_macro_scanned++;
            if (probability > probability_event) {
                activation_level_id = m->destination_level_id[transition_id];
                current_transition_type = m->transition_type[transition_id];
                found = 1;
                break;
            }
        }
        if (!found) {
            x->acc->error = TARDIS_ORACLE_ERR_MACRO_ATOM; /* MacroAtomError */
            *ttype = -1;
            return 0;
        }
    }
    *ttype = current_transition_type;
    return m->transition_line_id[transition_id];
}

/* interaction_event_callers.py:31-91 (classic branch: BB_EMISSION only) */
static void macro_atom_event(ctx_t *x, rpacket_t *p, int64_t destination_level_idx)
{
    int64_t ttype;
    int64_t transition_id = macro_atom_interaction(x, destination_level_idx, p->current_shell_id, &ttype);
    if (x->acc->error) return;
    if (ttype == -1) {
        line_emission(x, p, transition_id);
    } else {
        x->acc->error = TARDIS_ORACLE_ERR_MACRO_ATOM; /* "Interaction ... not known or implemented" */
    }
}

static void macro_atom_event_iip(ctx_t *x, rpacket_t *p, int64_t destination_level_idx);

/* interaction_event_callers.py:187-239 */
static void line_scatter_event(ctx_t *x, rpacket_t *p)
{
    const tardis_oracle_model *m = x->m;
    int full_rel = x->c->enable_full_relativity;
    double velocity = p->r / m->time_explosion;
    double old_doppler_factor = get_doppler_factor(velocity, p->mu, full_rel);
    p->mu = 2.0 * mt_next_double(&x->rng) - 1.0; /* utils.py:14-15 */
    double inverse_new_doppler_factor = get_inverse_doppler_factor(velocity, p->mu, full_rel);
    double comov_energy = p->energy * old_doppler_factor;
    p->energy = comov_energy * inverse_new_doppler_factor;
    if (x->c->line_interaction_type == 0) {
        line_emission(x, p, p->next_line_id);
    } else {
        double comov_nu = p->nu * old_doppler_factor;
        p->nu = comov_nu * inverse_new_doppler_factor;
        int64_t activation_level_id = m->line2macro_level_upper[p->next_line_id];
        if (x->c->continuum_processes_enabled) macro_atom_event_iip(x, p, activation_level_id);
        else macro_atom_event(x, p, activation_level_id);
    }
}

/* interaction_events.py:184-217 */
static void thomson_scatter(ctx_t *x, rpacket_t *p)
{
    const tardis_oracle_model *m = x->m;
    int full_rel = x->c->enable_full_relativity;
    double velocity = p->r / m->time_explosion;
    double old_doppler_factor = get_doppler_factor(velocity, p->mu, full_rel);
    double comov_nu = p->nu * old_doppler_factor;
    double comov_energy = p->energy * old_doppler_factor;
    p->mu = 2.0 * mt_next_double(&x->rng) - 1.0;
    double inverse_new_doppler_factor = get_inverse_doppler_factor(velocity, p->mu, full_rel);
    p->nu = comov_nu * inverse_new_doppler_factor;
    p->energy = comov_energy * inverse_new_doppler_factor;
    if (full_rel) p->mu = angle_aberration_CMF_to_LF(p->r, m->time_explosion, p->mu);
}

/* ------------------------------------------------------------ virtual packets */
static void vp_append(ctx_t *x, double nu, double energy, double mu, double r)
{
    if (x->vp_n >= x->vp_cap) {
        x->vp_cap = x->vp_cap * 2 + 16;
        x->vp_nu = (double *)realloc(x->vp_nu, sizeof(double) * x->vp_cap);
        x->vp_energy = (double *)realloc(x->vp_energy, sizeof(double) * x->vp_cap);
        x->vp_mu = (double *)realloc(x->vp_mu, sizeof(double) * x->vp_cap);
        x->vp_r = (double *)realloc(x->vp_r, sizeof(double) * x->vp_cap);
    }
    x->vp_nu[x->vp_n] = nu;
    x->vp_energy[x->vp_n] = energy;
    x->vp_mu[x->vp_n] = mu;
    x->vp_r[x->vp_n] = r;
    x->vp_n++;
}

/* packets/virtual_packet.py:77-165 */
static double trace_vpacket_within_shell(ctx_t *x, vpacket_t *v, double *distance_boundary_out, int64_t *delta_shell)
{
    const tardis_oracle_model *m = x->m;
    const tardis_oracle_config *c = x->c;
    int full_rel = c->enable_full_relativity;
    double r_inner = m->r_inner[v->current_shell_id];
    double r_outer = m->r_outer[v->current_shell_id];
    double distance_boundary = calculate_distance_boundary(v->r, v->mu, r_inner, r_outer, delta_shell);
    int64_t start_line_id = v->next_line_id;
    double chi_e = m->electron_density[v->current_shell_id] * c->sigma_thomson;
    double velocity = v->r / m->time_explosion;
    double doppler_factor = get_doppler_factor(velocity, v->mu, full_rel);
    double comov_nu = v->nu * doppler_factor;
    double chi_continuum = chi_e;
    if (full_rel) chi_continuum *= doppler_factor;
    double tau_continuum = chi_continuum * distance_boundary;
    double tau_trace_combined = tau_continuum;
    int64_t n_lines = m->n_lines;
    int64_t cur_line_id = start_line_id;
    int broke = 0;
    for (cur_line_id = start_line_id; cur_line_id < n_lines; cur_line_id++) {
        double nu_line = m->line_list_nu[cur_line_id];
        double tau_trace_line = m->tau_sobolev[cur_line_id * m->n_shells + v->current_shell_id];
        int is_last_line = cur_line_id == n_lines - 1;
        double distance_trace_line = calculate_distance_line(v->r, v->mu, v->nu, comov_nu, is_last_line, nu_line,
                                                             m->time_explosion, full_rel, &x->acc->error);
        x->acc->cnt.n_vpacket_line_steps++;
        if (distance_boundary <= distance_trace_line) {
            broke = 1;
            break;
        }
        tau_trace_combined += tau_trace_line;
    }
    if (!broke) {
        /* for-else: python leaves cur_line_id at the last iterated value (or start if empty) */
        if (start_line_id < n_lines) cur_line_id = n_lines - 1; else cur_line_id = start_line_id;
        if (cur_line_id == n_lines - 1) cur_line_id += 1;
    }
    v->next_line_id = cur_line_id;
    *distance_boundary_out = distance_boundary;
    return tau_trace_combined;
}

/* packets/virtual_packet.py:168-245 */
static double trace_vpacket(ctx_t *x, vpacket_t *v)
{
    const tardis_oracle_model *m = x->m;
    const tardis_oracle_config *c = x->c;
    double tau_trace_combined = 0.0;
    int64_t guard = 0;
    while (1) {
        double distance_boundary;
        int64_t delta_shell;
        double tau_shell = trace_vpacket_within_shell(x, v, &distance_boundary, &delta_shell);
        tau_trace_combined += tau_shell;
        move_packet_across_shell_boundary(&v->current_shell_id, &v->status, delta_shell, m->n_shells);
        if (tau_trace_combined > c->vpacket_tau_russian) {
            double event_random = mt_next_double(&x->rng);
            if (event_random > c->survival_probability) {
                v->energy = 0.0;
                v->status = ST_EMITTED;
            } else {
                v->energy = v->energy / c->survival_probability * exp(-tau_trace_combined);
                tau_trace_combined = 0.0;
            }
        }
        double new_r = sqrt(v->r * v->r + distance_boundary * distance_boundary + 2.0 * v->r * distance_boundary * v->mu);
        v->mu = (v->mu * v->r + distance_boundary) / new_r;
        v->r = new_r;
        if (v->status == ST_EMITTED) break;
        if (++guard > 4 * m->n_shells + 64) { /* the reference would spin forever here */
            x->acc->error = TARDIS_ORACLE_ERR_VPACKET_LOOP;
            break;
        }
    }
    return tau_trace_combined;
}

/* packets/virtual_packet.py:248-386 */
static void trace_vpacket_volley(ctx_t *x, const rpacket_t *p)
{
    const tardis_oracle_model *m = x->m;
    const tardis_oracle_config *c = x->c;
    int full_rel = c->enable_full_relativity;
    if ((p->nu < c->vpacket_spawn_start_frequency) || (p->nu > c->vpacket_spawn_end_frequency)) return;
    int64_t no_of_vpackets = c->number_of_vpackets;
    if (no_of_vpackets == 0) return;
    double mu_min, beta_inner = 0.0;
    int on_inner_boundary;
    if (p->r > m->r_inner[0]) {
        double r_inner_over_r = m->r_inner[0] / p->r;
        mu_min = -sqrt(1 - r_inner_over_r * r_inner_over_r);
        on_inner_boundary = 0;
        if (full_rel) mu_min = angle_aberration_LF_to_CMF(p->r, m->time_explosion, mu_min);
    } else {
        on_inner_boundary = 1;
        mu_min = 0.0;
        if (full_rel) {
            double inv_c = 1 / C_SPEED_OF_LIGHT;
            double inv_t = 1 / m->time_explosion;
            beta_inner = m->r_inner[0] * inv_t * inv_c;
        }
    }
    double mu_bin = (1.0 - mu_min) / no_of_vpackets;
    double r_packet_velocity = p->r / m->time_explosion;
    double r_packet_doppler_factor = get_doppler_factor(r_packet_velocity, p->mu, full_rel);
    for (int64_t i = 0; i < no_of_vpackets; i++) {
        double v_packet_mu = mu_min + i * mu_bin + mt_next_double(&x->rng) * mu_bin;
        double weight;
        if (on_inner_boundary) {
            if (!full_rel)
                weight = 2 * v_packet_mu / no_of_vpackets;
            else
                weight = 2 * (v_packet_mu + beta_inner) / (2 * beta_inner + 1) / no_of_vpackets;
        } else {
            weight = (1 - mu_min) / (2 * no_of_vpackets);
        }
        if (full_rel) v_packet_mu = angle_aberration_CMF_to_LF(p->r, m->time_explosion, v_packet_mu);
        double v_packet_doppler_factor = get_doppler_factor(r_packet_velocity, v_packet_mu, full_rel);
        double doppler_factor_ratio = r_packet_doppler_factor / v_packet_doppler_factor;
        double v_packet_nu = p->nu * doppler_factor_ratio;
        double v_packet_energy = p->energy * weight * doppler_factor_ratio;
        vpacket_t v;
        v.r = p->r;
        v.mu = v_packet_mu;
        v.nu = v_packet_nu;
        v.energy = v_packet_energy;
        v.current_shell_id = p->current_shell_id;
        v.next_line_id = p->next_line_id;
        v.status = ST_IN_PROCESS;
        double tau_vpacket = trace_vpacket(x, &v);
        v.energy *= exp(-tau_vpacket);
        x->acc->cnt.n_vpackets++;
        vp_append(x, v.nu, v.energy, v_packet_mu, p->r);
    }
}

/* modes/montecarlo_transport.py:166-195 */
static void add_vpacket_collection_to_histogram(ctx_t *x)
{
    const tardis_oracle_config *c = x->c;
    const double *grid = c->spectrum_frequency_grid;
    double delta_nu = grid[1] - grid[0];
    for (int64_t j = 0; j < x->vp_n; j++) {
        double nu = x->vp_nu[j];
        if ((nu < grid[0]) || (nu > grid[c->n_grid - 1])) continue;
        int64_t idx = (int64_t)floor((nu - grid[0]) / delta_nu);
        x->acc->vhist[idx] += x->vp_energy[j];
    }
}


/* ------------------------------------------------------------ continuum (IIP mode) */
/* opacities/opacities.py:89-246: chi_continuum_calculator = chi_bf_interpolator + chi_ff_calculator */
static void chi_continuum_calculator(ctx_t *x, double nu, int64_t shell)
{
    const tardis_oracle_model *m = x->m;
    x->n_current = 0;
    /* get_current_bound_free_continua, :89-107 */
    for (int64_t k = 0; k < m->n_continua; k++)
        if (nu >= m->photo_ion_nu_threshold_mins[k] && nu <= m->photo_ion_nu_threshold_maxs[k])
            x->current_continua[x->n_current++] = k;
    double running = 0.0;
    for (int64_t i = 0; i < x->n_current; i++) {
        int64_t k = x->current_continua[i];
        int64_t start = m->photo_ion_block_references[k], end = m->photo_ion_block_references[k + 1];
        const double *pn = m->phot_nus + start;
        int64_t n = end - start;
        int64_t lo = 0, hi = n; /* np.searchsorted(pn, nu) (left): first idx with pn[idx] >= nu */
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (pn[mid] < nu) lo = mid + 1; else hi = mid;
        }
        int64_t nu_idx = lo;
        if (nu_idx >= n) { x->acc->error = TARDIS_ORACLE_ERR_CONTINUUM; nu_idx = n - 1; }
        int64_t im1 = (nu_idx == 0) ? n - 1 : nu_idx - 1; /* python's [-1] wrap */
        double interval = pn[nu_idx] - pn[im1];
        double high_weight = nu - pn[im1];
        double low_weight = pn[nu_idx] - nu;
        double chi_hi = m->chi_bf[(start + nu_idx) * m->n_shells + shell], chi_lo = m->chi_bf[(start + im1) * m->n_shells + shell];
        double chi = (chi_hi * high_weight + chi_lo * low_weight) / interval;
        double xs = (m->x_sect[start + nu_idx] * high_weight + m->x_sect[start + im1] * low_weight) / interval;
        x->x_sect_bfs[i] = xs;
        running += chi; /* chi_bfs.cumsum() */
        x->chi_bf_contributions[i] = running;
    }
    if (x->n_current == 0) {
        x->chi_bf_tot = 0.0;
    } else {
        x->chi_bf_tot = x->chi_bf_contributions[x->n_current - 1];
        for (int64_t i = 0; i < x->n_current; i++) x->chi_bf_contributions[i] /= x->chi_bf_tot;
    }
    /* chi_ff_calculator, :181-205.  FF_OPAC_CONST, :25-27 */
    const double ff_opac_const = pow(2 * M_PI / (3 * M_ELECTRON * K_BOLTZMANN), 0.5) * 4 * pow(E_ESU, 6) / (3 * M_ELECTRON * H_PLANCK * C_SPEED_OF_LIGHT);
    x->chi_ff = ff_opac_const * m->ff_opacity_factor[shell] / (nu * nu * nu) * (1 - exp(-H_PLANCK * nu / (K_BOLTZMANN * m->t_electrons[shell])));
}

/* estimators/radfield_estimator_calcs.py:57-124 */
static void update_estimators_bound_free(ctx_t *x, double comov_nu, double comov_energy, int64_t shell, double distance, double chi_ff)
{
    const tardis_oracle_model *m = x->m;
    double t_electron = m->t_electrons[shell];
    double boltzmann_factor = exp(-(H_PLANCK * comov_nu) / (K_BOLTZMANN * t_electron));
    x->acc->ff_heating[shell] += comov_energy * distance * chi_ff;
    for (int64_t i = 0; i < x->n_current; i++) {
        int64_t k = x->current_continua[i];
        int64_t cell = k * m->n_shells + shell;
        double inc = comov_energy * distance * x->x_sect_bfs[i] / comov_nu;
        x->acc->photo_ion[cell] += inc;
        x->acc->stim_recomb[cell] += inc * boltzmann_factor;
        x->acc->photo_ion_stats[cell] += 1;
        double nu_th = m->bf_threshold_list_nu[k];
        double bfh = comov_energy * distance * x->x_sect_bfs[i] * (1 - nu_th / comov_nu);
        x->acc->bf_heating[cell] += bfh;
        x->acc->stim_recomb_cooling[cell] += bfh * boltzmann_factor;
        x->acc->cnt.n_bf_estimator_updates++;
    }
}

/* interaction_events.py:21-37: L - searchsorted(line_list[::-1], nu) == number of lines with nu_line >= nu (not clamped) */
static int64_t get_current_line_id(const tardis_oracle_model *m, double nu)
{
    int64_t lo = 0, hi = m->n_lines;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (m->line_list_nu[mid] >= nu) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* interaction_events.py:40-57 */
static double sample_nu_free_bound(ctx_t *x, int64_t shell, int64_t continuum_id)
{
    const tardis_oracle_model *m = x->m;
    if (continuum_id < 0 || continuum_id >= m->n_continua) { x->acc->error = TARDIS_ORACLE_ERR_CONTINUUM; return 1.0; }
    int64_t start = m->photo_ion_block_references[continuum_id], end = m->photo_ion_block_references[continuum_id + 1];
    const double *pn = m->phot_nus + start;
    int64_t n = end - start;
    double zrand = mt_next_double(&x->rng);
    int64_t lo = 0, hi = n; /* searchsorted(em, zrand, side='right'): first idx with em[idx] > zrand */
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (m->emissivities[(start + mid) * m->n_shells + shell] <= zrand) lo = mid + 1; else hi = mid;
    }
    int64_t idx = lo;
    if (idx >= n) { x->acc->error = TARDIS_ORACLE_ERR_CONTINUUM; idx = n - 1; }
    int64_t im1 = (idx == 0) ? n - 1 : idx - 1;
    double em_i = m->emissivities[(start + idx) * m->n_shells + shell], em_m = m->emissivities[(start + im1) * m->n_shells + shell];
    return pn[idx] - (em_i - zrand) / (em_i - em_m) * (pn[idx] - pn[im1]);
}

/* interaction_events.py:60-92 */
static void bound_free_emission(ctx_t *x, rpacket_t *p, int64_t continuum_id)
{
    const tardis_oracle_model *m = x->m;
    double velocity = p->r / m->time_explosion;
    double inverse_doppler_factor = get_inverse_doppler_factor(velocity, p->mu, 1);
    double comov_nu = sample_nu_free_bound(x, p->current_shell_id, continuum_id);
    p->nu = comov_nu * inverse_doppler_factor;
    p->next_line_id = get_current_line_id(m, comov_nu);
    p->mu = angle_aberration_CMF_to_LF(p->r, m->time_explosion, p->mu);
}

/* interaction_events.py:141-180 */
static void free_free_emission(ctx_t *x, rpacket_t *p)
{
    const tardis_oracle_model *m = x->m;
    double velocity = p->r / m->time_explosion;
    double inverse_doppler_factor = get_inverse_doppler_factor(velocity, p->mu, 1);
    double temperature = m->t_electrons[p->current_shell_id];
    double zrand = mt_next_double(&x->rng);
    double comov_nu = -K_BOLTZMANN * temperature / H_PLANCK * log(zrand);
    p->nu = comov_nu * inverse_doppler_factor;
    p->next_line_id = get_current_line_id(m, comov_nu);
    p->mu = angle_aberration_CMF_to_LF(p->r, m->time_explosion, p->mu);
}

/* macro_atom.py:108-184 */
static int64_t macro_atom_interaction_iip(ctx_t *x, int64_t activation_level_idx, int64_t shell, int64_t *emission_process)
{
    const tardis_oracle_model *m = x->m;
    *emission_process = 0;
    if (activation_level_idx < 0 || activation_level_idx >= m->n_markov) { x->acc->error = TARDIS_ORACLE_ERR_MACRO_ATOM; return 0; }
    double absorbing_state_probability = 0.0;
    double probability_event = mt_next_double(&x->rng);
    x->acc->cnt.n_macro_jumps++;
    const double *row = m->absorbing_markov_probabilities + (shell * m->n_markov + activation_level_idx) * m->n_markov;
    int64_t absorbing = -1;
    for (int64_t to = 0; to < m->n_markov; to++) {
        absorbing_state_probability += row[to];
        x->acc->cnt.n_macro_scanned++;
        if (absorbing_state_probability > probability_event) { absorbing = to; break; }
    }
    if (absorbing < 0 || absorbing >= m->n_blocks) { x->acc->error = TARDIS_ORACLE_ERR_MACRO_ATOM; return 0; }
    int64_t block_start = m->macro_block_edge_index[absorbing], block_end = m->macro_block_edge_index[absorbing + 1];
    double emission_transition_probability = 0.0;
    double probability_emission_event = mt_next_double(&x->rng);
    x->acc->cnt.n_macro_jumps++;
    for (int64_t ch = block_start; ch < block_end; ch++) {
        emission_transition_probability += m->transition_probabilities[ch * m->n_shells + shell];
        x->acc->cnt.n_macro_scanned++;
        if (emission_transition_probability > probability_emission_event) {
            *emission_process = m->transition_type[ch];
            return m->transition_line_id[ch];
        }
    }
    x->acc->error = TARDIS_ORACLE_ERR_MACRO_ATOM;
    return 0;
}

/* interaction_event_callers.py:31-91, CONTINUUM_PROCESSES_ENABLED branch */
static void macro_atom_event_iip(ctx_t *x, rpacket_t *p, int64_t destination_level_idx)
{
    int64_t ttype;
    int64_t transition_id = macro_atom_interaction_iip(x, destination_level_idx, p->current_shell_id, &ttype);
    if (x->acc->error) return;
    if (ttype == -3 || ttype == -21) free_free_emission(x, p);                                  /* FF_EMISSION, FF_COOLING */
    else if (ttype == -2 || ttype == -20 || ttype == -7) bound_free_emission(x, p, transition_id); /* BF_EMISSION, FB_COOLING, PHOTO_RECOMB_EMISSION */
    else if (ttype == -4) p->status = ST_ADIABATIC_COOLING;                                      /* adiabatic_cooling, interaction_events.py:130-138 */
    else if (ttype == -1) line_emission(x, p, transition_id);
    else x->acc->error = TARDIS_ORACLE_ERR_MACRO_ATOM;
}

/* interaction_event_callers.py:95-183 + interaction_events.py:262-299 */
static void continuum_event(ctx_t *x, rpacket_t *p)
{
    const tardis_oracle_model *m = x->m;
    double velocity = p->r / m->time_explosion;
    double old_doppler_factor = get_doppler_factor(velocity, p->mu, 1);
    p->mu = 2.0 * mt_next_double(&x->rng) - 1.0;
    double inverse_doppler_factor = get_inverse_doppler_factor(velocity, p->mu, 1);
    double comov_energy = p->energy * old_doppler_factor;
    double comov_nu = p->nu * old_doppler_factor;
    p->energy = comov_energy * inverse_doppler_factor;
    /* determine_continuum_macro_activation_idx */
    int64_t destination_level_idx;
    double fraction_bf = x->chi_bf_tot / (x->chi_bf_tot + x->chi_ff);
    if (mt_next_double(&x->rng) < fraction_bf) {
        /* determine_bf_macro_activation_idx: np.searchsorted(chi_bf_contributions, random) (left) */
        double z = mt_next_double(&x->rng);
        int64_t sampled = 0;
        while (sampled < x->n_current && x->chi_bf_contributions[sampled] < z) sampled++;
        if (sampled >= x->n_current) { x->acc->error = TARDIS_ORACLE_ERR_CONTINUUM; return; }
        int64_t active = x->current_continua[sampled];
        double nu_threshold = m->photo_ion_nu_threshold_mins[active];
        double fraction_ionization = nu_threshold / comov_nu;
        if (mt_next_double(&x->rng) < fraction_ionization) {
            if (active >= m->n_activation) { x->acc->error = TARDIS_ORACLE_ERR_CONTINUUM; return; }
            destination_level_idx = m->photo_ion_activation_idx[active];
        } else {
            destination_level_idx = m->k_packet_idx;
        }
    } else {
        destination_level_idx = m->k_packet_idx;
    }
    macro_atom_event_iip(x, p, destination_level_idx);
}

/* modes/iip/packet_propagation.py:55-270: always full relativity, no virtual packets */
static void packet_propagation_iip(ctx_t *x, rpacket_t *p)
{
    const tardis_oracle_model *m = x->m;
    const tardis_oracle_config *c = x->c; /* the caller passes a config with enable_full_relativity = 1 */
    {
        double beta = (p->r / m->time_explosion) / C_SPEED_OF_LIGHT;
        double velocity = p->r / m->time_explosion;
        double inverse_doppler_factor = get_inverse_doppler_factor(velocity, p->mu, 1);
        p->nu *= inverse_doppler_factor;
        p->energy *= inverse_doppler_factor;
        p->mu = (p->mu + beta) / (1 + beta * p->mu);
    }
    {
        double velocity = p->r / m->time_explosion;
        double comov_nu = p->nu * get_doppler_factor(velocity, p->mu, 1);
        int64_t next_line_id = get_current_line_id(m, comov_nu);
        if (next_line_id == m->n_lines) next_line_id -= 1;
        p->next_line_id = next_line_id;
    }
    track_boundary_event(x, p, -1, 0);
    while (p->status == ST_IN_PROCESS && !x->acc->error) {
        double velocity = p->r / m->time_explosion;
        double doppler_factor = get_doppler_factor(velocity, p->mu, 1);
        double comov_nu = p->nu * doppler_factor;
        double chi_e = m->electron_density[p->current_shell_id] * c->sigma_thomson;
        chi_continuum_calculator(x, comov_nu, p->current_shell_id);
        double chi_continuum = chi_e + x->chi_bf_tot + x->chi_ff;
        double escat_prob = chi_e / chi_continuum;
        chi_continuum *= doppler_factor;
        int interaction_type = 0;
        int64_t delta_shell = 0;
        double distance = trace_packet(x, p, chi_continuum, escat_prob, 1, &interaction_type, &delta_shell);
        if (x->acc->error) break;
        update_estimators_bound_free(x, comov_nu, p->energy * doppler_factor, p->current_shell_id, distance * doppler_factor,
                                     x->chi_ff * doppler_factor);
        if (interaction_type == IT_BOUNDARY) {
            move_r_packet(x, p, distance);
            track_boundary_event(x, p, p->current_shell_id, p->current_shell_id + delta_shell);
            move_packet_across_shell_boundary(&p->current_shell_id, &p->status, delta_shell, m->n_shells);
        } else if (interaction_type == IT_LINE) {
            move_r_packet(x, p, distance);
            track_interaction_before(x, p, IT_LINE);
            line_scatter_event(x, p);
            track_interaction_after(x, p, IT_LINE);
            x->acc->cnt.n_line_events++;
        } else if (interaction_type == IT_ESCATTERING) {
            move_r_packet(x, p, distance);
            track_interaction_before(x, p, IT_ESCATTERING);
            thomson_scatter(x, p);
            track_interaction_after(x, p, IT_ESCATTERING);
            x->acc->cnt.n_escat_events++;
        } else if (interaction_type == IT_CONTINUUM_PROCESS) {
            move_r_packet(x, p, distance);
            track_interaction_before(x, p, IT_CONTINUUM_PROCESS);
            continuum_event(x, p);
            track_interaction_after(x, p, IT_CONTINUUM_PROCESS);
            x->acc->cnt.n_continuum_events++;
        }
    }
    track_boundary_event(x, p, p->current_shell_id, p->current_shell_id + 1);
}

/* ------------------------------------------------------------ packet_propagation */
/* modes/classic/packet_propagation.py:53-251 */
static void packet_propagation(ctx_t *x, rpacket_t *p)
{
    const tardis_oracle_model *m = x->m;
    const tardis_oracle_config *c = x->c;
    int full_rel = c->enable_full_relativity;

    /* :99-102, :255-318 */
    {
        double velocity = p->r / m->time_explosion;
        double inverse_doppler_factor = get_inverse_doppler_factor(velocity, p->mu, full_rel);
        if (full_rel) {
            double beta = (p->r / m->time_explosion) / C_SPEED_OF_LIGHT;
            p->nu *= inverse_doppler_factor;
            p->energy *= inverse_doppler_factor;
            p->mu = (p->mu + beta) / (1 + beta * p->mu);
        } else {
            p->nu *= inverse_doppler_factor;
            p->energy *= inverse_doppler_factor;
        }
    }
    /* initialize_line_id, packets/radiative_packet.py:96-110 */
    {
        double velocity = p->r / m->time_explosion;
        double doppler_factor = get_doppler_factor(velocity, p->mu, full_rel);
        double comov_nu = p->nu * doppler_factor;
        /* L - searchsorted(nu[::-1], comov_nu, 'left') == number of lines with nu_line >= comov_nu */
        int64_t lo = 0, hi = m->n_lines; /* first index with nu[idx] < comov_nu (nu descending) */
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (m->line_list_nu[mid] >= comov_nu) lo = mid + 1; else hi = mid;
        }
        int64_t next_line_id = lo;
        if (next_line_id == m->n_lines) next_line_id -= 1;
        p->next_line_id = next_line_id;
    }
    trace_vpacket_volley(x, p);
    track_boundary_event(x, p, -1, 0);

    while (p->status == ST_IN_PROCESS && !x->acc->error) {
        double velocity = p->r / m->time_explosion;
        double doppler_factor = get_doppler_factor(velocity, p->mu, full_rel);
        double opacity_electron = m->electron_density[p->current_shell_id] * c->sigma_thomson; /* opacities/opacities.py:50-67 */
        if (full_rel) opacity_electron *= doppler_factor;

        int interaction_type = 0;
        int64_t delta_shell = 0;
        double distance = trace_packet(x, p, opacity_electron, 1.0, 0, &interaction_type, &delta_shell);
        if (x->acc->error) break;

        if (interaction_type == IT_BOUNDARY) {
            move_r_packet(x, p, distance);
            track_boundary_event(x, p, p->current_shell_id, p->current_shell_id + delta_shell);
            move_packet_across_shell_boundary(&p->current_shell_id, &p->status, delta_shell, m->n_shells);
        } else if (interaction_type == IT_LINE) {
            move_r_packet(x, p, distance);
            track_interaction_before(x, p, IT_LINE);
            line_scatter_event(x, p);
            track_interaction_after(x, p, IT_LINE);
            x->acc->cnt.n_line_events++;
            trace_vpacket_volley(x, p);
        } else if (interaction_type == IT_ESCATTERING) {
            move_r_packet(x, p, distance);
            track_interaction_before(x, p, IT_ESCATTERING);
            thomson_scatter(x, p);
            track_interaction_after(x, p, IT_ESCATTERING);
            x->acc->cnt.n_escat_events++;
            trace_vpacket_volley(x, p);
        }
    }
    track_boundary_event(x, p, p->current_shell_id, p->current_shell_id + 1); /* :247-251 */
}

/* ------------------------------------------------------------ main loop */
static void accum_alloc(accum_t *a, const tardis_oracle_model *m, const tardis_oracle_config *c, accum_t *share_with)
{
    memset(a, 0, sizeof(*a));
    a->j = (double *)calloc(m->n_shells, sizeof(double));
    a->nu_bar = (double *)calloc(m->n_shells, sizeof(double));
    if (share_with) {
        a->j_blue = share_with->j_blue;
        a->edotlu = share_with->edotlu;
        a->shared_line_estimators = 1;
    } else {
        a->j_blue = (double *)calloc((size_t)m->n_lines * m->n_shells, sizeof(double));
        a->edotlu = (double *)calloc((size_t)m->n_lines * m->n_shells, sizeof(double));
    }
    a->vhist = (double *)calloc(c->n_grid > 0 ? c->n_grid : 1, sizeof(double));
    if (c->continuum_processes_enabled) {
        size_t nc = (size_t)(m->n_continua > 0 ? m->n_continua : 1) * m->n_shells;
        a->photo_ion = (double *)calloc(nc, sizeof(double));
        a->stim_recomb = (double *)calloc(nc, sizeof(double));
        a->bf_heating = (double *)calloc(nc, sizeof(double));
        a->stim_recomb_cooling = (double *)calloc(nc, sizeof(double));
        a->photo_ion_stats = (int64_t *)calloc(nc, sizeof(int64_t));
        a->ff_heating = (double *)calloc(m->n_shells, sizeof(double));
    }
}
static void accum_free(accum_t *a, int owns_line_tables)
{
    free(a->j); free(a->nu_bar); free(a->vhist);
    free(a->photo_ion); free(a->stim_recomb); free(a->bf_heating); free(a->stim_recomb_cooling); free(a->photo_ion_stats); free(a->ff_heating);
    if (owns_line_tables) { free(a->j_blue); free(a->edotlu); }
}

/* worker: one host thread of the prange (modes/montecarlo_transport.py:316-354) */
typedef struct {
    const tardis_oracle_model *m;
    const tardis_oracle_config *c;
    const tardis_oracle_packets *pk;
    tardis_oracle_outputs *out;
    accum_t *acc;
    atomic_llong *next_packet;
    atomic_llong *vlog_n;
} worker_t;

#define ORACLE_CHUNK 64

static void *worker_main(void *arg)
{
    worker_t *w = (worker_t *)arg;
    const tardis_oracle_model *m = w->m;
    const tardis_oracle_config *c = w->c;
    const tardis_oracle_packets *pk = w->pk;
    tardis_oracle_outputs *out = w->out;
    int64_t n = pk->n_packets;
    ctx_t x;
    memset(&x, 0, sizeof(x));
    x.m = m; x.c = c; x.acc = w->acc; x.out = out;
    tardis_oracle_config c_iip;
    if (c->continuum_processes_enabled) {
        c_iip = *c;
        c_iip.enable_full_relativity = 1; /* modes/iip/packet_propagation.py passes enable_full_relativity=True everywhere */
        c_iip.number_of_vpackets = 0;     /* no virtual packets in this mode (modes/iip/solver.py:248) */
        x.c = &c_iip;
        size_t nc = (size_t)(m->n_continua > 0 ? m->n_continua : 1);
        x.chi_bf_contributions = (double *)malloc(nc * sizeof(double));
        x.x_sect_bfs = (double *)malloc(nc * sizeof(double));
        x.current_continua = (int64_t *)malloc(nc * sizeof(int64_t));
    }
    for (;;) {
        int64_t lo = atomic_fetch_add(w->next_packet, ORACLE_CHUNK);
        if (lo >= n || x.acc->error) break;
        int64_t hi = lo + ORACLE_CHUNK < n ? lo + ORACLE_CHUNK : n;
        for (int64_t i = lo; i < hi && !x.acc->error; i++) {
            rpacket_t p;
            /* make_r_packet, modes/montecarlo_transport.py:41-66 */
            p.r = pk->initial_radii[i];
            p.mu = pk->initial_mus[i];
            p.nu = pk->initial_nus[i];
            p.energy = pk->initial_energies[i];
            p.current_shell_id = 0;
            p.status = ST_IN_PROCESS;
            p.next_line_id = 0;
            p.index = i;
            mt_init(&x.rng, (uint32_t)pk->packet_seeds[i]);
            /* TrackerLastInteraction.__init__ */
            x.trk.radius = NAN; x.trk.before_nu = NAN; x.trk.before_mu = NAN; x.trk.before_energy = NAN;
            x.trk.after_nu = NAN; x.trk.after_mu = NAN; x.trk.after_energy = NAN;
            x.trk.shell_id = -1; x.trk.interaction_type = -1; x.trk.line_absorb_id = -1; x.trk.line_emit_id = -1;
            x.trk.interactions_count = 0; x.trk.boundary_buffer = -1;
            x.vp_n = 0;
            if (out->events && i < out->n_tracked_packets) {
                x.ev = out->events + i * out->max_events_per_packet;
                x.ev_cap = out->max_events_per_packet;
                x.ev_n = 0;
            } else {
                x.ev = NULL;
            }

            if (c->continuum_processes_enabled) packet_propagation_iip(&x, &p);
            else packet_propagation(&x, &p);

            /* set_packet_collection_output, modes/montecarlo_transport.py:70-90 */
            out->output_nus[i] = p.nu;
            if (p.status == ST_REABSORBED)
                out->output_energies[i] = -p.energy;
            else if (p.status == ST_EMITTED)
                out->output_energies[i] = p.energy;
            else
                out->output_energies[i] = -99.0;
            x.acc->cnt.n_rng_draws += x.rng.draws;
            if (out->events && i < out->n_tracked_packets) out->event_counts[i] = x.ev_n;
            if (out->last_interaction_type) {
                out->last_interaction_type[i] = x.trk.interaction_type;
                out->last_event_id[i] = x.trk.interactions_count;
                out->last_radius[i] = x.trk.radius;
                out->last_shell_id[i] = x.trk.shell_id;
                out->last_before_nu[i] = x.trk.before_nu;
                out->last_before_mu[i] = x.trk.before_mu;
                out->last_before_energy[i] = x.trk.before_energy;
                out->last_after_nu[i] = x.trk.after_nu;
                out->last_after_mu[i] = x.trk.after_mu;
                out->last_after_energy[i] = x.trk.after_energy;
                out->last_line_absorb_id[i] = x.trk.line_absorb_id;
                out->last_line_emit_id[i] = x.trk.line_emit_id;
            }
            add_vpacket_collection_to_histogram(&x);
            if (out->vlog_nus && x.vp_n > 0) {
                int64_t base = atomic_fetch_add(w->vlog_n, x.vp_n);
                for (int64_t j = 0; j < x.vp_n && base + j < out->vlog_capacity; j++) {
                    out->vlog_nus[base + j] = x.vp_nu[j];
                    out->vlog_energies[base + j] = x.vp_energy[j];
                    out->vlog_initial_mus[base + j] = x.vp_mu[j];
                    out->vlog_initial_rs[base + j] = x.vp_r[j];
                    out->vlog_packet_index[base + j] = i;
                }
            }
        }
    }
    free(x.vp_nu); free(x.vp_energy); free(x.vp_mu); free(x.vp_r);
    free(x.chi_bf_contributions); free(x.x_sect_bfs); free(x.current_continua);
    return NULL;
}

typedef struct {
    accum_t *accs;
    int nthreads;
    tardis_oracle_outputs *out;
    size_t lo, hi;
} reduce_t;

static void *reduce_main(void *arg)
{
    reduce_t *r = (reduce_t *)arg;
    for (size_t k = r->lo; k < r->hi; k++) {
        double jb = 0.0, ed = 0.0;
        for (int t = 0; t < r->nthreads; t++) { jb += r->accs[t].j_blue[k]; ed += r->accs[t].edotlu[k]; }
        r->out->j_blue[k] = jb;
        r->out->edotlu[k] = ed;
    }
    return NULL;
}

/* modes/montecarlo_transport.py:239-373.  nthreads plays the role of
 * numba.set_num_threads (modes/classic/solver.py:196): per-thread estimator
 * copies, summed serially in thread order afterwards (:356-360).  With
 * nthreads == 1 packets are processed in index order like the reference. */
int tardis_oracle_run(const tardis_oracle_model *m, const tardis_oracle_config *c,
                      const tardis_oracle_packets *pk, tardis_oracle_outputs *out, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    accum_t *accs = (accum_t *)malloc(sizeof(accum_t) * nthreads);
    worker_t *ws = (worker_t *)malloc(sizeof(worker_t) * nthreads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    atomic_llong next_packet = 0, vlog_n = 0;
    /* per-thread tables (the reference's layout) up to TARDIS_ORACLE_PRIVATE_MAX threads (default 8), one shared table
     * with atomic adds above that; bench.py calibrates both layouts and several thread counts and keeps the fastest */
    int private_max = 8;
    {
        const char *e = getenv("TARDIS_ORACLE_PRIVATE_MAX");
        if (e && *e) private_max = atoi(e);
    }
    const int shared = nthreads > private_max;
    for (int t = 0; t < nthreads; t++) {
        accum_alloc(&accs[t], m, c, (shared && t > 0) ? &accs[0] : NULL);
        if (shared) accs[t].shared_line_estimators = 1;
        ws[t].m = m; ws[t].c = c; ws[t].pk = pk; ws[t].out = out; ws[t].acc = &accs[t];
        ws[t].next_packet = &next_packet; ws[t].vlog_n = &vlog_n;
    }
    if (nthreads == 1) {
        worker_main(&ws[0]);
    } else {
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker_main, &ws[t]);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }

    /* reduce of thread estimators, modes/montecarlo_transport.py:356-360.  Same summation order per
     * element as the reference's serial loop (thread 0, 1, 2, ...), but the element range is split across
     * the host threads so that the CPU baseline is not dominated by a serial pass over
     * nthreads x 2 x L x S doubles. */
    int error = 0;
    memset(out->j, 0, sizeof(double) * m->n_shells);
    memset(out->nu_bar, 0, sizeof(double) * m->n_shells);
    if (out->vhist) memset(out->vhist, 0, sizeof(double) * c->n_grid);
    memset(&out->counters, 0, sizeof(out->counters));
    if (c->continuum_processes_enabled && out->photo_ion_estimator) {
        size_t nc = (size_t)m->n_continua * m->n_shells;
        memset(out->photo_ion_estimator, 0, nc * sizeof(double));
        memset(out->stim_recomb_estimator, 0, nc * sizeof(double));
        memset(out->bf_heating_estimator, 0, nc * sizeof(double));
        memset(out->stim_recomb_cooling_estimator, 0, nc * sizeof(double));
        memset(out->photo_ion_estimator_statistics, 0, nc * sizeof(int64_t));
        memset(out->ff_heating_estimator, 0, m->n_shells * sizeof(double));
    }
    {
        reduce_t *rs = (reduce_t *)malloc(sizeof(reduce_t) * nthreads);
        size_t ls = (size_t)m->n_lines * m->n_shells;
        for (int t = 0; t < nthreads; t++) {
            rs[t].accs = accs; rs[t].nthreads = shared ? 1 : nthreads; rs[t].out = out;
            rs[t].lo = ls * (size_t)t / (size_t)nthreads; rs[t].hi = ls * (size_t)(t + 1) / (size_t)nthreads;
        }
        if (nthreads == 1) {
            reduce_main(&rs[0]);
        } else {
            for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, reduce_main, &rs[t]);
            for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
        }
        free(rs);
    }
    for (int t = 0; t < nthreads; t++) {
        accum_t *a = &accs[t];
        if (a->error && !error) error = a->error;
        for (int64_t s = 0; s < m->n_shells; s++) { out->j[s] += a->j[s]; out->nu_bar[s] += a->nu_bar[s]; }
        if (out->vhist) for (int64_t k = 0; k < c->n_grid; k++) out->vhist[k] += a->vhist[k];
        out->counters.n_line_steps += a->cnt.n_line_steps;
        out->counters.n_boundary_events += a->cnt.n_boundary_events;
        out->counters.n_line_events += a->cnt.n_line_events;
        out->counters.n_escat_events += a->cnt.n_escat_events;
        out->counters.n_rng_draws += a->cnt.n_rng_draws;
        out->counters.n_macro_jumps += a->cnt.n_macro_jumps;
        out->counters.n_macro_scanned += a->cnt.n_macro_scanned;
        out->counters.n_vpackets += a->cnt.n_vpackets;
        out->counters.n_vpacket_line_steps += a->cnt.n_vpacket_line_steps;
        out->counters.n_continuum_events += a->cnt.n_continuum_events;
        out->counters.n_bf_estimator_updates += a->cnt.n_bf_estimator_updates;
        if (c->continuum_processes_enabled && out->photo_ion_estimator) {
            size_t nc = (size_t)m->n_continua * m->n_shells;
            for (size_t k = 0; k < nc; k++) {
                out->photo_ion_estimator[k] += a->photo_ion[k];
                out->stim_recomb_estimator[k] += a->stim_recomb[k];
                out->bf_heating_estimator[k] += a->bf_heating[k];
                out->stim_recomb_cooling_estimator[k] += a->stim_recomb_cooling[k];
                out->photo_ion_estimator_statistics[k] += a->photo_ion_stats[k];
            }
            for (int64_t sh = 0; sh < m->n_shells; sh++) out->ff_heating_estimator[sh] += a->ff_heating[sh];
        }
    }
    for (int t = nthreads - 1; t >= 0; t--) accum_free(&accs[t], !shared || t == 0);
    free(accs); free(ws); free(th);
    out->vlog_count = (int64_t)vlog_n;
    return error;
}

/* ---- single-function probes used by the known-answer tests ---- */
double tardis_oracle_distance_boundary(double r, double mu, double r_inner, double r_outer, int64_t *delta_shell)
{
    return calculate_distance_boundary(r, mu, r_inner, r_outer, delta_shell);
}
double tardis_oracle_distance_line(double r, double mu, double nu, double comov_nu, int is_last_line, double nu_line,
                                   double time_explosion, int full_rel, int *error)
{
    *error = 0;
    return calculate_distance_line(r, mu, nu, comov_nu, is_last_line, nu_line, time_explosion, full_rel, error);
}
double tardis_oracle_doppler_factor(double velocity, double mu, int full_rel) { return get_doppler_factor(velocity, mu, full_rel); }
double tardis_oracle_inverse_doppler_factor(double velocity, double mu, int full_rel) { return get_inverse_doppler_factor(velocity, mu, full_rel); }
