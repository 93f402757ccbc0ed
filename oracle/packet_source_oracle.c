/* packet_source_oracle.c -- CPU restatement of the reference's black-body packet source.  TEST INFRASTRUCTURE ONLY:
 * nothing under tardis_b200/ may use it (see tardis_oracle.c).
 *
 * Follows, statement by statement and in the reference's own SEQUENTIAL order,
 *   BasePacketSource.create_packets           tardis/transport/montecarlo/packet_source/base.py:195-253
 *   BlackBodySimpleSource.create_packet_radii / _nus / _mus / _energies   .../black_body.py:122-220
 * and the numpy arithmetic they run on (third-party dependency, numpy 2.x, not vendored under /root/reference):
 *   SeedSequence.mix_entropy / generate_state   numpy/random/bit_generator.pyx
 *   pcg64 (XSL-RR 128/64), pcg64_next32         numpy/random/src/pcg64/pcg64.h
 *   random_standard_uniform (53-bit doubles)    numpy/random/src/distributions/distributions.c
 *   Generator.choice -> integers -> random_bounded_uint64_fill -> buffered_bounded_lemire_uint32   (same file)
 * Pinned against numpy itself (tests/test_packet_source.py: numpy is present wherever the tests run) and against the
 * restated source in tardis_b200/synthetic.py::make_packets. */
#include <math.h>
#include <stdint.h>

typedef unsigned __int128 u128;

typedef struct { u128 state, inc; int has_uint32; uint32_t uinteger; } pcg64_t;

static const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
static const uint32_t MIX_MULT_L = 0xca01f9ddu, MIX_MULT_R = 0x4973f715u;

static uint32_t hashmix(uint32_t value, uint32_t *hash_const) {
    value ^= *hash_const; *hash_const *= MULT_A; value *= *hash_const; value ^= value >> 16;
    return value;
}
static uint32_t mix(uint32_t x, uint32_t y) {
    uint32_t result = MIX_MULT_L * x - MIX_MULT_R * y;
    result ^= result >> 16;
    return result;
}

static void pcg_step(pcg64_t *g) {
    const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | (u128)0x4385DF649FCCF645ull;
    g->state = g->state * mult + g->inc;
}

/* np.random.default_rng(seed) for 0 <= seed < 2^64 */
static void default_rng(pcg64_t *g, uint64_t seed) {
    uint32_t entropy[2]; int n_ent = 0;
    uint32_t pool[4], state32[8], hash_const = INIT_A;
    if (seed == 0) entropy[n_ent++] = 0;
    for (; seed != 0; seed >>= 32) entropy[n_ent++] = (uint32_t)seed;
    for (int i = 0; i < 4; i++) pool[i] = hashmix(i < n_ent ? entropy[i] : 0, &hash_const);
    for (int i_src = 0; i_src < 4; i_src++)
        for (int i_dst = 0; i_dst < 4; i_dst++)
            if (i_src != i_dst) pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src], &hash_const));
    hash_const = INIT_B;
    for (int i_dst = 0; i_dst < 8; i_dst++) {
        uint32_t data_val = pool[i_dst % 4];
        data_val ^= hash_const; hash_const *= MULT_B; data_val *= hash_const; data_val ^= data_val >> 16;
        state32[i_dst] = data_val;
    }
    const uint64_t s0 = state32[0] | ((uint64_t)state32[1] << 32), s1 = state32[2] | ((uint64_t)state32[3] << 32);
    const uint64_t s2 = state32[4] | ((uint64_t)state32[5] << 32), s3 = state32[6] | ((uint64_t)state32[7] << 32);
    const u128 initstate = ((u128)s0 << 64) | s1, initseq = ((u128)s2 << 64) | s3;
    g->state = 0; g->inc = (initseq << 1) | 1;
    pcg_step(g); g->state += initstate; pcg_step(g);
    g->has_uint32 = 0; g->uinteger = 0;
}

static uint64_t next_uint64(pcg64_t *g) {
    pcg_step(g);
    const uint64_t x = (uint64_t)(g->state >> 64) ^ (uint64_t)g->state;
    const unsigned rot = (unsigned)(g->state >> 122);
    return (x >> rot) | (x << ((64 - rot) & 63));
}
static uint32_t next_uint32(pcg64_t *g) {
    if (g->has_uint32) { g->has_uint32 = 0; return g->uinteger; }
    const uint64_t next = next_uint64(g);
    g->has_uint32 = 1; g->uinteger = (uint32_t)(next >> 32);
    return (uint32_t)(next & 0xffffffffu);
}
static double next_double(pcg64_t *g) { return (double)(next_uint64(g) >> 11) * (1.0 / 9007199254740992.0); }

/* buffered_bounded_lemire_uint32 with rng = high - 1 - low, 0 < rng < 0xFFFFFFFF */
static uint32_t bounded_lemire_uint32(pcg64_t *g, uint32_t rng) {
    const uint32_t rng_excl = rng + 1;
    uint64_t m = (uint64_t)next_uint32(g) * rng_excl;
    uint32_t leftover = (uint32_t)(m & 0xffffffffu);
    if (leftover < rng_excl) {
        const uint32_t threshold = (0xffffffffu - rng) % rng_excl;
        while (leftover < threshold) {
            m = (uint64_t)next_uint32(g) * rng_excl;
            leftover = (uint32_t)(m & 0xffffffffu);
        }
    }
    return (uint32_t)(m >> 32);
}

/* create_packets(no_of_packets, seed_offset) with seed = base_seed + seed_offset.  xis_scratch: 5 * n doubles. */
/* beta < 0: BlackBodySimpleSource; beta >= 0: BlackBodySimpleSourceRelativistic (packet_source/black_body_relativistic.py:120-177) */
int tardis_oracle_create_packets_beta(uint64_t seed, int64_t n, uint32_t max_seed_val, double radius, double temperature, double k_boltzmann,
                                      double h_planck, const double *l_array, int64_t n_l, double l_coef, double *xis_scratch, double *radii,
                                      double *nus, double *mus, double *energies, int64_t *seeds, double beta);

int tardis_oracle_create_packets(uint64_t seed, int64_t n, uint32_t max_seed_val, double radius, double temperature, double k_boltzmann,
                                 double h_planck, const double *l_array, int64_t n_l, double l_coef, double *xis_scratch, double *radii,
                                 double *nus, double *mus, double *energies, int64_t *seeds) {
    return tardis_oracle_create_packets_beta(seed, n, max_seed_val, radius, temperature, k_boltzmann, h_planck, l_array, n_l, l_coef, xis_scratch,
                                             radii, nus, mus, energies, seeds, -1.0);
}

int tardis_oracle_create_packets_beta(uint64_t seed, int64_t n, uint32_t max_seed_val, double radius, double temperature, double k_boltzmann,
                                      double h_planck, const double *l_array, int64_t n_l, double l_coef, double *xis_scratch, double *radii,
                                      double *nus, double *mus, double *energies, int64_t *seeds, double beta) {
    if (n < 0 || max_seed_val < 2) return 1;  /* (population 2^32 would be numpy's unbuffered special case: not representable here) */
    pcg64_t g;
    default_rng(&g, seed);                                             /* base.py:226 self._reseed(base_seed + seed_offset) */
    for (int64_t i = 0; i < n; i++)                                    /* base.py:227-229 rng.choice(MAX_SEED_VAL, n, replace=True) */
        seeds[i] = (int64_t)bounded_lemire_uint32(&g, max_seed_val - 1);
    for (int64_t i = 0; i < n; i++) radii[i] = 1.0 * radius;           /* black_body.py:138 np.ones(n) * radius */
    for (int64_t k = 0; k < 5 * n; k++) xis_scratch[k] = next_double(&g);  /* black_body.py:173 rng.random((5, n)) */
    for (int64_t i = 0; i < n; i++) {
        const double v = xis_scratch[i] * l_coef;                      /* :175 l_array.searchsorted(xis[0] * l_coef) + 1.0 */
        int64_t lo = 0, hi = n_l;
        while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (l_array[mid] < v) lo = mid + 1; else hi = mid; }
        const double l = (double)lo + 1.0;
        double prod = xis_scratch[n + i];                              /* :176 np.prod(xis[1:], 0) */
        prod *= xis_scratch[2 * n + i]; prod *= xis_scratch[3 * n + i]; prod *= xis_scratch[4 * n + i];
        const double x = -log(prod) / l;                               /* :177 */
        nus[i] = x * (k_boltzmann * temperature) / h_planck;           /* :179 */
    }
    if (beta < 0.0) {
        for (int64_t i = 0; i < n; i++) mus[i] = sqrt(next_double(&g));    /* :198 np.sqrt(rng.random(n)) */
        for (int64_t i = 0; i < n; i++) energies[i] = 1.0 / (double)n;     /* :219 np.ones(n) / n */
    } else {
        for (int64_t i = 0; i < n; i++) {                                  /* black_body_relativistic.py:148-150 */
            const double z = next_double(&g);
            mus[i] = -beta + sqrt(beta * beta + 2 * beta * z + z);
        }
        const double gamma = 1.0 / sqrt(1 - beta * beta);                   /* :168-177 */
        const double factor = (2 * beta + 1) / (1 - beta * beta);
        for (int64_t i = 0; i < n; i++) energies[i] = 1.0 / (double)n * factor / gamma;
    }
    return 0;
}
