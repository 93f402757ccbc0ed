// CPU build of tardis_b200/csrc/packet_source.cuh for the unit tests (tests/test_packet_source.py): the SAME functions
// the CUDA kernel runs, driven chunk by chunk the way the kernel drives them.  Test harness only.
#include "../tardis_b200/csrc/packet_source.cuh"
#include <algorithm>
#include <vector>

extern "C" int shim_create_packets(uint64_t seed, uint64_t n, uint32_t rng, double radius, double temperature_kb, double h_planck,
                                   const double *l_array, int n_l, double l_coef, uint64_t chunk, double *r, double *nu, double *mu,
                                   double *e, long long *seeds, uint64_t *n_rejected_out, int relativistic, double beta, double energy) {
    using namespace tbps;
    SourceParams P;
    P.origin = pcg64_from_seed(seed);
    P.n = n; P.rng_excl = rng + 1u; P.threshold = lemire_threshold(rng);
    std::vector<uint64_t> rej;
    for (int it = 0; it < 64; it++) {  // fixed point: rejected draws among the first n + R raw draws
        std::vector<uint64_t> found;
        const uint64_t n_raw = n + rej.size();
        for (uint64_t k0 = 0; k0 < n_raw; k0 += chunk) {
            uint64_t buf[4096];
            const uint64_t k1 = std::min(n_raw, k0 + chunk);
            const int c = scan_chunk(P.origin, k0, k1, P.rng_excl, P.threshold, buf, 4096);
            if (c > 4096) return 2;
            found.insert(found.end(), buf, buf + c);
        }
        if (found.size() == rej.size()) { rej = found; break; }
        rej = found;
    }
    std::sort(rej.begin(), rej.end());
    P.rejected = rej.data(); P.n_rej = (int)rej.size();
    P.dbl_start = (n + rej.size() + 1) / 2;
    P.l_array = l_array; P.n_l = n_l; P.l_coef = l_coef; P.k_b_t = temperature_kb; P.h_planck = h_planck;
    P.radius = radius; P.energy = n ? 1.0 / (double)n : 0.0;
    P.relativistic = relativistic; P.beta = beta;
    if (relativistic) P.energy = energy;
    for (uint64_t i0 = 0; i0 < n; i0 += chunk) fill_chunk(P, i0, std::min(n, i0 + chunk), r, nu, mu, e, seeds);
    if (n_rejected_out) *n_rejected_out = rej.size();
    return 0;
}

extern "C" void shim_advance(uint64_t seed, uint64_t delta, uint64_t *out4) {
    tbps::Pcg64 g = tbps::pcg64_from_seed(seed);
    tbps::pcg_advance(g, delta);
    out4[0] = (uint64_t)(g.state >> 64); out4[1] = (uint64_t)g.state; out4[2] = (uint64_t)(g.inc >> 64); out4[3] = (uint64_t)g.inc;
}
