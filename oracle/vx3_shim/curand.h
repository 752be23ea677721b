// TEST INFRASTRUCTURE: host stand-in for cuRAND as VX3.cuh uses it (random(max, seed): one draw from a freshly seeded state;
// the reference uses it to pick which link to test for divergence).
#pragma once
struct curandState_t {
    unsigned long long s;
};
inline void curand_init(unsigned long long seed, unsigned long long, unsigned long long, curandState_t *st) { st->s = seed * 6364136223846793005ULL + 1442695040888963407ULL; }
inline unsigned curand(curandState_t *st) {
    st->s = st->s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (unsigned)(st->s >> 33);
}
