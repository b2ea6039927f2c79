// Host side of the seam, native: the support / load sampling of generate_sample()
// (OpenPyStruct_BeamOpt_training_SingleCore.py:133-160, MultiCore:137-162, GPU:141-170) drawn from a bit-exact replica
// of CPython's `random` -- MT19937 seeded like random.seed(int), randint -> randrange -> _randbelow_with_getrandbits,
// sample() with its two selection strategies, choice(), uniform() -- in the reference's call order, written straight
// into the arrays of the C ABI.  Same stream as the Python sampler (sampling.sample_case) for the same seed, so the
// datasets are identical; it exists because the Python loop costs 3.8 us per beam, the GPU 0.5 us.
// Integer / bit work: checked bit for bit against `random` (tests/test_host_logic.py).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/openpystruct_b200.h"

namespace {

struct MT {
    uint32_t s[624];
    int idx;
    void init_genrand(uint32_t seed)
    {
        s[0] = seed;
        for (int i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void init_by_array(const uint32_t *key, int len)
    {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = (624 > len ? 624 : len); k; --k) {
            s[i] = (s[i] ^ ((s[i - 1] ^ (s[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            if (++i >= 624) { s[0] = s[623]; i = 1; }
            if (++j >= len) j = 0;
        }
        for (int k = 623; k; --k) {
            s[i] = (s[i] ^ ((s[i - 1] ^ (s[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            if (++i >= 624) { s[0] = s[623]; i = 1; }
        }
        s[0] = 0x80000000u;
    }
    uint32_t u32()
    {
        if (idx >= 624) {
            int kk = 0;
            for (; kk < 624 - 397; ++kk) {
                const uint32_t y = (s[kk] & 0x80000000u) | (s[kk + 1] & 0x7fffffffu);
                s[kk] = s[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            for (; kk < 623; ++kk) {
                const uint32_t y = (s[kk] & 0x80000000u) | (s[kk + 1] & 0x7fffffffu);
                s[kk] = s[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            const uint32_t y = (s[623] & 0x80000000u) | (s[0] & 0x7fffffffu);
            s[623] = s[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            idx = 0;
        }
        uint32_t y = s[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    // random.random(): 53 bits from two words
    double random()
    {
        const uint32_t a = u32() >> 5, b = u32() >> 6;
        return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
    }
    // _randbelow_with_getrandbits(n), 0 < n < 2^32
    uint32_t below(uint32_t n)
    {
        int k = 0;
        for (uint32_t t = n; t; t >>= 1) ++k;                       // n.bit_length()
        uint32_t r = u32() >> (32 - k);                              // getrandbits(k), k <= 32
        while (r >= n) r = u32() >> (32 - k);
        return r;
    }
    int randint(int a, int b) { return a + (int)below((uint32_t)(b - a + 1)); }
    double uniform(double a, double b)
    {
        volatile double w = (b - a) * random();                     // (two roundings, like the interpreter's a + (b-a)*r)
        return a + w;
    }
};

}  // namespace

struct OpsSampler {
    MT mt;
};

extern "C" {

int ops_sampler_create(uint64_t seed, OpsSampler **out)
{
    if (!out) return OPS_E_BADARG;
    OpsSampler *s = (OpsSampler *)malloc(sizeof(OpsSampler));
    if (!s) return OPS_E_BADARG;
    // random.seed(int): the absolute value as little-endian 32-bit digits (at least one) through init_by_array
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    s->mt.init_by_array(key, key[1] ? 2 : 1);
    *out = s;
    return 0;
}

void ops_sampler_destroy(OpsSampler *s) { free(s); }

double ops_sampler_random(OpsSampler *s) { return s ? s->mt.random() : NAN; }

int ops_sampler_randint(OpsSampler *s, int a, int b) { return (s && b >= a) ? s->mt.randint(a, b) : 0; }

int ops_sampler_draw_cases(OpsSampler *sp, int64_t count, int32_t num_nodes, int32_t flag, double L,
                           const int32_t *roller_nodes, int32_t n_rollers, const int32_t *available_nodes,
                           int32_t n_available, double L_max, double L_min, int32_t N_rollers_max, int32_t M_forces_max,
                           double max_force, double min_force, int32_t num_cases, int32_t max_forces,
                           uint8_t *fixed_uy, int32_t *force_nodes, double *force_vals, double *L_out,
                           int32_t *roller_tags, int32_t roller_width, int32_t *force_tags, double *case_L)
{
    if (!sp || count < 0 || num_nodes < 2 || num_cases < 1 || count % num_cases != 0 || max_forces < 1 ||
        M_forces_max < 1 || M_forces_max > max_forces || !fixed_uy || !force_nodes || !force_vals || !L_out ||
        !roller_tags || !force_tags || !case_L || roller_width < 1)
        return OPS_E_BADARG;
    if (flag == 0 && (!roller_nodes || !available_nodes || n_rollers < 0 || n_rollers > roller_width || n_available < 0))
        return OPS_E_BADARG;
    if (flag != 0 && (N_rollers_max < 1 || N_rollers_max > roller_width)) return OPS_E_BADARG;
    MT &mt = sp->mt;
    int32_t *avail = (int32_t *)malloc(sizeof(int32_t) * (size_t)(num_nodes > n_available ? num_nodes : n_available + 1));
    int32_t *pool = (int32_t *)malloc(sizeof(int32_t) * (size_t)(num_nodes > n_available ? num_nodes : n_available + 1));
    if (!avail || !pool) { free(avail); free(pool); return OPS_E_BADARG; }
    for (int64_t i = 0; i < count; ++i) {
        const int64_t b = i / num_cases;
        const int c = (int)(i % num_cases);
        double Lb = L;
        int32_t *rt = roller_tags + i * roller_width;
        for (int j = 0; j < roller_width; ++j) rt[j] = 0;
        int na, nr;
        if (flag == 1) {
            Lb = L_min + mt.uniform(0.0, L_max);
            na = num_nodes - 2;
            for (int j = 0; j < na; ++j) avail[j] = j + 2;               // list(range(2, num_nodes))
            nr = 0;
            const int num_rollers = mt.randint(1, N_rollers_max);
            for (int j = 0; j < num_rollers; ++j) {
                if (na > 0) {
                    const uint32_t q = mt.below((uint32_t)na);          // choice(avail)
                    rt[nr++] = avail[q];
                    memmove(avail + q, avail + q + 1, sizeof(int32_t) * (size_t)(na - 1 - (int)q));   // avail.remove(r)
                    --na;
                }
            }
        } else {
            nr = n_rollers;
            for (int j = 0; j < nr; ++j) rt[j] = roller_nodes[j];
            na = n_available;
            memcpy(avail, available_nodes, sizeof(int32_t) * (size_t)na);
        }
        int k = mt.randint(1, M_forces_max);
        if (k > na) k = na;
        int32_t *ft = force_tags + i * max_forces;
        for (int j = 0; j < max_forces; ++j) ft[j] = 0;
        // random.sample(avail, k): pool strategy for n <= setsize, selection set otherwise (k <= 5 here: setsize 21)
        int setsize = 21;
        if (k > 5) setsize += (int)pow(4.0, ceil(log((double)(k * 3)) / log(4.0)));
        if (na <= setsize) {
            memcpy(pool, avail, sizeof(int32_t) * (size_t)na);
            for (int j = 0; j < k; ++j) {
                const uint32_t q = mt.below((uint32_t)(na - j));
                ft[j] = pool[q];
                pool[q] = pool[na - j - 1];
            }
        } else {
            uint32_t chosen[64];
            for (int j = 0; j < k; ++j) {
                uint32_t q;
                bool again;
                do {
                    q = mt.below((uint32_t)na);
                    again = false;
                    for (int t = 0; t < j; ++t) again = again || (chosen[t] == q);
                } while (again);
                chosen[j] = q;
                ft[j] = avail[q];
            }
        }
        double *fv = force_vals + i * max_forces;
        int32_t *fn = force_nodes + i * max_forces;
        for (int j = 0; j < max_forces; ++j) { fn[j] = -1; fv[j] = 0.0; }
        for (int j = 0; j < k; ++j) { fv[j] = mt.uniform(min_force, max_force); fn[j] = ft[j] - 1; }
        case_L[i] = Lb;
        if (c == 0) {                                                // supports and length of the beam: its first case
            L_out[b] = Lb;
            uint8_t *fx = fixed_uy + b * num_nodes;
            memset(fx, 0, (size_t)num_nodes);
            fx[0] = 1;
            for (int j = 0; j < nr; ++j) if (rt[j] >= 1 && rt[j] <= num_nodes) fx[rt[j] - 1] = 1;
        }
    }
    free(avail); free(pool);
    return 0;
}

}  // extern "C"
