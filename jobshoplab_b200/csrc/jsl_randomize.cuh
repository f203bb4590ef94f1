// jsl_randomize.cuh -- InstanceRandomizer (compiler/manipulators.py:99-150) for the variants of a batch.
//
// The reference manipulator keeps the shape of an instance and, per job, (1) shuffles the job's operations
// (np.random.permutation over the operation objects: machine and tool travel together) and (2) redraws every
// duration as random.randrange(min_duration, max_duration) where min / max are taken over the instance's
// current durations.  Its RNGs are Python's / numpy's global generators, so sample values are not part of
// any contract (no test of the reference pins them): the distribution is the contract.  Here one call draws
// a fresh instance for every variant of a batch from a counter-based generator,
//     word d of job j of variant v = Philox4x32-10(key = seed, ctr = (j, d / 4, v_lo, v_hi))[d % 4]
// (words 0..K-2: Fisher-Yates swaps from the back, words K-1..2K-2: durations), and recomputes what the
// compiler derives from the tables: per-job work, the Taillard lower bound (utils/utils.py:310-331) and
// max_allowed_time (utils/utils.py:334-352).
//
// randomize_variant() is plain sequential code shared by the CUDA kernel (one thread per variant) and the CPU
// debug build (tests/hostsim), so the two can be compared bit for bit.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define JSL_RHD __host__ __device__ inline
#else
#define JSL_RHD static inline
#endif

namespace jsl {

JSL_RHD uint32_t rnd_word(uint64_t seed, uint64_t variant, uint32_t job, uint32_t d) {
    uint32_t c0 = job, c1 = d >> 2, c2 = (uint32_t)variant, c3 = (uint32_t)(variant >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const uint32_t w = d & 3u;
    return w == 0 ? c0 : w == 1 ? c1 : w == 2 ? c2 : c3;
}
// uniform integer in [0, n) (multiply-high; bias < n / 2^32)
JSL_RHD uint32_t rnd_below(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

constexpr int RND_MAX_MACHINES = 32;  // engine limit; for m >= #machines the bound's term is a job length, never the max

// Randomise one variant in place.  op_mt: packed machine | tool << 16 words [N]; op_dur [N]; job_work [J];
// lb_mat [2] = (lower bound, max_allowed_time).  Durations are drawn from [min_d, max_d).
JSL_RHD void randomize_variant(uint64_t seed, uint64_t variant, int J, const int32_t* job_op_start, int min_d,
                               int max_d, int32_t* op_mt, int32_t* op_dur, int32_t* job_work, int32_t* lb_mat) {
    int32_t head[RND_MAX_MACHINES], tail[RND_MAX_MACHINES], load[RND_MAX_MACHINES];
    const uint32_t range = (uint32_t)(max_d - min_d);
    const int K0 = job_op_start[1] - job_op_start[0];
    for (int m = 0; m < K0 && m < RND_MAX_MACHINES; m++) { head[m] = 0x7fffffff; tail[m] = 0x7fffffff; load[m] = 0; }
    int64_t total = 0;
    int longest = 0;
    for (int j = 0; j < J; j++) {
        const int o0 = job_op_start[j], K = job_op_start[j + 1] - o0;
        // np.random.permutation: Fisher-Yates from the back
        for (int i = K - 1; i > 0; i--) {
            const uint32_t q = rnd_below(rnd_word(seed, variant, (uint32_t)j, (uint32_t)(K - 1 - i)), (uint32_t)(i + 1));
            const int32_t t = op_mt[o0 + i]; op_mt[o0 + i] = op_mt[o0 + (int)q]; op_mt[o0 + (int)q] = t;
        }
        int work = 0;
        for (int k = 0; k < K; k++) {
            const int d = min_d + (int)rnd_below(rnd_word(seed, variant, (uint32_t)j, (uint32_t)(K - 1 + k)), range);
            op_dur[o0 + k] = d;
            work += d;
        }
        job_work[j] = work;
        total += work;
        if (work > longest) longest = work;
        // Taillard bound pieces of this job: time before / after its first visit of every machine
        uint32_t seen[RND_MAX_MACHINES / 32];
        for (int w = 0; w < RND_MAX_MACHINES / 32; w++) seen[w] = 0;
        int before = 0;
        for (int k = 0; k < K; k++) {
            const int m = op_mt[o0 + k] & 0xffff, d = op_dur[o0 + k];
            if (m < K0 && m < RND_MAX_MACHINES) {
                load[m] += d;
                if (!((seen[m >> 5] >> (m & 31)) & 1u)) {
                    seen[m >> 5] |= 1u << (m & 31);
                    if (before < head[m]) head[m] = before;
                    const int after = work - before - d;
                    if (after < tail[m]) tail[m] = after;
                }
            }
            before += d;
        }
        for (int m = 0; m < K0 && m < RND_MAX_MACHINES; m++)
            if (!((seen[m >> 5] >> (m & 31)) & 1u)) {  // machine not visited: the whole job lies before it
                if (work < head[m]) head[m] = work;
                tail[m] = tail[m] < 0 ? tail[m] : 0;
            }
    }
    int best = longest;
    for (int m = 0; m < K0 && m < RND_MAX_MACHINES; m++) {
        const int v = head[m] + load[m] + tail[m];
        if (v > best) best = v;
    }
    lb_mat[0] = best;
    lb_mat[1] = (int32_t)total;
}

}  // namespace jsl
