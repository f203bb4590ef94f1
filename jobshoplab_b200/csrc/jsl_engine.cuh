// jsl_engine.cuh -- warp-per-env discrete-event engine of the JobShopLab env-step path (sm_100a).
//
// One warp owns one env.  The env's mutable state (EnvS) is staged in shared memory; control flow
// is warp-uniform (all 32 lanes execute the same scalar code on the same shared data) and every
// scan over jobs / machines / transports is lane-parallel through three primitives:
//     wp_ballot  -- predicate over i in [0,n)  -> bit mask          (__ballot_sync)
//     wp_min     -- min over i in [0,n)                             (__reduce_min_sync)
//     wp_for     -- independent per-element work + __syncwarp()
// Ordered buffers (FIFO/LIFO stores of the reference) are not stored as lists: every job carries the
// buffer it is in and a monotone sequence stamp; queue head/tail, fill level and position tests are
// ballots / min-reductions over (loc, seq).  possible_transitions is four bit masks (cached in the state),
// element k is recovered arithmetically in the reference's order.
//
// Template parameters of everything below: JW = lane words (ceil(max(J,T)/32)), ST = the instance has
// stochastic time objects, CL = specialisation level: 0 general; 1 "simple" (all buffers FLEX and unbounded, no
// setup times, no outages, deterministic: capacity / position / TimeDependency / outage / setup code is compiled
// out); 2 "classic" (simple + all travel times zero: no travel lookups, start / completion chains in one batch);
// 3 "queues" (deterministic, no setup times, no outages, but FIFO / LIFO / DUMMY buffers and finite capacities:
// only the outage and setup code is compiled out); 4 "features" (setup times, outages, stochastic times, but all
// buffers FLEX and unbounded: the queue / capacity / TimeDependency code is compiled out).  JSL_LVL_BUF / JSL_LVL_OUT
// name what a level keeps.
// The hot loop has to stay inside the instruction cache: one call site per phase (sm_step, run_env_op),
// rare paths out of line (JSL_DN), rarely used per-env constants in EnvAux (shared memory).
//
// The header also compiles as plain C++ (JSL_HOST_SIM) with the primitives written as loops.  That
// build exists only for CPU-side debugging of this very source (tests/hostsim) -- the product never
// loads it and has no CPU execution path.
//
// Reference semantics implemented here (file:line relative to /root/reference/jobshoplab/):
//   state_machine/core/state_machine/state.py:154-467, handler.py:55-1161, manipulate.py:45-444,
//   validate.py:30-161, core/transitions.py:82-186, state_machine/time_machines.py:61-221,
//   utils/state_machine_utils/{possible_transition_utils,buffer_type_utils,outage_utils}.py,
//   state_machine/middleware/middleware.py:264-443, env/factories/{actions,observations,rewards}.py,
//   env/env.py:427-454, utils/dipatching_benchmark.py:17-67, types/stochasticy_models.py:6-121.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/jobshoplab_b200.h"

#if defined(__CUDACC__) && !defined(JSL_HOST_SIM)
#define JSL_D __device__ __forceinline__
#define JSL_M __device__ __forceinline__
#define JSL_DN static __device__ __noinline__
#define JSL_LANE ((int)(threadIdx.x & 31))
#define JSL_SYNCWARP() __syncwarp()
#define JSL_POPC(x) __popc(x)
#define JSL_FFS(x) __ffs(x)
#else
#define JSL_D static inline
#define JSL_M inline
#define JSL_DN static inline
#define JSL_SYNCWARP() ((void)0)
#define JSL_POPC(x) __builtin_popcount(x)
#define JSL_FFS(x) __builtin_ffs((int)(x))
#endif

namespace jsl {

// chain fusion (apply_machine / apply_transport): always on in the kernels; the CPU debug build can switch the
// individual fusions off (bit mask, tests/hostsim: JSL_FUSE) to bisect against the oracle
#ifdef JSL_HOST_SIM
static int g_fuse = -1;
static long g_fused[5] = {0, 0, 0, 0, 0};  // how often each fusion fired (coverage of the fuzz / golden runs)
#define JSL_FUSE_ON(bit) ((g_fuse >> (bit)) & 1)
#define JSL_FUSED(bit) (g_fused[bit]++)
#else
#define JSL_FUSE_ON(bit) true
#define JSL_FUSED(bit) ((void)0)
#endif
#ifdef JSL_HOST_SIM
static long g_invariant_checks = 0, g_invariant_violations = 0;  // CPU debug build only
static long g_fast_start = 0, g_fast_complete = 0, g_general_batches = 0;
#endif

constexpr int NONE8 = 0xff;
constexpr int OCC_NOTIME = 0, OCC_TIME = 1, OCC_TD = 2;
constexpr int INT_BIG = 0x7fffffff;
constexpr int MAX_OUT = 2;  // outage slots per component

// ------------------------------------------------------------------------------------ instance
// Device-resident lowered instance set (one shape, n_inst variants of the op tables).
struct InstDev {
    int J, M, T, S, NT, N, NB, NP;
    int n_inst;
    const int32_t* job_op_start;  // [J+1]
    const int32_t* op_mt;         // [n_inst][N]  machine | tool << 16
    const int32_t* op_dur;        // [n_inst][N]
    const int32_t* job_work;      // [n_inst][J]  sum of the job's durations ("mwkr" key)
    const int32_t* lb_mat;        // [n_inst][2]  lower_bound, max_allowed_time
    const uint8_t* pre_type;      // [M]
    const int32_t* pre_cap;       // [M]
    const uint8_t* post_type;     // [M]
    const int32_t* post_cap;      // [M]
    const int32_t* setup;         // [M*NT*NT] or nullptr (= all zero)
    const uint8_t* sbuf_type;     // [S]
    const int32_t* sbuf_cap;      // [S]
    const uint8_t* sbuf_role;     // [S]
    const int32_t* buf_ref_id;    // [NB]
    const int32_t* travel;        // [NP*NP] or nullptr (= all zero)
    const uint8_t* agv_init_place;  // [T]
    const uint16_t* job_init_buf;   // [J]
    const int32_t* m_out_start;     // [M+1]
    const int32_t* m_out_freq;      // [n]
    const int32_t* m_out_dur;       // [n]
    const int32_t* t_out_freq;      // [n_t_out]
    const int32_t* t_out_dur;       // [n_t_out]
    const float* tr_comp;           // [M+T] float32(comp_idx / (M+T))  (observations.py:541)
    const float* tr_job;            // [J+1] float32(job / J), [J] = "no job"  (:536-539)
    const uint8_t* job_lex;         // [J] r-th job in string-id order
    const uint8_t* mach_lex;        // [M]
    int n_t_out;
    int out_buf;      // first OUTPUT-role buffer idx or -1
    int default_tool;
    int start_time;
    int flex_pre;     // every prebuffer is FLEX  -> no automatic machine starts
    int flex_post;    // every postbuffer and standalone buffer is FLEX -> "ready" needs no position test
    int caps_inf;     // no finite capacity anywhere
    int has_m_out, has_t_out;
    // stochastic time objects (stochasticy_models.py): one flat table over every time object of the
    // instance -- ops [0,N), setup, travel, machine-outage freq/dur, transport-outage freq/dur.
    // nullptr for deterministic instances (the static tables above are used directly).
    const uint8_t* tobj_kind;   // [n_tobj] JSL_TIME_*
    const int32_t* tobj_base;   // [n_tobj] base_time
    const float* tobj_param;    // [n_tobj] offset | std | scale
    const int32_t* tobj_init;   // [n_tobj] .time at compile time
    const int32_t* sample_table; // [rows][sample_depth]: value of a time object for numpy seed k
    const int32_t* sample_row;   // [n_tobj] row of each object, -1 = static
    int sample_depth;
    int n_tobj, off_setup, off_travel, off_mof, off_mod, off_tof, off_tod;
    int sb0;           // 3M+T: index of the first standalone buffer
    double neg_inv_n;  // -1.0 / N (rewards.py:135-138)
    uint32_t out_mask; // bit s set: standalone buffer s has the OUTPUT role (S <= 32)
};

struct CfgDev {
    int allow_early_transport, truncation_joker, truncation_active, obs_kind, record_ops;
    double sparse_bias, dense_bias, truncation_bias;
};

// ------------------------------------------------------------------------------------ state
// Mutable state of one env.  JW = ceil(max(J,T)/32) lane words; machines <= 32.
// The shared-memory working copy.  Its first IMAGE_BYTES bytes are the HBM image (step mode): everything that
// has to survive between launches, in one contiguous 16-byte-aligned run; behind it the TimeDependency fields
// (part of the image only in the levels that have ordered buffers) and the per-batch scratch of sm_step, which
// never leaves the SM.
template <int JW, bool ST = false, int CL = false>
struct alignas(16) EnvS {
    static constexpr int MJ = 32 * JW, MT = 32 * JW, MM = 32;
    // scalars (uniform)
    int32_t time;
    int32_t seq_ctr;
    int32_t n_possible;       // len(possible_transitions) after the last state-machine call
    int32_t skip;             // candidates dropped by "no-op with >1 candidates" since then
    int32_t joker;            // middleware.py:260
    int32_t step_noop, step_action;  // SubTimeStepper, middleware.py:45-124
    int32_t reward_noop;      // rewards.py:106 no_op_counter
    int32_t n_steps, noop_count;
    int32_t max_done_end;     // max end_time over DONE ops (state.py:270-275)
    int32_t status;           // JSL_ST_*
    uint8_t terminated, truncated, last_action_empty, obs_done;
    int32_t all_out;          // core_utils.is_done of the current state
    int32_t draw_ctr;         // StochasticTimeConfig.update() calls so far (tape position / Philox counter)
    int32_t lower_bound;      // of this env's instance variant, copied at reset (utils.py:310-331)
    double ret;               // sum of rewards
    double inv_mat;           // 1.0 / max_allowed_time (see obs_time_fraction)
    int32_t max_allowed_time; // utils.py:334-352, copied at reset
    int32_t pad_[3];
    // candidate masks of the state as of the last state-machine call (possible_transitions, unexpanded)
    uint32_t c_p1[JW], c_idle[JW], c_lr[JW], c_li[JW];
    // jobs
    int32_t j_start[MJ], j_end[MJ];  // of the PROCESSING op
    uint32_t j_seq[MJ];              // order stamp inside the current buffer
    uint16_t j_loc[MJ];              // buffer idx
    uint8_t j_k[MJ];                 // number of DONE ops == index of the next not-done op
    uint8_t j_proc[MJ];              // op j_k is PROCESSING
    uint8_t j_agv[MJ];               // transport whose transport_job is this job, or NONE8
    uint8_t j_mach[MJ];              // machine of op j_k (NONE8 when all ops are done)
    uint32_t j_exec[MJ];             // bit m: the job has a DONE op on machine m (observation row)
    // machines
    int32_t m_occ[MM];
    uint16_t m_done[MM];             // DONE ops on this machine (observation)
    uint8_t m_state[MM], m_tool[MM], m_job[MM];
    // transports
    int32_t t_occ[MT];               // time | blocking job (TimeDependency)
    uint16_t t_loc1[MT];
    uint8_t t_state[MT], t_job[MT], t_carry[MT], t_occkind[MT], t_lockind[MT], t_loc0[MT], t_loc2[MT];
    // ---- end of the image of the levels without ordered buffers (no TimeDependency can arise there)
    uint16_t t_tdbuf[MT];
    uint8_t t_tdcomp[MT], t_tdjob[MT];
    // ---- end of the image of the other levels; scratch: timed transitions of the current batch
    uint8_t x_mkind[MM], x_mjob[MM];
    uint8_t x_tkind[MT], x_tcomp[MT], x_tjob[MT];
};
#define JSL_LVL_BUF_(CL) ((CL) == 0 || (CL) == 3)
template <int JW, int CL>
struct Image {  // bytes of EnvS that travel to HBM
    static constexpr int BYTES = JSL_LVL_BUF_(CL) ? (int)offsetof(EnvS<JW>, x_mkind) : (int)offsetof(EnvS<JW>, t_tdbuf);
    static_assert(BYTES % 16 == 0, "the state image is copied in 16-byte units");
};

// Outage bookkeeping (types/state_types.py:178-197), only allocated for instances with outages.
template <int JW>
struct alignas(16) OutS {
    static constexpr int MT = 32 * JW, MM = 32;
    int32_t mo_a[MM * MAX_OUT], mo_b[MM * MAX_OUT];
    int32_t to_a[MT * MAX_OUT], to_b[MT * MAX_OUT];
    uint8_t mo_active[MM * MAX_OUT], to_active[MT * MAX_OUT];
};

// Rarely used per-env constants: kept in the env's shared-memory slot, not in registers.
struct alignas(16) EnvAux {
    const int32_t* op_mt;        // this env's variant of the op tables
    const int32_t* op_dur;
    const int32_t* job_work;     // [J] sum of each job's durations ("mwkr" key)
    int32_t* op_log;             // [N][2] or nullptr
    void* outs;                  // OutS<JW>* or nullptr
    int32_t* tcur;               // [n_tobj] current .time of every time object (stochastic instances) or nullptr
    uint16_t* tseed;             // [n_tobj] current numpy seed of every stochastic time object
    const int32_t* start_seeds;  // [n_tobj] recorded start seeds, or nullptr (drawn 0..30 from Philox)
    const int32_t* tape;         // replayed update() results, or nullptr (table mode)
    const uint8_t* action_tape;  // TAPE policy: this env's actions
    int64_t env;                 // index of the env inside its batch
    int32_t step_limit;          // rollout: stop once n_steps reaches this
    uint64_t seed, env_id;       // Philox key / stream of this env
    double neg_inv_n;            // -1.0 / N: the dense reward of a penalised step (rewards.py:135-138)
    int32_t tape_len;
    int32_t pad0_;
    const int32_t* lb_mat;       // {lower_bound, max_allowed_time} of this env's instance variant (read at reset)
    uint32_t act_block;          // RANDOM policy: 1 + index of the 32-step block whose actions are in act_bits (0 = none)
    uint32_t act_bits;
    int32_t pad;
    int32_t* leg_log;            // transport-task log of this env (see LegLog) or nullptr
    uint32_t loop_scratch[14];   // stochastic kernels: sm_step keeps its transition masks and loop guard here (see there)
    uint64_t bar;                // k_step: mbarrier of the bulk copy that brings the state image in
};

// per-call context: the few uniform values the hot loop keeps in registers
struct Ctx {
    const InstDev* I;
    const CfgDev* C;
    EnvAux* aux;
    int raise;              // JSL_ST_* of an escaping exception, 0 = none
    int now;
    bool log;               // aux->leg_log is set (derived from a kernel parameter: free to test)
};

#define JSL_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define JSL_LVL_BUF(CL) ((CL) == 0 || (CL) == 3)  // ordered / bounded buffers, TimeDependency, state-graph validation
#define JSL_LVL_OUT(CL) ((CL) == 0 || (CL) == 4)  // outages, setup times
#define JSL_RAISE(c, code)          \
    do {                            \
        if (!(c).raise) (c).raise = (code); \
        return;                     \
    } while (0)
#define JSL_RAISE_V(c, code, v)     \
    do {                            \
        if (!(c).raise) (c).raise = (code); \
        return (v);                 \
    } while (0)

// ------------------------------------------------------------------------------------ primitives
// position of the k-th (0-based) set bit of x, k < popc(x).  (__fns is emulated by a long software sequence.)
JSL_M int nth_bit(uint32_t x, int k) {
    if (k < 4) {  // the usual case: the agent looks at one of the first candidates
        for (; k > 0; k--) x &= x - 1;
        return JSL_FFS(x) - 1;
    }
    int pos = 0, c;
    c = JSL_POPC(x & 0xffffu); if (k >= c) { k -= c; pos += 16; x >>= 16; }
    c = JSL_POPC(x & 0xffu);   if (k >= c) { k -= c; pos += 8;  x >>= 8; }
    c = JSL_POPC(x & 0xfu);    if (k >= c) { k -= c; pos += 4;  x >>= 4; }
    c = JSL_POPC(x & 0x3u);    if (k >= c) { k -= c; pos += 2;  x >>= 2; }
    c = (int)(x & 1u);         if (k >= c) pos += 1;
    return pos;
}

template <int W>
struct Mask {
    uint32_t w[W];
    JSL_M bool any() const {
        uint32_t a = 0;
#pragma unroll
        for (int i = 0; i < W; i++) a |= w[i];
        return a != 0;
    }
    JSL_M int count() const {
        int c = 0;
#pragma unroll
        for (int i = 0; i < W; i++) c += JSL_POPC(w[i]);
        return c;
    }
    // (a single-word mask is never indexed dynamically: that would put it into local memory)
    JSL_M bool test(int i) const { return W == 1 ? ((w[0] >> (i & 31)) & 1u) : ((w[i >> 5] >> (i & 31)) & 1u); }
    JSL_M void set(int i) { if (W == 1) w[0] |= 1u << (i & 31); else w[i >> 5] |= 1u << (i & 31); }
    JSL_M void clear(int i) { if (W == 1) w[0] &= ~(1u << (i & 31)); else w[i >> 5] &= ~(1u << (i & 31)); }
    JSL_M void assign_only(int i) {
#pragma unroll
        for (int k = 0; k < W; k++) w[k] = 0;
        set(i);
    }
    JSL_M uint32_t word_of(int i) const { return W == 1 ? w[0] : w[i >> 5]; }
    JSL_M int first() const {  // -1 if empty
#pragma unroll
        for (int i = 0; i < W; i++)
            if (w[i]) return i * 32 + JSL_FFS(w[i]) - 1;
        return -1;
    }
    JSL_M int pop_first() {
#pragma unroll
        for (int i = 0; i < W; i++)
            if (w[i]) {
                int b = JSL_FFS(w[i]) - 1;
                w[i] &= w[i] - 1;
                return i * 32 + b;
            }
        return -1;
    }
    JSL_M int nth(int k) const {  // k-th set bit (0-based), -1 if fewer
#pragma unroll
        for (int i = 0; i < W; i++) {
            int c = JSL_POPC(w[i]);
            if (k < c) return i * 32 + nth_bit(w[i], k);
            k -= c;
        }
        return -1;
    }
};

#if defined(__CUDACC__) && !defined(JSL_HOST_SIM)
template <int W, class F>
JSL_D Mask<W> wp_ballot(int n, F f) {
    Mask<W> m;
    const int lane = JSL_LANE;
#pragma unroll
    for (int w = 0; w < W; w++) {
        int i = w * 32 + lane;
        bool p = (i < n) && f(i);
        m.w[w] = __ballot_sync(0xffffffffu, p);
    }
    return m;
}
template <int W, class F>
JSL_D int wp_min(int n, F f) {  // f(i) -> int key (INT_BIG = not participating)
    int v = INT_BIG;
    const int lane = JSL_LANE;
#pragma unroll
    for (int w = 0; w < W; w++) {
        int i = w * 32 + lane;
        if (i < n) {
            int k = f(i);
            v = k < v ? k : v;
        }
    }
    return __reduce_min_sync(0xffffffffu, v);
}
template <int W, class F>
JSL_D uint32_t wp_or(int n, F f) {  // OR of f(i) -> uint32 over i in [0,n)
    uint32_t v = 0;
    const int lane = JSL_LANE;
#pragma unroll
    for (int w = 0; w < W; w++) {
        int i = w * 32 + lane;
        if (i < n) v |= f(i);
    }
    return __reduce_or_sync(0xffffffffu, v);
}
template <int W, class F>
JSL_D void wp_for(int n, F f) {
    const int lane = JSL_LANE;
#pragma unroll
    for (int w = 0; w < W; w++) {
        int i = w * 32 + lane;
        if (i < n) f(i);
    }
    __syncwarp();
}
template <class F>
JSL_D void wp_for_dyn(int n, F f) {
    for (int i = JSL_LANE; i < n; i += 32) f(i);
    __syncwarp();
}
#else
template <int W, class F>
JSL_D Mask<W> wp_ballot(int n, F f) {
    Mask<W> m;
    for (int w = 0; w < W; w++) m.w[w] = 0;
    for (int i = 0; i < n && i < 32 * W; i++)
        if (f(i)) m.w[i >> 5] |= 1u << (i & 31);
    return m;
}
template <int W, class F>
JSL_D int wp_min(int n, F f) {
    int v = INT_BIG;
    for (int i = 0; i < n && i < 32 * W; i++) {
        int k = f(i);
        v = k < v ? k : v;
    }
    return v;
}
template <int W, class F>
JSL_D uint32_t wp_or(int n, F f) {
    uint32_t v = 0;
    for (int i = 0; i < n && i < 32 * W; i++) v |= f(i);
    return v;
}
template <int W, class F>
JSL_D void wp_for(int n, F f) {
    for (int i = 0; i < n && i < 32 * W; i++) f(i);
}
template <class F>
JSL_D void wp_for_dyn(int n, F f) {
    for (int i = 0; i < n; i++) f(i);
}
#endif

// ------------------------------------------------------------------------------------ lookups
JSL_D int buf_type(const InstDev& I, int b) {
    if (b < 3 * I.M) {
        int m = b / 3, r = b - 3 * m;
        return r == 0 ? I.pre_type[m] : r == 1 ? I.post_type[m] : JSL_BUF_FLEX;
    }
    if (b < 3 * I.M + I.T) return JSL_BUF_FLEX;
    return I.sbuf_type[b - 3 * I.M - I.T];
}
JSL_D int buf_cap(const InstDev& I, int b) {
    if (b < 3 * I.M) {
        int m = b / 3, r = b - 3 * m;
        return r == 0 ? I.pre_cap[m] : r == 1 ? I.post_cap[m] : 1;
    }
    if (b < 3 * I.M + I.T) return 1;
    return I.sbuf_cap[b - 3 * I.M - I.T];
}
JSL_D bool is_standalone(const InstDev& I, int b) { return b >= I.sb0; }
JSL_D bool is_agv_buf(const InstDev& I, int b) { return b >= 3 * I.M && b < 3 * I.M + I.T; }
JSL_D bool is_output(const InstDev& I, int b) { return b >= I.sb0 && ((I.out_mask >> (b - I.sb0)) & 1u); }
// place of a buffer: machine for machine buffers, M+s for standalone, -1 for an AGV buffer
JSL_D int place_of_buf(const InstDev& I, int b) {
    if (b < 3 * I.M) return b / 3;
    if (b >= 3 * I.M + I.T) return I.M + (b - 3 * I.M - I.T);
    return -1;
}
JSL_D int job_nops(const InstDev& I, int j) { return I.job_op_start[j + 1] - I.job_op_start[j]; }

// ------------------------------------------------------------------------------------ buffers as (loc, seq)
template <int JW, bool ST, int CL>
JSL_DN int buf_count(const EnvS<JW, ST, CL>& s, const InstDev& I, int b) {
    return wp_ballot<JW>(I.J, [&](int j) { return s.j_loc[j] == b; }).count();
}

// buffer_type_utils.py:236-262 get_next_job_from_buffer; -1 = None
template <int JW, bool ST, int CL>
JSL_DN int next_job_from_buffer(const EnvS<JW, ST, CL>& s, const InstDev& I, int b) {
    int ty = buf_type(I, b);
    if (ty == JSL_BUF_FLEX) return -1;
    // seq stamps are unique per env, so the stamp identifies the job
    if (ty == JSL_BUF_LIFO) {
        int k = wp_min<JW>(I.J, [&](int j) { return s.j_loc[j] == b ? (int)(0x7ffffffe - s.j_seq[j]) : INT_BIG; });
        if (k == INT_BIG) return -1;
        uint32_t sq = (uint32_t)(0x7ffffffe - k);
        return wp_ballot<JW>(I.J, [&](int j) { return s.j_loc[j] == b && s.j_seq[j] == sq; }).first();
    }
    int k = wp_min<JW>(I.J, [&](int j) { return s.j_loc[j] == b ? (int)s.j_seq[j] : INT_BIG; });
    if (k == INT_BIG) return -1;
    return wp_ballot<JW>(I.J, [&](int j) { return s.j_loc[j] == b && (int)s.j_seq[j] == k; }).first();
}

// buffer_type_utils.py:328-353 is_job_ready_for_pickup_from_postbuffer (uniform j)
template <int JW, bool ST, int CL>
JSL_DN bool ready_for_pickup(const EnvS<JW, ST, CL>& s, const InstDev& I, int j) {
    int b = s.j_loc[j];
    bool right = is_standalone(I, b) || (b < 3 * I.M && (b % 3) == 1);
    if (!right) return false;
    int ty = buf_type(I, b);
    if (ty == JSL_BUF_FLEX) return true;
    uint32_t sq = s.j_seq[j];
    if (ty == JSL_BUF_LIFO)
        return !wp_ballot<JW>(I.J, [&](int q) { return s.j_loc[q] == b && s.j_seq[q] > sq; }).any();
    return !wp_ballot<JW>(I.J, [&](int q) { return s.j_loc[q] == b && s.j_seq[q] < sq; }).any();
}
// same test evaluated by one lane for its own job (serial scan; only used when early transport is off)
template <int JW, bool ST, int CL>
JSL_D bool ready_for_pickup_lane(const EnvS<JW, ST, CL>& s, const InstDev& I, int j) {
    int b = s.j_loc[j];
    bool right = is_standalone(I, b) || (b < 3 * I.M && (b % 3) == 1);
    if (!right) return false;
    if (!JSL_LVL_BUF(CL)) return true;  // every buffer is FLEX
    int ty = buf_type(I, b);
    if (ty == JSL_BUF_FLEX) return true;
    uint32_t sq = s.j_seq[j];
    bool ok = true;
#pragma unroll 1
    for (int q = 0; q < I.J; q++)
        if (s.j_loc[q] == b && (ty == JSL_BUF_LIFO ? s.j_seq[q] > sq : s.j_seq[q] < sq)) ok = false;
    return ok;
}

// get_next_job_from_buffer (buffer_type_utils.py:236-262) evaluated by ONE lane with a serial scan over the jobs:
// every lane can look at a different buffer, so all queues of a batch are resolved in one pass.  -1 = None.
template <int JW, bool ST, int CL>
JSL_D int queue_next_lane(const EnvS<JW, ST, CL>& s, const InstDev& I, int b, int ty) {
    if (ty == JSL_BUF_FLEX) return -1;
    const bool lifo = ty == JSL_BUF_LIFO;
    int best = -1;
    uint32_t bs = 0;
#pragma unroll 1
    for (int j = 0; j < I.J; j++) {
        const uint32_t sq = s.j_seq[j];
        if (s.j_loc[j] == b && (best < 0 || (lifo ? sq > bs : sq < bs))) { best = j; bs = sq; }
    }
    return best;
}

// buffer_type_utils.py:77-105 put_in_buffer (multi-slot buffers)
template <int JW, bool ST, int CL>
JSL_D void put_in_buffer(EnvS<JW, ST, CL>& s, Ctx& c, int b, int j) {
    const InstDev& I = *c.I;
    if (JSL_UNLIKELY(JSL_LVL_BUF(CL) && !I.caps_inf)) {
        int cap = buf_cap(I, b);
        if (cap != JSL_CAP_INF && buf_count(s, I, b) >= cap) JSL_RAISE(c, JSL_ST_BUFFER_FULL);
    }
    s.j_loc[j] = (uint16_t)b;
    s.j_seq[j] = (uint32_t)s.seq_ctr;
    s.seq_ctr = s.seq_ctr + 1;
}

// ------------------------------------------------------------------------------------ possible transitions
template <int JW>
struct Poss {
    Mask<JW> p1;    // jobs with a possible machine start        (state.py:317-338)
    Mask<JW> idle;  // idle AGVs                                 (possible_transition_utils.py:121-139)
    Mask<JW> lr;    // lonely running jobs                       (:258-273)
    Mask<JW> li;    // lonely idle transportable jobs            (:264-273)
    JSL_M int n_lonely() const { return lr.count() + li.count(); }
    JSL_M int total() const { return p1.count() + idle.count() * n_lonely(); }
};

// machine of the next IDLE op of job j and whether it has one (lane-local)
template <int JW, bool ST, int CL>
JSL_D int next_idle_place(const EnvS<JW, ST, CL>& s, const Ctx& c, int j) {  // -1 = none -> output buffer
    const InstDev& I = *c.I;
    int k = s.j_k[j] + (s.j_proc[j] ? 1 : 0);
    if (k >= job_nops(I, j)) return -1;
    return c.aux->op_mt[I.job_op_start[j] + k] & 0xffff;
}

template <int JW, bool ST, int CL>
JSL_D Poss<JW> compute_poss(const EnvS<JW, ST, CL>& s, const Ctx& c) {
    const InstDev& I = *c.I;
    Poss<JW> p;
    // possible_transition_utils.py:68-118 is_action_possible
    p.p1 = wp_ballot<JW>(I.J, [&](int j) {
        int m = s.j_mach[j];
        return !s.j_proc[j] && m != NONE8 && s.j_loc[j] == 3 * m && s.m_state[m] == JSL_M_IDLE;
    });
    p.idle = wp_ballot<JW>(I.T, [&](int t) { return s.t_state[t] == JSL_T_IDLE; });
    const bool early = c.C->allow_early_transport != 0;
    p.lr = wp_ballot<JW>(I.J, [&](int j) {
        if (!s.j_proc[j] || s.j_agv[j] != NONE8) return false;
        return early || ready_for_pickup_lane(s, I, j);
    });
    // possible_transition_utils.py:167-213 is_transportable
    p.li = wp_ballot<JW>(I.J, [&](int j) {
        if (s.j_proc[j] || s.j_agv[j] != NONE8) return false;
        int m = s.j_mach[j];
        bool tr = (m == NONE8) ? !is_output(I, s.j_loc[j]) : (s.j_loc[j] != 3 * m);
        if (!tr) return false;
        return early || ready_for_pickup_lane(s, I, j);
    });
    return p;
}

struct Trans {
    int kind, comp, new_state, job;
};

// idx-th element of get_possible_transitions(state) (state.py:298-341 ordering)
template <int JW, bool ST, int CL>
JSL_D Trans nth_possible(const EnvS<JW, ST, CL>& s, const Ctx& c, const Poss<JW>& p, int idx) {
    Trans t;
    int n1 = p.p1.count();
    if (idx < n1) {
        int j = p.p1.nth(idx);
        t.kind = 0; t.comp = s.j_mach[j]; t.new_state = JSL_M_SETUP; t.job = j;
        return t;
    }
    int r = idx - n1, nl = p.n_lonely(), nr = p.lr.count();
    int a = r / nl, q = r - a * nl;
    t.kind = 1; t.comp = p.idle.nth(a); t.new_state = JSL_T_WORKING;
    t.job = q < nr ? p.lr.nth(q) : p.li.nth(q - nr);
    return t;
}

template <int JW, bool ST, int CL>
JSL_D Poss<JW> cached_poss(const EnvS<JW, ST, CL>& s) {
    Poss<JW> p;
#pragma unroll
    for (int w = 0; w < JW; w++) { p.p1.w[w] = s.c_p1[w]; p.idle.w[w] = s.c_idle[w]; p.lr.w[w] = s.c_lr[w]; p.li.w[w] = s.c_li[w]; }
    return p;
}
template <int JW, bool ST, int CL>
JSL_D void store_poss(EnvS<JW, ST, CL>& s, const Poss<JW>& p, bool clear) {
#pragma unroll
    for (int w = 0; w < JW; w++) {
        s.c_p1[w] = clear ? 0u : p.p1.w[w]; s.c_idle[w] = clear ? 0u : p.idle.w[w];
        s.c_lr[w] = clear ? 0u : p.lr.w[w]; s.c_li[w] = clear ? 0u : p.li.w[w];
    }
}
// the candidate the agent is looking at: possible_transitions[0] after `skip` dropped ones
template <int JW, bool ST, int CL>
JSL_D Trans current_candidate(const EnvS<JW, ST, CL>& s, const Ctx& c) {
    return nth_possible(s, c, cached_poss(s), s.skip);
}

// ------------------------------------------------------------------------------------ time objects
// .time of a time object (no resample)
template <bool ST>
JSL_D int op_duration(const Ctx& c, int o) { return ST ? c.aux->tcur[o] : c.aux->op_dur[o]; }
template <bool ST, int CL>
JSL_D int setup_time_cur(const Ctx& c, int i) {
    if (ST) return c.aux->tcur[c.I->off_setup + i];
    return (JSL_LVL_OUT(CL) && c.I->setup) ? c.I->setup[i] : 0;
}
template <bool ST, int CL>
JSL_D int travel_time_cur(const Ctx& c, int a, int b) {
    if (ST) return c.aux->tcur[c.I->off_travel + a * c.I->NP + b];
    return (CL != 2 && c.I->travel) ? c.I->travel[a * c.I->NP + b] : 0;
}

struct Rng4 { uint32_t x, y, z, w; };
JSL_D Rng4 philox4(uint64_t seed, uint64_t stream, uint32_t c0, uint32_t c1) {
    uint32_t a = c0, b = c1, cc = (uint32_t)stream, d = (uint32_t)(stream >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * a;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * cc;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ b ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ d ^ k1;
        b = (uint32_t)p1; d = (uint32_t)p0; a = n0; cc = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Rng4 o = {a, b, cc, d};
    return o;
}
JSL_D double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

// one draw of a stochastic time object in Philox mode: stochasticy_models.py:44-121, then max(0, .)
// (:15,:30).  Distribution-exact, not stream-compatible with numpy; uses sub-counter k
// of draw `n` of stream `env` for rejection loops.
JSL_DN int draw_time(uint64_t seed, uint64_t env, uint32_t n, int kind, int base, float param) {
    double v = (double)base;
    if (kind == JSL_TIME_UNIFORM) {
        Rng4 r = philox4(seed, env, n, 1u << 16);
        double lo = (double)base - (double)param, hi = (double)base + (double)param;
        v = trunc(lo + (hi - lo) * u01(r.x));
    } else if (kind == JSL_TIME_GAUSSIAN) {
        Rng4 r = philox4(seed, env, n, 1u << 16);
        double z = sqrt(-2.0 * log(u01(r.x))) * cos(6.283185307179586 * u01(r.y));
        v = trunc((double)base + (double)param * z);
    } else if (kind == JSL_TIME_POISSON) {
        const double lam = (double)base + 0.5;
        int k = 0;
        if (lam < 10.0) {  // inversion by sequential search
            Rng4 r = philox4(seed, env, n, 1u << 16);
            double u = u01(r.x), p = exp(-lam), cdf = p;
            while (u > cdf && k < 1000) { k++; p *= lam / (double)k; cdf += p; }
        } else {  // PTRS, Hoermann 1993
            const double slam = sqrt(lam), loglam = log(lam);
            const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
            const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
            for (uint32_t it = 0;; it++) {
                Rng4 r = philox4(seed, env, n, (1u << 16) + it);
                double U = u01(r.x) - 0.5, V = u01(r.y);
                double us = 0.5 - fabs(U);
                k = (int)floor((2.0 * a / us + b) * U + lam + 0.43);
                if (us >= 0.07 && V <= vr) break;
                if (k < 0 || (us < 0.013 && V > us)) { if (it > 64) { k = (int)lam; break; } continue; }
                if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + (double)k * loglam - lgamma((double)k + 1.0)) break;
                if (it > 64) { k = (int)lam; break; }
            }
        }
        v = (double)base + (double)k;
    } else if (kind == JSL_TIME_GAMMA) {  // gamma(shape = base_time, scale): Marsaglia & Tsang 2000
        double shape = (double)base, g = 0.0;
        if (shape > 0.0) {
            const bool boost = shape < 1.0;
            const double sh = boost ? shape + 1.0 : shape;
            const double dd = sh - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
            for (uint32_t it = 0;; it++) {
                Rng4 r = philox4(seed, env, n, (1u << 16) + it);
                double z = sqrt(-2.0 * log(u01(r.x))) * cos(6.283185307179586 * u01(r.y));
                double t = 1.0 + cc * z;
                if (t <= 0.0) { if (it > 64) { g = dd; break; } continue; }
                t = t * t * t;
                double u = u01(r.z);
                if (log(u) < 0.5 * z * z + dd - dd * t + dd * log(t) || it > 64) {
                    g = dd * t;
                    if (boost) g *= pow(u01(r.w), 1.0 / shape);
                    break;
                }
            }
        }
        v = rint((double)base + g * (double)param);  // Python round(): half to even
    }
    if (!(v > 0.0)) v = 0.0;
    if (v > 2.0e9) v = 2.0e9;
    return (int)v;
}

// StochasticTimeConfig.update() of time object `idx` (stochasticy_models.py:27-30): a new sample becomes
// its .time.  Out of line and free of Ctx so the deterministic hot path only pays a null test.
// Returns the new time; *raise is set when the replay tape runs out.
JSL_DN int stoch_update(const InstDev* I, int32_t* tcur, uint16_t* tseed, const int32_t* tape, int tape_len,
                        int32_t* draw_ctr, uint64_t seed, uint64_t env, int idx, int* raise) {
    const int kind = I->tobj_kind[idx];
    if (kind == JSL_TIME_STATIC) return tcur[idx];
    const int n = *draw_ctr;
    int v;
    if (tape) {
        if (n >= tape_len) { if (!*raise) *raise = JSL_ST_OTHER_INVALID; return 0; }
        v = tape[n];
    } else {
        // stochasticy_models.py:22-30: seed += 1; time = max(0, first variate of default_rng(seed))
        const int sd = (int)tseed[idx] + 1;
        tseed[idx] = (uint16_t)sd;
        const int row = I->sample_row ? I->sample_row[idx] : -1;
        if (row >= 0 && sd < I->sample_depth) v = I->sample_table[row * I->sample_depth + sd];
        else v = draw_time(seed, env, (uint32_t)n, kind, I->tobj_base[idx], I->tobj_param[idx]);
    }
    *draw_ctr = n + 1;
    tcur[idx] = v;
    return v;
}
template <bool ST>
JSL_D int tobj_update(Ctx& c, int32_t& draw_ctr, int idx, int static_value) {
    if (!ST) return static_value;
    const InstDev& I = *c.I;
    if (I.tobj_kind[idx] == JSL_TIME_STATIC) return c.aux->tcur[idx];
#ifndef JSL_NO_INLINE_TABLE
    // the usual case inline -- table mode, seed within the table: no call, nothing spilled around it
    if (c.aux->tape == nullptr && I.sample_row != nullptr) {
        const int row = I.sample_row[idx], sd = (int)c.aux->tseed[idx] + 1;
        if (row >= 0 && sd < I.sample_depth) {
            const int v = I.sample_table[row * I.sample_depth + sd];
            c.aux->tseed[idx] = (uint16_t)sd;
            c.aux->tcur[idx] = v;
            draw_ctr = draw_ctr + 1;
            return v;
        }
    }
#endif
    int raise = 0;
    int32_t ctr = draw_ctr;
    int v = stoch_update(c.I, c.aux->tcur, c.aux->tseed, c.aux->tape, c.aux->tape_len, &ctr, c.aux->seed, c.aux->env_id, idx, &raise);
    draw_ctr = ctr;
    if (raise && !c.raise) c.raise = raise;
    return v;
}

// .time of every time object as compiled.  Free-running mode: each stochastic object gets a numpy start seed
// (recorded, or uniform 0..30 like np.random.randint(0, 31), stochasticy_models.py:12) and its
// construction-time sample table[row][seed] (:15).  Executed by all lanes of the env's warp.
JSL_DN void stoch_init(const InstDev* I, int32_t* tcur, uint16_t* tseed, const int32_t* start_seeds,
                       const int32_t* tape, uint64_t seed, uint64_t env, const int32_t* op_dur) {
    const bool table_mode = tape == nullptr && I->sample_row != nullptr;
    wp_for_dyn(I->n_tobj, [&](int i) {
        // a static operation duration is the env's own variant's (batches of several instances share every other
        // time object; InstanceRandomizer makes durations deterministic, manipulators.py:133-137)
        int v = (i < I->N && I->tobj_kind[i] == JSL_TIME_STATIC) ? op_dur[i] : I->tobj_init[i];
        if (table_mode) {
            const int row = I->sample_row[i];
            if (row >= 0) {
                int sd;
                if (start_seeds) sd = start_seeds[i];
                else sd = (int)(philox4(seed, env, (uint32_t)i, 2u << 16).x % 31u);
                if (sd < 0) sd = 0;
                if (sd >= I->sample_depth) sd = I->sample_depth - 1;
                tseed[i] = (uint16_t)sd;
                v = I->sample_table[row * I->sample_depth + sd];  // the construction-time sample of that seed (:15)
            }
        }
        tcur[i] = v;
    });
}

// ------------------------------------------------------------------------------------ outages
// outage_utils.py:20-131: sample the outages of one component, return the occupied time
template <bool ST>
JSL_D int sample_outages(Ctx& c, int32_t& draw_ctr, int now, int n, const int32_t* freq, const int32_t* dur,
                         int freq_obj, int dur_obj, uint8_t* active, int32_t* a, int32_t* b) {
    int occupied = 0;
    for (int i = 0; i < n; i++) {
        if (active[i]) JSL_RAISE_V(c, JSL_ST_OTHER_INVALID, 0);
        int last = a[i] < 0 ? 0 : a[i];
        int elapsed = now - last;
        bool apply;
        int d = dur[i];
        if (ST) {
            if (c.I->tobj_kind[freq_obj + i] != JSL_TIME_STATIC) {
                apply = elapsed > c.aux->tcur[freq_obj + i];  // outage_utils.py:32-42 _sample_frq
                if (apply) tobj_update<ST>(c, draw_ctr, freq_obj + i, 0);
            } else {
                apply = freq[i] >= elapsed;  // sic, outage_utils.py:48
            }
            if (apply) d = tobj_update<ST>(c, draw_ctr, dur_obj + i, d);
            if (c.raise) return 0;
        } else {
            apply = freq[i] >= elapsed;
        }
        if (apply) { active[i] = 1; a[i] = now; b[i] = now + d; }
    }
    bool any = false;
    for (int i = 0; i < n; i++)
        if (active[i]) {
            int len = b[i] - a[i];
            if (!any || len > occupied) occupied = len;
            any = true;
        }
    return any ? occupied : 0;
}
JSL_D void release_outages(int n, uint8_t* active, int32_t* a, int32_t* b) {
    for (int i = 0; i < n; i++)
        if (active[i]) { active[i] = 0; a[i] = b[i]; b[i] = -1; }
}

// ------------------------------------------------------------------------------------ machine handlers
// validate.py:30-82.  returns false on a validation error (StateMachineResult.success = False)
template <int JW, bool ST, int CL>
JSL_D bool machine_valid(const EnvS<JW, ST, CL>& s, Ctx& c, int m, int to, int job) {
    int from = s.m_state[m];
    if (from == to) return false;
    bool ok = from == JSL_M_IDLE ? to == JSL_M_SETUP
              : from == JSL_M_SETUP ? to == JSL_M_WORKING
              : from == JSL_M_WORKING ? (to == JSL_M_OUTAGE)
                                      : to == JSL_M_IDLE;
    if (JSL_UNLIKELY(!ok)) return false;
    if (from == JSL_M_OUTAGE || from == JSL_M_WORKING) return true;
    if (job != NONE8) {
        if (s.j_mach[job] == NONE8) JSL_RAISE_V(c, JSL_ST_OTHER_INVALID, false);
        if (s.j_mach[job] != m) return false;
    }
    return true;
}

// handler.py:414-457 + manipulate.py:378-444
// Returns true when the SETUP -> WORKING transition that the next batch would create can be applied right away
// (chain fusion, see apply_machine): a static setup time of zero.
template <int JW, bool ST, int CL>
JSL_D bool machine_idle_to_setup(EnvS<JW, ST, CL>& s, Ctx& c, int m, int j) {
    const InstDev& I = *c.I;
    if (j == NONE8) JSL_RAISE_V(c, JSL_ST_OTHER_INVALID, false);
    if (s.j_loc[j] != 3 * m) JSL_RAISE_V(c, JSL_ST_OTHER_INVALID, false);
    int o = I.job_op_start[j] + s.j_k[j];
    int tool = (c.aux->op_mt[o] >> 16) & 0xffff;
    int setup = 0;
    bool drawn = false;
    if (JSL_UNLIKELY(JSL_LVL_OUT(CL) && (I.setup || ST))) {
        const int si = (m * I.NT + s.m_tool[m]) * I.NT + tool;
        setup = setup_time_cur<ST, CL>(c, si);
        if (ST) drawn = I.tobj_kind[I.off_setup + si] != JSL_TIME_STATIC;
        if (setup == JSL_NO_TIME && !drawn) JSL_RAISE_V(c, JSL_ST_OTHER_INVALID, false);
        setup = tobj_update<ST>(c, s.draw_ctr, I.off_setup + si, setup);
        if (JSL_UNLIKELY(c.raise)) return false;
    }
    if (s.m_job[m] != NONE8) JSL_RAISE_V(c, JSL_ST_BUFFER_FULL, false);
    s.j_proc[j] = 1; s.j_start[j] = c.now; s.j_end[j] = c.now + setup;
    s.j_loc[j] = (uint16_t)(3 * m + 2);
    s.m_job[m] = (uint8_t)j;
    s.m_state[m] = JSL_M_SETUP; s.m_occ[m] = c.now + setup; s.m_tool[m] = (uint8_t)tool;
    return setup == 0 && !drawn;
}
// handler.py:460-503 + manipulate.py:225-276
template <int JW, bool ST, int CL>
JSL_D void machine_setup_to_working(EnvS<JW, ST, CL>& s, Ctx& c, int m, int j) {
    const InstDev& I = *c.I;
    if (j == NONE8 || s.j_loc[j] != 3 * m + 2) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    const int o = I.job_op_start[j] + s.j_k[j];
    int dur = tobj_update<ST>(c, s.draw_ctr, o, c.aux->op_dur[o]);
    if (JSL_UNLIKELY(c.raise)) return;
    s.j_proc[j] = 1; s.j_start[j] = c.now; s.j_end[j] = c.now + dur;
    s.m_state[m] = JSL_M_WORKING; s.m_occ[m] = c.now + dur;
}
// handler.py:506-543 + manipulate.py:328-375
template <int JW, bool ST, int CL>
JSL_D void machine_working_to_outage(EnvS<JW, ST, CL>& s, Ctx& c, int m, int j) {
    const InstDev& I = *c.I;
    int occupied_for = 0;
    if (JSL_UNLIKELY(JSL_LVL_OUT(CL) && I.has_m_out)) {
        int b = I.m_out_start[m], n = I.m_out_start[m + 1] - b;
        OutS<JW>& o = *(OutS<JW>*)c.aux->outs;
        occupied_for = sample_outages<ST>(c, s.draw_ctr, c.now, n, I.m_out_freq + b, I.m_out_dur + b, I.off_mof + b,
                                      I.off_mod + b, o.mo_active + m * MAX_OUT, o.mo_a + m * MAX_OUT, o.mo_b + m * MAX_OUT);
        if (JSL_UNLIKELY(c.raise)) return;
    }
    s.m_state[m] = JSL_M_OUTAGE; s.m_occ[m] = c.now + occupied_for;
    if (j == NONE8 || !s.j_proc[j]) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    s.j_end[j] = c.now + occupied_for;
}
// handler.py:546-574 + manipulate.py:117-193
template <int JW, bool ST, int CL>
JSL_D void machine_outage_to_idle(EnvS<JW, ST, CL>& s, Ctx& c, int m) {
    const InstDev& I = *c.I;
    int j = s.m_job[m];
    if (j == NONE8 || !s.j_proc[j]) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    int k = s.j_k[j];
    if (c.aux->op_log) {
        int o = I.job_op_start[j] + k;
        c.aux->op_log[2 * o] = s.j_start[j];
        c.aux->op_log[2 * o + 1] = c.now;
    }
    s.j_k[j] = (uint8_t)(k + 1);
    s.j_proc[j] = 0;
    s.j_exec[j] |= 1u << (c.aux->op_mt[I.job_op_start[j] + k] & 31);
    s.j_mach[j] = (k + 1 < job_nops(I, j)) ? (uint8_t)(c.aux->op_mt[I.job_op_start[j] + k + 1] & 0xffff) : (uint8_t)NONE8;
    if (c.now > s.max_done_end) s.max_done_end = c.now;
    s.m_done[m] = (uint16_t)(s.m_done[m] + 1);
    s.m_job[m] = NONE8;
    put_in_buffer(s, c, 3 * m + 1, j);
    if (JSL_UNLIKELY(c.raise)) return;
    if (JSL_UNLIKELY(JSL_LVL_OUT(CL) && I.has_m_out)) {
        OutS<JW>& o = *(OutS<JW>*)c.aux->outs;
        release_outages(I.m_out_start[m + 1] - I.m_out_start[m], o.mo_active + m * MAX_OUT, o.mo_a + m * MAX_OUT,
                        o.mo_b + m * MAX_OUT);
    }
    s.m_state[m] = JSL_M_IDLE;
}

// process one machine transition (validate + handle); false = validation error
template <int JW, bool ST, int CL>
JSL_D bool apply_machine(EnvS<JW, ST, CL>& s, Ctx& c, int m, int to, int job, bool fuse) {
    // every machine transition is created from that machine's own current state (timed) or from the candidate
    // masks of the unchanged state (action); only the TimeDependency hand-me-downs of non-FLEX postbuffers
    // (handler.py:632) can put two transitions of one component into a batch, so the state-graph check
    // (validate.py:30-82) can only fail there
    if (JSL_LVL_BUF(CL) && !c.I->flex_post && !machine_valid(s, c, m, to, job)) return false;
    if (JSL_UNLIKELY(c.raise)) return true;
    // Chain fusion.  The reference applies the follow-up of a zero-duration state in the NEXT batch of the same `now`
    // (state.py:236-267: one loop iteration per batch).  A follow-up may be applied here, in the batch that creates it,
    // when (1) nothing that runs in between observes the intermediate state and (2) the component then has nothing more
    // to do at this `now` -- otherwise the rest of its same-time cascade would run one batch early and could overtake
    // another component's (the order of two drop-offs into one buffer is state).  The state every env.step ends in is
    // then the same, the loop runs fewer batches.  For machines:
    //   IDLE -> SETUP -> WORKING   a static setup time of zero and a static duration > 0.  In between the reference only
    //                              has end_time = now instead of now + duration, which an AGV reads as its waiting time
    //                              and re-reads one batch later.
    // (WORKING -> OUTAGE -> IDLE is NOT fused: the job would reach the postbuffer, and its AGV leave, a batch early.)
    // Never for the agent's own action (`fuse` false): the time machine and the teleports that follow it
    // (state.py:214-232) read the intermediate occupied_till / end_time.
    int from = s.m_state[m];
    if (from == JSL_M_IDLE) {
        bool chain = machine_idle_to_setup(s, c, m, job);
        if (!chain || !fuse || !JSL_FUSE_ON(0) || JSL_UNLIKELY(c.raise)) return true;
        const int o = c.I->job_op_start[job] + s.j_k[job];
        if (c.aux->op_dur[o] <= 0 || (ST && c.I->tobj_kind[o] != JSL_TIME_STATIC)) return true;
        from = JSL_M_SETUP;
        JSL_FUSED(0);
    }
    if (from == JSL_M_SETUP) machine_setup_to_working(s, c, m, job);
    else if (from == JSL_M_WORKING) machine_working_to_outage(s, c, m, job);
    else machine_outage_to_idle(s, c, m);
    return true;
}

// ------------------------------------------------------------------------------------ transport-task log
// What env.render() needs of the transports (rendering/gant_dashboard.py:299-359): one record per transport task
// -- the states an AGV passes through between IDLE->PICKUP and the drop-off share one `location` triple -- with
//   start = `time` of the first recorded state carrying the triple.  state.py:199/266 records a state after the
//           action's transitions (time unchanged) and after every batch of the loop once jump_to_event has run,
//           so a task created inside the loop starts at the post-jump time (LEG_PENDING until then);
//   end   = occupied_till of the last such state (the TRANSIT state's arrival time for a finished task);
//   job, route = transport_job and the triple.
// Layout (int32): [0] count, [1] overflow, [2 .. 2+MT) open record of each AGV (-1), then 5 per record:
// agv (< 0: not visible, see run_env_op), job, start, end, route = place | pickup buffer << 8 | drop place << 24.
// Only with cfg.record_ops.
constexpr int LEG_PENDING = -2, LEG_NO_TIME = -1;
JSL_D int leg_hdr(int MT) { return 2 + MT; }
// Compiled in only with -DJSL_LEG_LOG (libjsl_b200_log.so, used for batches created with record_ops): even a
// never-taken branch per transition costs the throughput kernels ~5 % of their instruction budget.
#ifdef JSL_LEG_LOG
template <int JW, bool ST, int CL>
JSL_D void leg_open(EnvS<JW, ST, CL>& s, Ctx& c, int t, int j, int start, int end) {
    if (!c.log) return;
    int32_t* g = c.aux->leg_log;
    constexpr int MT = EnvS<JW, ST, CL>::MT;
    const int cap = c.I->N + 2 * c.I->J + 8;
    const int idx = g[0];
    if (idx >= cap) { g[1] = 1; g[2 + t] = -1; return; }
    int32_t* e = g + leg_hdr(MT) + 5 * idx;
    e[0] = t; e[1] = j; e[2] = start; e[3] = end;
    e[4] = (int)s.t_loc0[t] | ((int)s.t_loc1[t] << 8) | ((int)s.t_loc2[t] << 24);
    g[2 + t] = idx;
    g[0] = idx + 1;
}
template <int JW, bool ST, int CL>
JSL_D void leg_end(EnvS<JW, ST, CL>& s, Ctx& c, int t, int end, bool close) {
    if (!c.log) return;
    int32_t* g = c.aux->leg_log;
    constexpr int MT = EnvS<JW, ST, CL>::MT;
    const int idx = g[2 + t];
    if (idx >= 0 && end != LEG_PENDING) g[leg_hdr(MT) + 5 * idx + 3] = end;
    if (close) g[2 + t] = -1;
}
// the state after the current batch is recorded at time `now`: tasks opened in it start there
template <int JW, bool ST, int CL>
JSL_D void leg_stamp(EnvS<JW, ST, CL>& s, Ctx& c) {
    if (!c.log) return;
    int32_t* g = c.aux->leg_log;
    constexpr int MT = EnvS<JW, ST, CL>::MT;
    const int now = c.now;
    JSL_SYNCWARP();
    wp_for<JW>(c.I->T, [&](int t) {
        const int idx = g[2 + t];
        if (idx >= 0 && g[leg_hdr(MT) + 5 * idx + 2] == LEG_PENDING) g[leg_hdr(MT) + 5 * idx + 2] = now;
    });
}
constexpr bool LEG_LOG_BUILD = true;
#else
template <int JW, bool ST, int CL>
JSL_D void leg_open(EnvS<JW, ST, CL>&, Ctx&, int, int, int, int) {}
template <int JW, bool ST, int CL>
JSL_D void leg_end(EnvS<JW, ST, CL>&, Ctx&, int, int, bool) {}
template <int JW, bool ST, int CL>
JSL_D void leg_stamp(EnvS<JW, ST, CL>&, Ctx&) {}
constexpr bool LEG_LOG_BUILD = false;
#endif

// ------------------------------------------------------------------------------------ transport handlers
// validate.py:85-118 + transitions.py:82-107,161-186
template <int JW, bool ST, int CL>
JSL_D bool transport_valid(const EnvS<JW, ST, CL>& s, int t, int to) {
    // bit (from*6 + to): idle->running, running->running|outage (same state only for WAITINGPICKUP), outage->idle
    constexpr uint64_t R = (1ull << JSL_T_WORKING) | (1ull << JSL_T_PICKUP) | (1ull << JSL_T_TRANSIT) | (1ull << JSL_T_WAITINGPICKUP);
    constexpr uint64_t RO = R | (1ull << JSL_T_OUTAGE);
    constexpr uint64_t TABLE = (R << (6 * JSL_T_IDLE)) | ((RO & ~(1ull << JSL_T_WORKING)) << (6 * JSL_T_WORKING)) |
                               ((RO & ~(1ull << JSL_T_PICKUP)) << (6 * JSL_T_PICKUP)) |
                               ((RO & ~(1ull << JSL_T_TRANSIT)) << (6 * JSL_T_TRANSIT)) |
                               ((1ull << JSL_T_IDLE) << (6 * JSL_T_OUTAGE)) | (RO << (6 * JSL_T_WAITINGPICKUP));
    return (TABLE >> (s.t_state[t] * 6 + to)) & 1ull;
}

JSL_D int output_place(Ctx& c) {
    if (c.I->out_buf < 0) JSL_RAISE_V(c, JSL_ST_OTHER_INVALID, 0);
    return c.I->M + (c.I->out_buf - 3 * c.I->M - c.I->T);
}

// handler.py:808-902 idle -> working (PICKUP)
template <int JW, bool ST, int CL>
JSL_D void agv_idle_to_working(EnvS<JW, ST, CL>& s, Ctx& c, int t, int j) {
    const InstDev& I = *c.I;
    if (j == NONE8) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    if (s.t_lockind[t] != 0) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    int target = next_idle_place(s, c, j);
    if (target < 0) target = output_place(c);
    if (JSL_UNLIKELY(c.raise)) return;
    int loc = s.j_loc[j];
    int src = place_of_buf(I, loc);
    if (src < 0) JSL_RAISE(c, JSL_ST_OTHER_INVALID);  // TransportConfigError
    int ttp = travel_time_cur<ST, CL>(c, s.t_loc0[t], src);    // no resample (handler.py:876-885)
    if (ttp == JSL_NO_TIME) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    s.t_occkind[t] = OCC_TIME; s.t_occ[t] = c.now + ttp;
    s.t_lockind[t] = 1; s.t_loc1[t] = (uint16_t)loc; s.t_loc2[t] = (uint8_t)target;
    s.t_state[t] = JSL_T_PICKUP;
    s.t_job[t] = (uint8_t)j;
    s.j_agv[j] = (uint8_t)t;
    leg_open(s, c, t, j, LEG_PENDING, c.now + ttp);
}

// handler.py:577-639 _get_waiting_time, :642-702, :989-1030
template <int JW, bool ST, int CL>
JSL_D void agv_to_waitingpickup(EnvS<JW, ST, CL>& s, Ctx& c, int t, int j) {
    const InstDev& I = *c.I;
    if (j == NONE8) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    int loc = s.j_loc[j];
    int kind = OCC_TIME, val = c.now;
    if (is_standalone(I, loc)) {
        // now
    } else if (is_agv_buf(I, loc)) {
        JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    } else if (loc % 3 == 1) {  // in the machine's postbuffer
        if (JSL_LVL_BUF(CL) && !ready_for_pickup(s, I, j)) {  // FLEX postbuffers: always ready
            int nxt = next_job_from_buffer(s, I, loc);
            if (nxt < 0) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
            int other = s.j_agv[nxt];
            if (other == NONE8) {
                kind = OCC_TD; val = nxt;
                s.t_tdbuf[t] = (uint16_t)loc; s.t_tdcomp[t] = (uint8_t)t; s.t_tdjob[t] = (uint8_t)j;
            } else {  // the other transport's occupied_till, whatever it is
                kind = s.t_occkind[other]; val = s.t_occ[other];
                if (kind == OCC_TD) {
                    s.t_tdbuf[t] = s.t_tdbuf[other]; s.t_tdcomp[t] = s.t_tdcomp[other]; s.t_tdjob[t] = s.t_tdjob[other];
                }
            }
        }
    } else {  // prebuffer or machine buffer: wait for the processing op
        if (!s.j_proc[j]) JSL_RAISE(c, JSL_ST_OTHER_INVALID);  // MissingProcessingOperationError
        val = s.j_end[j];
    }
    s.t_occkind[t] = (uint8_t)kind; s.t_occ[t] = val;
    s.t_state[t] = JSL_T_WAITINGPICKUP;
    leg_end(s, c, t, kind == OCC_TIME ? val : LEG_NO_TIME, false);
}

// handler.py:705-805 (waiting)pickup -> transit
template <int JW, bool ST, int CL>
JSL_D void agv_to_transit(EnvS<JW, ST, CL>& s, Ctx& c, int t, int j) {
    const InstDev& I = *c.I;
    if (j == NONE8) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    int loc = s.j_loc[j];
    int src = place_of_buf(I, loc);
    if (src < 0) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    int dst;
    int nip = next_idle_place(s, c, j);
    if (nip < 0) dst = output_place(c);
    else dst = s.j_mach[j];  // get_next_not_done_operation(job).machine_id
    if (JSL_UNLIKELY(c.raise)) return;
    if (!(src < I.M || dst < I.M)) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    int travel = travel_time_cur<ST, CL>(c, src, dst);
    const int tobj = I.off_travel + src * I.NP + dst;
    if (travel == JSL_NO_TIME && !(ST && I.tobj_kind[tobj] != JSL_TIME_STATIC)) JSL_RAISE(c, JSL_ST_OTHER_INVALID);  // TravelTimeError
    travel = tobj_update<ST>(c, s.draw_ctr, tobj, travel);  // resampled (time_utils.py:15)
    if (JSL_UNLIKELY(c.raise)) return;
    if (s.t_carry[t] != NONE8) JSL_RAISE(c, JSL_ST_BUFFER_FULL);
    if (loc % 3 == 2 && loc < 3 * I.M) s.m_job[loc / 3] = NONE8;  // lifted out of a machine buffer
    s.j_loc[j] = (uint16_t)(3 * I.M + t);
    s.t_carry[t] = (uint8_t)j;
    s.t_state[t] = JSL_T_TRANSIT;
    s.t_occkind[t] = OCC_TIME; s.t_occ[t] = c.now + travel;
    leg_end(s, c, t, c.now + travel, false);
}

// handler.py:905-958 + manipulate.py:45-114 transit -> outage (drop-off)
template <int JW, bool ST, int CL>
JSL_D void agv_transit_to_outage(EnvS<JW, ST, CL>& s, Ctx& c, int t, int j) {
    const InstDev& I = *c.I;
    if (j == NONE8) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    if (s.t_lockind[t] != 1) JSL_RAISE(c, JSL_ST_OTHER_INVALID);
    if (s.t_carry[t] != j) JSL_RAISE(c, JSL_ST_OTHER_INVALID);  // JobNotInBufferError
    int dest = s.t_loc2[t];
    int to_buf = dest < I.M ? 3 * dest : 3 * I.M + I.T + (dest - I.M);
    s.t_carry[t] = NONE8;
    put_in_buffer(s, c, to_buf, j);
    if (JSL_UNLIKELY(c.raise)) return;
    int occupied_for = 0;
    if (JSL_UNLIKELY(JSL_LVL_OUT(CL) && I.has_t_out)) {
        OutS<JW>& o = *(OutS<JW>*)c.aux->outs;
        occupied_for = sample_outages<ST>(c, s.draw_ctr, c.now, I.n_t_out, I.t_out_freq, I.t_out_dur, I.off_tof, I.off_tod,
                                      o.to_active + t * MAX_OUT, o.to_a + t * MAX_OUT, o.to_b + t * MAX_OUT);
        if (JSL_UNLIKELY(c.raise)) return;
    }
    // (chain fusion: without transport outages the OUTAGE state lasts zero time and OUTAGE -> IDLE, whose only effect is
    // the state itself, would follow in the next batch of the same `now`; no handler reads the state of an AGV without a job)
    s.t_state[t] = ((JSL_LVL_OUT(CL) && I.has_t_out) || !JSL_FUSE_ON(2)) ? JSL_T_OUTAGE : JSL_T_IDLE;
    if (s.t_state[t] == JSL_T_IDLE) JSL_FUSED(2);
    s.t_occkind[t] = OCC_TIME; s.t_occ[t] = c.now + occupied_for;
    s.t_lockind[t] = 0; s.t_loc0[t] = (uint8_t)dest;
    s.j_agv[j] = NONE8;
    s.t_job[t] = NONE8;
    leg_end(s, c, t, LEG_PENDING, true);
}

// process one transport transition; false = validation error
template <int JW, bool ST, int CL>
JSL_D bool apply_transport(EnvS<JW, ST, CL>& s, Ctx& c, int t, int to, int job, bool fuse) {
    if (JSL_LVL_BUF(CL) && !c.I->flex_post && !transport_valid(s, t, to)) return false;
    // Chain fusion (rules: apply_machine), only with FLEX postbuffers / standalone buffers -- otherwise another AGV's
    // waiting time is derived from this one's occupied_till and queue positions matter:
    //   IDLE -> PICKUP -> WAITINGPICKUP   the AGV already stands at the pickup place (travel time 0) and its job is still
    //                                     being processed beyond `now`: it ends up waiting for the future
    //   PICKUP -> WAITINGPICKUP -> TRANSIT  the job is ready on arrival (in a postbuffer / standalone buffer) and the
    //                                     trip takes a static time > 0; needs unbounded capacities (the job leaves its
    //                                     buffer one batch early)
    //   TRANSIT -> OUTAGE -> IDLE         no transport outages (in agv_transit_to_outage)
    const InstDev& I = *c.I;
    const bool flex = fuse && (!JSL_LVL_BUF(CL) || I.flex_post != 0);
    int from = s.t_state[t];
    if (from == JSL_T_IDLE && to == JSL_T_WORKING) {
        agv_idle_to_working(s, c, t, job);
        if (!flex || !JSL_FUSE_ON(3) || JSL_UNLIKELY(c.raise) || s.t_occ[t] != c.now) return true;
        const int b = s.j_loc[job];
        const bool in_machine = b < 3 * I.M && (b % 3) != 1;  // prebuffer or machine buffer: waits for the op's end
        if (!(in_machine && s.j_proc[job] && s.j_end[job] > c.now)) return true;
        from = JSL_T_PICKUP; to = JSL_T_WAITINGPICKUP;
        JSL_FUSED(3);
    }
    if ((from == JSL_T_PICKUP || from == JSL_T_WAITINGPICKUP) && to == JSL_T_WAITINGPICKUP) {  // handler.py:642-702, :989-1030
        const bool arriving = from == JSL_T_PICKUP;
        agv_to_waitingpickup(s, c, t, job);
        if (!flex || !arriving || !JSL_FUSE_ON(4) || JSL_UNLIKELY(c.raise)) return true;
        if (JSL_LVL_BUF(CL) && !I.caps_inf) return true;
        if (s.t_occkind[t] != OCC_TIME || s.t_occ[t] != c.now) return true;
        const int b = s.j_loc[job];
        if (!(is_standalone(I, b) || (b < 3 * I.M && (b % 3) == 1))) return true;  // not ready: keeps waiting
        {   // the trip must end strictly later, by a time that is not redrawn (agv_to_transit resamples stochastic ones)
            const int src = place_of_buf(I, b);
            const int nip = next_idle_place(s, c, job);
            const int dst = nip < 0 ? (I.out_buf < 0 ? -1 : I.M + (I.out_buf - 3 * I.M - I.T)) : (int)s.j_mach[job];
            if (src < 0 || dst < 0) return true;
            if (ST && I.tobj_kind[I.off_travel + src * I.NP + dst] != JSL_TIME_STATIC) return true;
            if (travel_time_cur<ST, CL>(c, src, dst) <= 0) return true;
        }
        from = JSL_T_WAITINGPICKUP; to = JSL_T_TRANSIT;
        JSL_FUSED(4);
    }
    if ((from == JSL_T_WAITINGPICKUP || from == JSL_T_PICKUP) && to == JSL_T_TRANSIT) agv_to_transit(s, c, t, job);
    else if ((from == JSL_T_WORKING || from == JSL_T_TRANSIT) && to == JSL_T_OUTAGE) agv_transit_to_outage(s, c, t, job);
    else if (from == JSL_T_OUTAGE && to == JSL_T_IDLE) {  // handler.py:961-986
        if (JSL_LVL_OUT(CL) && c.I->has_t_out) {
            OutS<JW>& o = *(OutS<JW>*)c.aux->outs;
            release_outages(c.I->n_t_out, o.to_active + t * MAX_OUT, o.to_a + t * MAX_OUT, o.to_b + t * MAX_OUT);
        }
        s.t_state[t] = JSL_T_IDLE;
    } else { if (!c.raise) c.raise = JSL_ST_OTHER_INVALID; }  // NotImplementedError, handler.py:1125
    return true;
}

// ------------------------------------------------------------------------------------ time machines
// time_machines.py:133-221
template <int JW, bool ST, int CL>
JSL_D int force_jump_to_event(const EnvS<JW, ST, CL>& s, Ctx& c) {
    const InstDev& I = *c.I;
    int bj = wp_min<JW>(I.J, [&](int j) { return s.j_proc[j] ? s.j_end[j] : INT_BIG; });
    bool bad = false;
    int bt = INT_BIG;
    if (I.T > 0) {
        Mask<JW> nt = wp_ballot<JW>(I.T, [&](int t) {
            return s.t_state[t] != JSL_T_IDLE && s.t_occkind[t] == OCC_NOTIME;
        });
        bad = nt.any();
        bt = wp_min<JW>(I.T, [&](int t) {
            return (s.t_state[t] != JSL_T_IDLE && s.t_occkind[t] == OCC_TIME) ? s.t_occ[t] : INT_BIG;
        });
    }
    if (bad) JSL_RAISE_V(c, JSL_ST_OTHER_INVALID, c.now);
    // INT_BIG doubles as "nobody": a real end time never reaches it
    if (bj == INT_BIG && bt == INT_BIG) return c.now + 1;  // jump_by_one
    return bj < bt ? bj : bt;
}

// ------------------------------------------------------------------------------------ is anything due?
// The predicates create_timed applies to one component (handler.py:55-114, :329-379).  Handlers only ever write their
// own component's state / occupied_till, so after a batch the components that can be due at the unchanged `now` are
// the ones the batch touched -- sm_step tests those and skips the next scan when none is (see there).
template <int JW, bool ST, int CL>
JSL_D bool machine_due(const EnvS<JW, ST, CL>& s, int m, int now) {
    return s.m_state[m] != JSL_M_IDLE && s.m_occ[m] != JSL_NO_TIME && s.m_occ[m] <= now;
}
template <int JW, bool ST, int CL>
JSL_D bool transport_due(const EnvS<JW, ST, CL>& s, int t, int now) {
    const int st = s.t_state[t];
    return s.t_occkind[t] == OCC_TIME && s.t_occ[t] <= now && st != JSL_T_IDLE && st != JSL_T_WORKING;
}
// ------------------------------------------------------------------------------------ timed transitions
// handler.py:55-153, :329-379: evaluate on the current snapshot into the x_* scratch arrays.
// returns true when at least one transition was created.
template <int JW, bool ST, int CL>
JSL_D bool create_timed(EnvS<JW, ST, CL>& s, Ctx& c, Mask<1>& mm, Mask<JW>& tm, bool& time_due) {
    const InstDev& I = *c.I;
    const int now = c.now;
    // machines: occupied_till reached -> next state
    mm = wp_ballot<1>(I.M, [&](int m) {
        int st = s.m_state[m];
        bool due = s.m_occ[m] != JSL_NO_TIME && s.m_occ[m] <= now && st != JSL_M_IDLE;
        if (due) { s.x_mkind[m] = (uint8_t)((st + 1) & 3); s.x_mjob[m] = s.m_job[m]; }
        return due;
    });
    time_due = mm.any();
    JSL_SYNCWARP();
    if (JSL_UNLIKELY(JSL_LVL_BUF(CL) && !I.flex_pre)) {  // handler.py:117-153: automatic start from FIFO/LIFO/DUMMY prebuffers
        // machines whose prebuffer holds a job (one OR-reduction over the jobs), then one min-reduction per
        // machine that really starts: the head (smallest stamp) or, for LIFO, the tail of its queue
        const int M3 = 3 * I.M;
        const uint32_t pre_occ = wp_or<JW>(I.J, [&](int j) {
            const int b = s.j_loc[j];
            return (b < M3 && b % 3 == 0) ? 1u << (b / 3) : 0u;
        });
        if (pre_occ) {
            Mask<1> lifo;
            lifo.w[0] = 0;
            Mask<1> cand = wp_ballot<1>(I.M, [&](int m) {
                return ((pre_occ >> m) & 1u) && !mm.test(m) && s.m_state[m] == JSL_M_IDLE && I.pre_type[m] != JSL_BUF_FLEX;
            });
            if (cand.any()) lifo = wp_ballot<1>(I.M, [&](int m) { return I.pre_type[m] == JSL_BUF_LIFO; });
            while (cand.any()) {
                const int m = cand.pop_first(), b = 3 * m;
                const bool lf = lifo.test(m);
                // seq stamps are unique per env, so the stamp identifies the job
                const int k = wp_min<JW>(I.J, [&](int j) {
                    return s.j_loc[j] == b ? (lf ? (int)(0x7ffffffe - s.j_seq[j]) : (int)s.j_seq[j]) : INT_BIG;
                });
                const uint32_t sq = lf ? (uint32_t)(0x7ffffffe - k) : (uint32_t)k;
                const int j = wp_ballot<JW>(I.J, [&](int q) { return s.j_loc[q] == b && s.j_seq[q] == sq; }).first();
                s.x_mkind[m] = JSL_M_SETUP; s.x_mjob[m] = (uint8_t)j;
                mm.w[0] |= 1u << m;
            }
        }
    }
    // (a due machine with an empty buffer -- IndexError in the reference -- is caught by the handlers: job == NONE8)
    // transports (a PICKUP/WAITINGPICKUP transport without a job or a TRANSIT one without cargo -- raises in
    // the reference -- is caught by the handlers: job == NONE8)
    Mask<JW> td;
    const bool flex_post = !JSL_LVL_BUF(CL) || I.flex_post != 0;
    tm = wp_ballot<JW>(I.T, [&](int t) {
        s.x_tkind[t] = NONE8;
        if (s.t_occkind[t] != OCC_TIME || s.t_occ[t] > now) return false;
        int st = s.t_state[t];
        if (st == JSL_T_IDLE || st == JSL_T_WORKING) return false;
        int kind = JSL_T_WAITINGPICKUP, job = s.t_job[t];
        if (st == JSL_T_WAITINGPICKUP) {
            if (flex_post && job != NONE8) {  // handler.py:191-252: ready <=> in a postbuffer / standalone buffer
                int b = s.j_loc[job];
                if (is_standalone(I, b) || (b < 3 * I.M && (b % 3) == 1)) kind = JSL_T_TRANSIT;
            }
        } else if (st == JSL_T_TRANSIT) { kind = JSL_T_OUTAGE; job = s.t_carry[t]; }
        else if (st == JSL_T_OUTAGE) { kind = JSL_T_IDLE; job = NONE8; }
        s.x_tkind[t] = (uint8_t)kind; s.x_tjob[t] = (uint8_t)job;
        if (JSL_LVL_BUF(CL)) s.x_tcomp[t] = (uint8_t)t;
        return true;
    });
    time_due = time_due || tm.any();
    JSL_SYNCWARP();
    if (JSL_UNLIKELY(!flex_post)) {
        // handler.py:191-252: WAITINGPICKUP & ready -> TRANSIT (lane t tests its own job's queue position)
        wp_for<JW>(I.T, [&](int t) {
            if (tm.test(t) && s.t_state[t] == JSL_T_WAITINGPICKUP && s.t_job[t] != NONE8 &&
                ready_for_pickup_lane(s, I, s.t_job[t]))
                s.x_tkind[t] = JSL_T_TRANSIT;
        });
    }
    if (JSL_UNLIKELY(JSL_LVL_BUF(CL) && !I.flex_post)) {  // handler.py:278-326,356-359 resolved TimeDependency -> stored transition
        td = wp_ballot<JW>(I.T, [&](int t) {
            if (s.t_occkind[t] != OCC_TD) return false;
            const int b = s.t_tdbuf[t];
            const int nxt = queue_next_lane(s, I, b, buf_type(I, b));
            const int mine = s.t_job[t] == NONE8 ? -1 : (int)s.t_job[t];
            if (mine != nxt && s.j_agv[s.t_occ[t]] == NONE8) return false;  // still blocked, nobody fetches the blocker
            s.x_tkind[t] = JSL_T_WAITINGPICKUP; s.x_tcomp[t] = s.t_tdcomp[t]; s.x_tjob[t] = s.t_tdjob[t];
            return true;
        });
#pragma unroll
        for (int w = 0; w < JW; w++) tm.w[w] |= td.w[w];
    }
    JSL_SYNCWARP();
    return mm.any() || tm.any();
}

// ------------------------------------------------------------------------------------ classic fast paths
// In classic instances (CL: FLEX unbounded buffers, zero setup / travel / outage times) an operation's start
// and its completion each unroll into a fixed chain of same-time transitions, one per loop iteration of
// state.py:154-295.  The two functions below apply such a chain in one go when -- and only when -- the batch
// has exactly the chain's shape; every other situation takes the general loop.  Each is the composition of the
// handlers named in its comment, so the state after it equals the state after the iterations it replaces.

// Start of an operation by the agent: machine IDLE->SETUP (stage 0), SETUP->WORKING plus the teleport
// IDLE->PICKUP of the first idle AGV for the now running job (first batch), PICKUP->WAITINGPICKUP (second
// batch).  Preconditions: no lonely job in the quiet state the step starts from (so the started job is the only
// teleport candidate) and a positive duration (else WORKING->OUTAGE is due in the second batch).  Leaves the
// candidate masks of the resulting state in s.c_*.
template <int JW, bool ST, int CL>
JSL_D bool classic_start(EnvS<JW, ST, CL>& s, Ctx& c, const Trans& act) {
    const InstDev& I = *c.I;
    Poss<JW> p = cached_poss(s);
    if (p.lr.any() || p.li.any()) return false;
    const int m = act.comp, j = act.job;
    const int k = s.j_k[j];
    const int o = I.job_op_start[j] + k;
    const int dur = c.aux->op_dur[o];
    if (dur <= 0 || I.out_buf < 0) return false;
    if (s.j_loc[j] != 3 * m || s.m_job[m] != NONE8 || s.m_state[m] != JSL_M_IDLE) return false;
    const int now = c.now, end = now + dur;
    // machine_idle_to_setup + machine_setup_to_working
    s.j_proc[j] = 1; s.j_start[j] = now; s.j_end[j] = end;
    s.j_loc[j] = (uint16_t)(3 * m + 2);
    s.m_job[m] = (uint8_t)j;
    s.m_state[m] = JSL_M_WORKING; s.m_occ[m] = end; s.m_tool[m] = (uint8_t)((c.aux->op_mt[o] >> 16) & 0xffff);
    if (c.C->allow_early_transport != 0) {
        const int a = p.idle.first();
        if (a >= 0) {  // agv_idle_to_working + agv_to_waitingpickup (waits for the processing op's end)
            int target = (k + 1 < job_nops(I, j)) ? (c.aux->op_mt[o + 1] & 0xffff) : I.M + (I.out_buf - 3 * I.M - I.T);
            s.t_occkind[a] = OCC_TIME; s.t_occ[a] = end;
            s.t_lockind[a] = 1; s.t_loc1[a] = (uint16_t)(3 * m + 2); s.t_loc2[a] = (uint8_t)target;
            s.t_state[a] = JSL_T_WAITINGPICKUP;
            s.t_job[a] = (uint8_t)j;
            s.j_agv[j] = (uint8_t)a;
            leg_open(s, c, a, j, now, end);
            p.idle.clear(a);
        } else {
            p.lr.set(j);
        }
    }
    // machine m is busy now: nobody waiting in front of it can start
    const Mask<JW> at_m = wp_ballot<JW>(I.J, [&](int q) { return s.j_mach[q] == m; });
#pragma unroll
    for (int w = 0; w < JW; w++) p.p1.w[w] &= ~at_m.w[w];
    store_poss(s, p, false);
    JSL_SYNCWARP();
    return true;
}

// Completion of operations: the batch holds only WORKING->OUTAGE of machines and WAITINGPICKUP->WAITINGPICKUP
// of the AGVs waiting for exactly those jobs.  The chain is OUTAGE->IDLE (job into the postbuffer) |
// WAITINGPICKUP->TRANSIT | TRANSIT->OUTAGE (job into the next prebuffer / output buffer) | OUTAGE->IDLE, all at
// `now`; buffer stamps keep the reference's order (machines by index, then AGVs by index).
template <int JW, bool ST, int CL>
JSL_D bool classic_complete(EnvS<JW, ST, CL>& s, Ctx& c, const Mask<1>& mm, const Mask<JW>& tm) {
    const InstDev& I = *c.I;
    const bool bad_m = wp_ballot<1>(I.M, [&](int m) {
        if (!mm.test(m)) return false;
        const int j = s.m_job[m];
        return s.x_mkind[m] != JSL_M_OUTAGE || j == NONE8 || !s.j_proc[j];
    }).any();
    if (bad_m) return false;
    const bool bad_t = wp_ballot<JW>(I.T, [&](int t) {
        if (!tm.test(t)) return false;
        const int j = s.t_job[t];
        if (s.x_tkind[t] != JSL_T_WAITINGPICKUP || s.t_state[t] != JSL_T_WAITINGPICKUP || j == NONE8 ||
            s.t_lockind[t] != 1 || s.t_carry[t] != NONE8)
            return true;
        const int b = s.j_loc[j];
        if (b >= 3 * I.M || (b % 3) != 2) return true;
        return !mm.test(b / 3) || s.m_job[b / 3] != j;
    }).any();
    if (bad_t) return false;
    const int now = c.now;
    const int seq0 = s.seq_ctr, nm = mm.count();
    JSL_SYNCWARP();
    // machine_working_to_outage + machine_outage_to_idle
    wp_for<1>(I.M, [&](int m) {
        if (!mm.test(m)) return;
        const int j = s.m_job[m];
        const int k = s.j_k[j];
        const int o = I.job_op_start[j] + k;
        if (c.aux->op_log) { c.aux->op_log[2 * o] = s.j_start[j]; c.aux->op_log[2 * o + 1] = now; }
        s.j_end[j] = now;
        s.j_k[j] = (uint8_t)(k + 1);
        s.j_proc[j] = 0;
        s.j_exec[j] |= 1u << m;
        s.j_mach[j] = (k + 1 < job_nops(I, j)) ? (uint8_t)(c.aux->op_mt[o + 1] & 0xffff) : (uint8_t)NONE8;
        s.m_done[m] = (uint16_t)(s.m_done[m] + 1);
        s.m_job[m] = NONE8;
        s.j_loc[j] = (uint16_t)(3 * m + 1);
        s.j_seq[j] = (uint32_t)(seq0 + JSL_POPC(mm.w[0] & ((1u << m) - 1u)));
        s.m_state[m] = JSL_M_IDLE; s.m_occ[m] = now;
    });
    if (now > s.max_done_end) s.max_done_end = now;
    // agv_to_waitingpickup (x2) + agv_to_transit + agv_transit_to_outage + OUTAGE->IDLE
    wp_for<JW>(I.T, [&](int t) {
        if (!tm.test(t)) return;
        int rank = JSL_POPC(tm.word_of(t) & ((1u << (t & 31)) - 1u));
#pragma unroll
        for (int w = 0; w < JW; w++)
            if (w < (t >> 5)) rank += JSL_POPC(tm.w[w]);
        const int j = s.t_job[t];
        const int dest = s.t_loc2[t];
        s.j_loc[j] = (uint16_t)(dest < I.M ? 3 * dest : 3 * I.M + I.T + (dest - I.M));
        s.j_seq[j] = (uint32_t)(seq0 + nm + rank);
        s.j_agv[j] = NONE8;
        s.t_state[t] = JSL_T_IDLE;
        s.t_occkind[t] = OCC_TIME; s.t_occ[t] = now;
        s.t_lockind[t] = 0; s.t_loc0[t] = (uint8_t)dest;
        s.t_job[t] = NONE8;
        leg_end(s, c, t, now, true);
    });
    s.seq_ctr = seq0 + nm + tm.count();
    JSL_SYNCWARP();
    return true;
}

// ------------------------------------------------------------------------------------ state.step
constexpr int TM_JUMP = 0, TM_FORCE = 1;

// state.py:154-295 as ONE loop with a single call site per phase (the whole hot path has to stay
// inside the instruction cache):
//   stage 0 : the action's transition (0 or 1) is queued like a timed transition and applied
//   stage 1+: candidates -> time machine -> timed transitions (+ teleports in the first batch) -> apply
// The candidate masks computed at the top of the last iteration describe the final state (nothing
// was applied after them), so they double as the new possible_transitions.
// Returns success (false = validation error -> StateMachineResult.success=False); escaping exceptions
// are left in c.raise.
template <int JW, bool ST, int CL>
JSL_D bool sm_step(EnvS<JW, ST, CL>& s, Ctx& c, bool has_action, Trans act, int time_machine) {
    const InstDev& I = *c.I;
    c.now = s.time;
    // transition masks of the current batch and the loop guard.  The stochastic kernels are short of registers (the
    // samplers are inlined into the handlers): there these live in the env's shared-memory slot instead of being
    // spilled to local memory around every handler; everywhere else they are plain registers.
    struct LoopState { Mask<1> mm; Mask<JW> tm, tele; int guard; bool f_action, f_first, f_eval, f_quiet, f_fresh, f_moved; };
    static_assert(sizeof(LoopState) <= sizeof(EnvAux::loop_scratch), "loop_scratch too small");
    LoopState* const ls = reinterpret_cast<LoopState*>(c.aux->loop_scratch);
    Mask<1> mm_r;
    Mask<JW> tm_r, tele_r;
    int guard_r;
    Mask<1>& mm = ST ? ls->mm : mm_r;
    Mask<JW>& tm = ST ? ls->tm : tm_r;
    Mask<JW>& tele = ST ? ls->tele : tele_r;
    int& guard = ST ? ls->guard : guard_r;
    mm.w[0] = 0;
#pragma unroll
    for (int w = 0; w < JW; w++) { tm.w[w] = 0; tele.w[w] = 0; }
    // the candidate masks live in the state (c_p1 / c_idle / c_lr / c_li), not in registers
    // (the loop's flags as well: in the stochastic kernels they were spilled and reloaded around every handler)
    bool action_r, first_r, eval_r, quiet_r, fresh_r, moved_r;
    bool& action_stage = ST ? ls->f_action : action_r;
    bool& first = ST ? ls->f_first : first_r;
    bool& evaluated = ST ? ls->f_eval : eval_r;      // the time machine already ran for the current state
    bool& known_quiet = ST ? ls->f_quiet : quiet_r;
    bool& poss_fresh = ST ? ls->f_fresh : fresh_r;   // s.c_* already describe the current state
    bool& moved = ST ? ls->f_moved : moved_r;
    action_stage = has_action;
    first = true;
    evaluated = false;
    // a step starts from the quiet state the previous call ended in (its loop only exits when nothing is due at
    // `now`), and the classic fast paths end in one: no timed transition can exist in the next batch
    known_quiet = !has_action && time_machine == TM_FORCE;  // (a reset starts from the initial state instead)
    poss_fresh = known_quiet;
    moved = CL != 2 || (!has_action && time_machine == TM_JUMP);  // a job may have reached an output buffer (reset: unknown)
    guard = 0;
    if (CL == 2 && has_action && act.kind == 0 && classic_start(s, c, act)) {
        action_stage = false; first = false; known_quiet = true; poss_fresh = true;
#ifdef JSL_HOST_SIM
        g_fast_start++;
#endif
    }
#pragma unroll 1
    for (;;) {
        if (action_stage) {
            if (act.kind == 0) {
                s.x_mkind[act.comp] = (uint8_t)act.new_state; s.x_mjob[act.comp] = (uint8_t)act.job;
                mm.w[0] = 1u << act.comp;
            } else {
                s.x_tkind[act.comp] = (uint8_t)act.new_state;
                if (JSL_LVL_BUF(CL)) s.x_tcomp[act.comp] = (uint8_t)act.comp;
                s.x_tjob[act.comp] = (uint8_t)act.job;
                tm.set(act.comp);
            }
        } else {
            bool time_due = false;
            bool have = false;
            if (!known_quiet) {
                have = create_timed(s, c, mm, tm, time_due);
                if (JSL_UNLIKELY(c.raise)) return true;
            }
#ifdef JSL_HOST_SIM
            else {  // the skipped scan must have found nothing
                Mask<1> qm; Mask<JW> qt; bool qd = false;
                int save = c.raise;
                g_invariant_checks++;
                if (create_timed(s, c, qm, qt, qd) || c.raise != save) g_invariant_violations++;
                c.raise = save;
            }
#endif
            known_quiet = false;
            // Time machine (time_machines.py:61-130): `now` stays while something is possible, else it jumps to
            // min(processing-op ends, busy-AGV times).  When an occupied_till-driven transition is due at `now`
            // that minimum IS `now` (nothing pending can lie in the past: every due component fires in the very
            // next batch), so the candidate scan is only needed in quiet batches and in the first one.
            if (!evaluated && (first || !time_due)) {
                if (!poss_fresh) store_poss(s, compute_poss(s, c), false);
                poss_fresh = false;
                const Poss<JW> p = cached_poss(s);
                const bool anyp = p.p1.any() || (p.idle.any() && (p.lr.any() || p.li.any()));
                if ((first && time_machine == TM_FORCE) || !anyp) {
                    const int t2 = force_jump_to_event(s, c);
                    if (JSL_UNLIKELY(c.raise)) return true;
                    if (t2 != c.now) {  // re-evaluate the timed transitions at the new time
                        c.now = t2;
                        evaluated = true;
                        leg_stamp(s, c);
                        if (++guard > 100000) { c.raise = JSL_ST_OTHER_INVALID; return true; }
                        continue;
                    }
                }
            }
#ifdef JSL_HOST_SIM
            if (!evaluated && !(first || !time_due)) {  // the skipped evaluation must have been a no-op
                Poss<JW> q = compute_poss(s, c);
                const bool anyq = q.p1.any() || (q.idle.any() && (q.lr.any() || q.li.any()));
                int save = c.raise;
                g_invariant_checks++;
                if (!anyq && force_jump_to_event(s, c) != c.now) g_invariant_violations++;
                c.raise = save;
            }
#endif
            evaluated = false;
            leg_stamp(s, c);
            const Poss<JW> p = cached_poss(s);
            if (first && p.idle.any() && (p.lr.any() || p.li.any())) {
                // teleports (state.py:229-232, :414-467): i-th idle AGV <- i-th lonely job with zero travel time
                auto zero_travel = [&](int j, bool& bad) {
                    int cur = place_of_buf(I, s.j_loc[j]);
                    int nxt = next_idle_place(s, c, j);
                    if (nxt < 0) nxt = I.out_buf < 0 ? -2 : I.M + (I.out_buf - 3 * I.M - I.T);
                    if (nxt == -2) { bad = true; return false; }
                    if (cur == nxt) return true;
                    if (cur < 0) { bad = true; return false; }
                    int tt = travel_time_cur<ST, CL>(c, cur, nxt);
                    if (tt == JSL_NO_TIME) { bad = true; return false; }
                    return tt == 0;
                };
                Mask<JW> zr, zi;
                if (CL == 2) {  // every travel time is zero: all lonely jobs qualify
                    zr = p.lr; zi = p.li;
                } else {
                    zr = wp_ballot<JW>(I.J, [&](int j) { bool bd = false; return p.lr.test(j) && zero_travel(j, bd); });
                    zi = wp_ballot<JW>(I.J, [&](int j) { bool bd = false; return p.li.test(j) && zero_travel(j, bd); });
                    if (I.travel || ST || I.out_buf < 0) {
                        Mask<JW> badm = wp_ballot<JW>(I.J, [&](int j) {
                            bool bd = false;
                            if (p.lr.test(j) || p.li.test(j)) zero_travel(j, bd);
                            return bd;
                        });
                        if (badm.any()) { c.raise = JSL_ST_OTHER_INVALID; return true; }
                    }
                }
                Mask<JW> idle = p.idle;
                while (idle.any() && (zr.any() || zi.any())) {
                    int t = idle.pop_first();
                    int j = zr.any() ? zr.pop_first() : zi.pop_first();
                    s.x_tkind[t] = JSL_T_WORKING; s.x_tjob[t] = (uint8_t)j;
                    if (JSL_LVL_BUF(CL)) s.x_tcomp[t] = (uint8_t)t;
                    tele.set(t);
                }
            }
            if (!have && !tele.any()) break;
            if (CL == 2 && !tele.any() && classic_complete(s, c, mm, tm)) {
                mm.w[0] = 0;
#pragma unroll
                for (int w = 0; w < JW; w++) tm.w[w] = 0;
                first = false; known_quiet = true; moved = true;
#ifdef JSL_HOST_SIM
                g_fast_complete++;
#endif
                continue;  // (ends in a quiet state: the loop only comes back here after a guarded time jump)
            }
        }
#ifdef JSL_HOST_SIM
        if (CL == 2) g_general_batches++;
#endif
        // apply: machines, then timed transports, then teleports -- in creation order
        moved = true;
        // Levels with FLEX buffers only (simple, features): does this batch leave anything due at `now`?  Only the
        // components it touches can be (above); if none is, the next iteration's scan would find nothing (42-52 % of
        // the scans did) and is skipped like the one of a step that starts from a quiet state.  (The CPU debug build
        // runs the skipped scan and counts what it finds.)  In the levels with ordered buffers a prebuffer start or a
        // TimeDependency hand-over may be pending as well; testing for those cost more than the skipped scans saved
        // (queues kernel 370 -> 361 M env-steps/s), so they keep scanning.
        constexpr bool TRACK_DUE = CL != 2 && !JSL_LVL_BUF(CL);
        if (TRACK_DUE) known_quiet = true;
        while (mm.any()) {
            int m = mm.pop_first();
            bool ok = apply_machine(s, c, m, s.x_mkind[m], s.x_mjob[m], !action_stage);
            if (JSL_UNLIKELY(c.raise)) return true;
            if (JSL_UNLIKELY(!ok)) return false;
            if (TRACK_DUE && machine_due(s, m, c.now)) known_quiet = false;
        }
        if (CL == 2) {  // (the classic kernel keeps the loop it was tuned with: its layout is worth 2 % of the headline)
#pragma unroll 1
            for (int ph = 0; ph < 2; ph++) {
                Mask<JW> cur = ph == 0 ? tm : tele;
                while (cur.any()) {
                    int t = cur.pop_first();
                    bool ok = apply_transport(s, c, t, s.x_tkind[t], s.x_tjob[t], !action_stage);
                    if (JSL_UNLIKELY(c.raise)) return true;
                    if (JSL_UNLIKELY(!ok)) return false;
                }
            }
#pragma unroll
            for (int w = 0; w < JW; w++) { tm.w[w] = 0; tele.w[w] = 0; }
        } else {
            // timed transports first, then the teleports: one loop, one call site, the masks consumed in place (no
            // copy of the current mask to keep live -- or, in the stochastic kernels, to spill -- across the handlers)
            for (;;) {
                int t = tm.pop_first();
                if (t < 0) t = tele.pop_first();
                if (t < 0) break;
                const int comp = JSL_LVL_BUF(CL) ? (int)s.x_tcomp[t] : t;
                bool ok = apply_transport(s, c, comp, s.x_tkind[t], s.x_tjob[t], !action_stage);
                if (JSL_UNLIKELY(c.raise)) return true;
                if (JSL_UNLIKELY(!ok)) return false;
                if (TRACK_DUE && transport_due(s, comp, c.now)) known_quiet = false;
            }
        }
        JSL_SYNCWARP();
        if (action_stage) { action_stage = false; leg_stamp(s, c); }
        else first = false;
        if (++guard > 100000) { c.raise = JSL_ST_OTHER_INVALID; return true; }  // the reference would spin
    }
    s.time = c.now;
    s.skip = 0;
    // core_utils.py:44-68 is_done, state.py:270-283
    bool done = s.all_out != 0;
    if (moved) done = !wp_ballot<JW>(I.J, [&](int j) { return !is_output(I, s.j_loc[j]); }).any();
    if (done) {
        if (s.max_done_end >= 0) s.time = s.max_done_end;
        s.n_possible = 0;
    } else {
        s.n_possible = cached_poss(s).total();
    }
    s.all_out = done ? 1 : 0;
    if (done) store_poss(s, cached_poss(s), true);
    return true;
}

// ------------------------------------------------------------------------------------ reset
template <int JW, bool ST, int CL>
JSL_D void env_init(EnvS<JW, ST, CL>& s, Ctx& c) {
    const InstDev& I = *c.I;
    s.time = I.start_time; s.seq_ctr = 0; s.n_possible = 0; s.skip = 0;
    s.joker = c.C->truncation_joker; s.step_noop = 0; s.step_action = 0; s.reward_noop = 0;
    s.n_steps = 0; s.noop_count = 0; s.max_done_end = -1; s.status = JSL_ST_OK; s.all_out = 0;
    s.terminated = 0; s.truncated = 0; s.last_action_empty = 1; s.obs_done = 0;
    s.ret = 0.0;
    {   // per-variant constants travel with the state: no dependent table reads on the step path
        const int mat = c.aux->lb_mat[1];
        s.lower_bound = c.aux->lb_mat[0]; s.max_allowed_time = mat;
        s.inv_mat = 1.0 / (double)mat;
        s.pad_[0] = 0; s.pad_[1] = 0; s.pad_[2] = 0;
    }
    JSL_SYNCWARP();
    wp_for<JW>(EnvS<JW, ST, CL>::MJ, [&](int j) {
        bool live = j < I.J;
        s.j_start[j] = -1; s.j_end[j] = -1;
        s.j_seq[j] = (uint32_t)j;  // initial stores list jobs in job order (mapper.py:1296-1321)
        s.j_loc[j] = live ? I.job_init_buf[j] : (uint16_t)0xffff;
        s.j_k[j] = 0; s.j_proc[j] = 0; s.j_agv[j] = NONE8; s.j_exec[j] = 0;
        s.j_mach[j] = (live && job_nops(I, j) > 0) ? (uint8_t)(c.aux->op_mt[I.job_op_start[j]] & 0xffff) : (uint8_t)NONE8;
    });
    wp_for<1>(32, [&](int m) {
        s.m_occ[m] = JSL_NO_TIME; s.m_done[m] = 0; s.m_state[m] = JSL_M_IDLE;
        s.m_tool[m] = (uint8_t)I.default_tool; s.m_job[m] = NONE8; s.x_mkind[m] = 0; s.x_mjob[m] = NONE8;
    });
    wp_for<JW>(EnvS<JW, ST, CL>::MT, [&](int t) {
        s.t_occ[t] = -1; s.t_loc1[t] = 0; s.t_tdbuf[t] = 0;
        s.t_state[t] = JSL_T_IDLE; s.t_job[t] = NONE8; s.t_carry[t] = NONE8; s.t_occkind[t] = OCC_NOTIME;
        s.t_lockind[t] = 0; s.t_loc0[t] = t < I.T ? I.agv_init_place[t] : 0; s.t_loc2[t] = 0;
        s.t_tdcomp[t] = NONE8; s.t_tdjob[t] = NONE8; s.x_tkind[t] = NONE8; s.x_tcomp[t] = 0; s.x_tjob[t] = NONE8;
    });
    if (c.aux->outs) {
        OutS<JW>& o = *(OutS<JW>*)c.aux->outs;
        wp_for<1>(32, [&](int m) {
            for (int k = 0; k < MAX_OUT; k++) { o.mo_active[m * MAX_OUT + k] = 0; o.mo_a[m * MAX_OUT + k] = -1; o.mo_b[m * MAX_OUT + k] = -1; }
        });
        wp_for<JW>(EnvS<JW, ST, CL>::MT, [&](int t) {
            for (int k = 0; k < MAX_OUT; k++) { o.to_active[t * MAX_OUT + k] = 0; o.to_a[t * MAX_OUT + k] = -1; o.to_b[t * MAX_OUT + k] = -1; }
        });
    }
    s.seq_ctr = EnvS<JW, ST, CL>::MJ;
    s.draw_ctr = 0;
    if (ST) stoch_init(c.I, c.aux->tcur, c.aux->tseed, c.aux->start_seeds, c.aux->tape, c.aux->seed, c.aux->env_id, c.aux->op_dur);
    JSL_SYNCWARP();
}

// ------------------------------------------------------------------------------------ reward / step
// rewards.py:118-148 BinaryActionJsspReward.make, IEEE double without contraction
template <int JW, bool ST, int CL>
JSL_D double make_reward(EnvS<JW, ST, CL>& s, const Ctx& c) {
    double sr;
    if (s.truncated) sr = (c.C->truncation_bias * 1.0) / c.C->sparse_bias;
    else if (!s.terminated) sr = 0.0;
    else sr = (double)(s.max_allowed_time - s.time) / (double)(s.max_allowed_time - s.lower_bound);
    if (s.last_action_empty) s.reward_noop = s.reward_noop + 1;
    else s.reward_noop = 0;
    const double dr = s.reward_noop >= c.I->J ? c.aux->neg_inv_n : 0.0;  // -int(flag) / N
#if defined(__CUDACC__) && !defined(JSL_HOST_SIM)
    return __dadd_rn(__dmul_rn(sr, c.C->sparse_bias), __dmul_rn(dr, c.C->dense_bias));
#else
    volatile double a = sr * c.C->sparse_bias;
    volatile double b = dr * c.C->dense_bias;
    return a + b;
#endif
}

// User hook (the reference's RewardFactory ABC, rewards.py:11-58): a library built with
// -DJSL_USER_HEADER='"my_hooks.cuh"' includes that header here; it must define
//     template <int JW, bool ST, int CL>
//     JSL_D double jsl_user_reward(const EnvS<JW, ST, CL>& s, const Ctx& c, double stock_reward);
// whose result replaces the reward of every env.step (the stock BinaryActionJsspReward value is handed in).  The
// state fields it may read are the ones jsl_batch_state_layout lists.  Example: csrc/examples/jsl_user_example.cuh;
// build: JSL_NVCC_EXTRA='-DJSL_USER_HEADER="\"examples/jsl_user_example.cuh\""' JSL_LIB_OUT=... python -m jobshoplab_b200.build
#ifdef JSL_USER_HEADER
#include JSL_USER_HEADER
#else
template <int JW, bool ST, int CL>
JSL_D double jsl_user_reward(const EnvS<JW, ST, CL>&, const Ctx&, double stock_reward) { return stock_reward; }
#endif

// env.py:427-454 step + middleware.py:384-431 step + :304-368 _get_no_op_result, split around the single
// sm_step call site of run_env_op().
struct StepPlan {
    bool run_sm, has_action;
    int time_machine;
    Trans act;
};

// everything before the state machine: argument checks, SubTimeStepper, what to run.
// returns false when env.step raises before touching the state machine (status is set).
template <int JW, bool ST, int CL>
JSL_D bool env_step_pre(EnvS<JW, ST, CL>& s, Ctx& c, int action, StepPlan& plan) {
    plan.run_sm = false; plan.has_action = false; plan.time_machine = TM_JUMP;
    plan.act.kind = 0; plan.act.comp = 0; plan.act.new_state = 0; plan.act.job = 0;
    if (s.status > JSL_ST_TRUNCATED) return false;                                     // a raise already escaped
    if (s.terminated || s.truncated) { s.status = JSL_ST_ENV_DONE; return false; }     // env.py:441
    const int remaining = s.n_possible - s.skip;
    if (remaining <= 0) { s.status = JSL_ST_NO_POSSIBLE; return false; }                // actions.py:141
    if (action != 0 && action != 1) { s.status = JSL_ST_ACTION_OUT_OF_SPACE; return false; }
    if (action == 0) {
        s.step_noop = s.step_noop + 1;
        if (remaining == 1) { plan.run_sm = true; plan.time_machine = TM_FORCE; }       // middleware.py:333-336
        else s.skip = s.skip + 1;                                                       // :355-368
    } else {
        s.step_action = s.step_action + 1;
        plan.run_sm = true; plan.has_action = true;
        plan.act = current_candidate(s, c);
    }
    return true;
}

// everything after it: middleware bookkeeping, terminated/truncated, reward.
template <int JW, bool ST, int CL>
JSL_D double env_step_post(EnvS<JW, ST, CL>& s, Ctx& c, int action, const StepPlan& plan, bool success) {
    if (c.raise) { s.status = c.raise; return 0.0; }
    const bool action_empty = action == 0;
    if (action == 0 && plan.run_sm) {
        if (!success) { s.status = JSL_ST_UNSUCCESSFUL; return 0.0; }  // failed result, old state is not done
        if (s.n_possible == 0) {
            if (!s.all_out) { s.status = JSL_ST_UNSUCCESSFUL; return 0.0; }             // middleware.py:344
        } else {
            if (c.C->truncation_active && s.step_action == 0) s.joker = s.joker - 1;    // :347
            s.step_noop = 1; s.step_action = 0;                                         // :351
        }
    }
    if (success) {
        s.last_action_empty = action_empty;
        s.obs_done = (s.n_possible - s.skip) == 0;
        s.noop_count = s.noop_count + (action_empty ? 1 : 0);
        s.terminated = s.all_out != 0;
        s.truncated = s.joker < 0;
    } else {
        s.obs_done = 1;
        s.truncated = 1; s.terminated = 0;
    }
    double r = jsl_user_reward(s, c, make_reward(s, c));
    s.ret = s.ret + r;
    s.n_steps = s.n_steps + 1;
    s.status = !success ? JSL_ST_STEP_FAILED : s.terminated ? JSL_ST_DONE : s.truncated ? JSL_ST_TRUNCATED : JSL_ST_OK;
    return r;
}

// One env operation: OP_RESET (env.reset: init + dummy step) or an env.step(action).  The only place
// sm_step is instantiated.
constexpr int OP_RESET = -1;
template <int JW, bool ST, int CL>
JSL_D double run_env_op(EnvS<JW, ST, CL>& s, Ctx& c, int op) {
    StepPlan plan;
    if (op == OP_RESET) {
        env_init(s, c);
        if (c.aux->op_log) wp_for_dyn(2 * c.I->N, [&](int i) { c.aux->op_log[i] = -1; });
        if (LEG_LOG_BUILD && c.log) {
            int32_t* g = c.aux->leg_log;
            wp_for_dyn(leg_hdr(EnvS<JW, ST, CL>::MT), [&](int i) { g[i] = i < 2 ? 0 : -1; });
        }
        plan.run_sm = true; plan.has_action = false; plan.time_machine = TM_JUMP;
        plan.act.kind = 0; plan.act.comp = 0; plan.act.new_state = 0; plan.act.job = 0;
    } else if (!env_step_pre(s, c, op, plan)) {
        return 0.0;
    }
    bool success = true;
    if (plan.run_sm) success = sm_step(s, c, plan.has_action, plan.act, plan.time_machine);
    if (op == OP_RESET) {
        if (LEG_LOG_BUILD && c.log) {
            // env.reset keeps no history (env.py:536): tasks that ran to their end inside the reset step are
            // never seen (marked agv < 0), open ones are first seen in the state the next env.step starts from
            int32_t* g = c.aux->leg_log;
            constexpr int MT = EnvS<JW, ST, CL>::MT;
            const int n = g[0], t0 = s.time;
            JSL_SYNCWARP();
            wp_for_dyn(n, [&](int i) {
                int32_t* e = g + leg_hdr(MT) + 5 * i;
                if (g[2 + e[0]] == i) e[2] = t0;
                else e[0] = -1 - e[0];
            });
        }
        if (c.raise) s.status = c.raise;
        else if (!success) { s.n_possible = 0; s.skip = 0; }
        return 0.0;
    }
    return env_step_post(s, c, op, plan, success);
}

// ------------------------------------------------------------------------------------ policies
JSL_D uint32_t philox_bit(uint64_t seed, uint64_t env_idx, uint32_t step) {
    uint32_t c0 = step, c1 = 0u, c2 = (uint32_t)env_idx, c3 = (uint32_t)(env_idx >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0 & 1u;
}

// dipatching_benchmark.py:17-67 + dispatch_rules_uitls.py:57-111
template <int JW, bool ST, int CL>
JSL_D int policy_action(const EnvS<JW, ST, CL>& s, const Ctx& c, int policy) {
    const InstDev& I = *c.I;
    if (policy == JSL_POLICY_RANDOM) {
        // the action of step n is philox_bit(seed, env, n): lane l draws the one of step 32*block + l, so the
        // ten Philox rounds run once per 32 steps
        const uint32_t step = (uint32_t)s.n_steps, block = (step >> 5) + 1u;
        if (c.aux->act_block != block) {
            const uint64_t seed = c.aux->seed, env_id = c.aux->env_id;
            const Mask<1> bits = wp_ballot<1>(32, [&](int l) { return philox_bit(seed, env_id, (step & ~31u) + (uint32_t)l) != 0u; });
            JSL_SYNCWARP();
            c.aux->act_bits = bits.w[0];
            c.aux->act_block = block;
            JSL_SYNCWARP();
        }
        return (int)((c.aux->act_bits >> (step & 31u)) & 1u);
    }
    if (policy == JSL_POLICY_FIFO || s.n_possible - s.skip <= 0) return 1;
    Trans cand = current_candidate(s, c);
    if (cand.kind == 1) return 1;
    const int pre = 3 * cand.comp, cj = cand.job;
    auto key = [&](int j) {
        if (policy == JSL_POLICY_SPT) return op_duration<ST>(c, I.job_op_start[j] + s.j_k[j]);
        if (!ST) return (int)c.aux->job_work[j];
        int w = 0;
        for (int o = I.job_op_start[j]; o < I.job_op_start[j + 1]; o++) w += c.aux->tcur[o];
        return w;
    };
    const int ck = key(cj);
    const uint32_t cs = s.j_seq[cj];
    // accept iff the candidate is Python's min(): first job in buffer order with the smallest key
    bool better = wp_ballot<JW>(I.J, [&](int j) {
        if (s.j_loc[j] != pre || j == cj) return false;
        int k = key(j);
        return k < ck || (k == ck && s.j_seq[j] < cs);
    }).any();
    return better ? 0 : 1;
}

// ------------------------------------------------------------------------------------ observations
// Packed record of BinaryActionObservationFactory.make (observations.py:515-551 over :419-479):
//   f32 current_time | f32 current_transition[3] | i32 job_progression[J] | i32 machine_progression[M]
//   | i8 job_running[J] | i8 available_jobs[J] | i8 machine_running[M] | i8 job_executed_on_machine[J*M]
//   zero-padded to a multiple of 16 bytes.  Rows follow the reference's string-id ordering quirk.
#if defined(__CUDACC__) && !defined(JSL_HOST_SIM)
#define JSL_HD __host__ __device__ inline
#else
#define JSL_HD static inline
#endif
JSL_HD int obs_bytes_binary(int J, int M) { return (16 + 4 * J + 4 * M + 2 * J + M + J * M + 15) & ~15; }
JSL_HD int obs_bytes_oparray(int J, int N) { return (16 + 4 * N + 4 * J + 15) & ~15; }
JSL_HD int obs_bytes_tassel(int J) { return (7 * 4 * J + 15) & ~15; }

// np.float32(time / max_allowed_time) (observations.py:470-471): the float64 quotient rounded to float32.  A
// float64 division is a ~20-instruction sequence on the SM; t * (1/m) is within 2 ulp(f64) of t / m, while the
// quotient of two integers with m < 2^22 is either a float32 value itself or at least 2^-47 (relative) away from the
// nearest float32 rounding boundary, so both round to the same float32 (checked exhaustively for m < 3000 and on
// 8e7 random pairs, tests/test_host_api.py).  Larger / degenerate m take the division.
template <int JW, bool ST, int CL>
JSL_D float time_fraction(const EnvS<JW, ST, CL>& s) {
    const int m = s.max_allowed_time;
    if (m > 0 && m < (1 << 22)) {
#if defined(__CUDACC__) && !defined(JSL_HOST_SIM)
        return (float)__dmul_rn((double)s.time, s.inv_mat);
#else
        volatile double q = (double)s.time * s.inv_mat;
        return (float)q;
#endif
    }
    return (float)((double)s.time / (double)m);
}

template <int JW, bool ST, int CL>
JSL_D void cand_floats(const EnvS<JW, ST, CL>& s, const Ctx& c, float (&tr)[3], bool known, const Trans& kc) {
    const InstDev& I = *c.I;
    if (!s.obs_done && s.n_possible - s.skip > 0) {
        const Trans cand = known ? kc : current_candidate(s, c);
        int comp_idx = cand.kind == 0 ? cand.comp : I.M + cand.comp;
        tr[0] = I.tr_comp[comp_idx];  // float32 of the float64 quotient, tabulated on the host
        tr[1] = I.tr_job[cand.job];
        tr[2] = cand.kind == 0 ? 0.0f : (float)0.33;
    } else {
        tr[0] = tr[1] = tr[2] = 1.0f;
    }
}

template <int Q>
JSL_D void exec_row_words(uint32_t* rw, uint32_t mask) {
#pragma unroll
    for (int q = 0; q < Q; q++) rw[q] = (((mask >> (4 * q)) & 15u) * 0x00204081u) & 0x01010101u;
}

template <int JW, bool ST, int CL>
JSL_D void write_obs_binary(const EnvS<JW, ST, CL>& s, const Ctx& c, uint8_t* out, bool have_cand, const Trans& cand) {
    const InstDev& I = *c.I;
    const int J = I.J, M = I.M;
    float tr[3];
    cand_floats(s, c, tr, have_cand, cand);
    float* f = (float*)out;
    int32_t* jprog = (int32_t*)(out + 16);
    int32_t* mprog = jprog + J;
    int8_t* jrun = (int8_t*)(mprog + M);
    int8_t* avail = jrun + J;
    int8_t* mrun = avail + J;
    int8_t* exec = mrun + M;
    const int total = obs_bytes_binary(J, M);
    wp_for_dyn(1, [&](int) {
        f[0] = time_fraction(s);
        f[1] = tr[0]; f[2] = tr[1]; f[3] = tr[2];
        for (int i = 16 + 4 * J + 4 * M + 2 * J + M + J * M; i < total; i++) out[i] = 0;
    });
    // rows of job_executed_on_machine are written as 32-bit words when they are word aligned
    const bool word_rows = (M & 3) == 0 && (((16 + 4 * J + 4 * M + 2 * J + M) & 3) == 0) && M <= 32;
    wp_for<JW>(J, [&](int r) {
        int j = I.job_lex[r];
        jrun[r] = (int8_t)s.j_proc[j];
        bool any_idle = (s.j_k[j] + (s.j_proc[j] ? 1 : 0)) < job_nops(I, j);
        // numeric index into the lexicographically ordered tuple (observations.py:437-441)
        avail[r] = (int8_t)(any_idle && !s.j_proc[I.job_lex[j]]);
        jprog[r] = s.j_k[r];
        const uint32_t mask = s.j_exec[j];
        int8_t* row = exec + r * M;
        if (word_rows) {
            uint32_t* rw = (uint32_t*)row;
            // 4 mask bits -> 4 bytes per word; the common row lengths are unrolled (a counted loop costs twice
            // the instructions of its body here)
            if (M == 20) exec_row_words<5>(rw, mask);
            else if (M == 16) exec_row_words<4>(rw, mask);
            else if (M == 12) exec_row_words<3>(rw, mask);
            else if (M == 8) exec_row_words<2>(rw, mask);
            else for (int q = 0; q < M / 4; q++) rw[q] = (((mask >> (4 * q)) & 15u) * 0x00204081u) & 0x01010101u;
        } else {
            for (int m = 0; m < M; m++) row[m] = (int8_t)((mask >> m) & 1u);
        }
    });
    wp_for<1>(M, [&](int r) {
        int m = I.mach_lex[r];
        mrun[r] = (int8_t)(s.m_state[m] == JSL_M_WORKING);
        mprog[r] = s.m_done[m];
    });
}

// BinaryOperationArrayObservation.make (observations.py:631-693 over :588-618):
//   f32 current_transition[3] | f32 0 | f32 operation_state[N] | f32 job_locations[J]
template <int JW, bool ST, int CL>
JSL_D void write_obs_oparray(const EnvS<JW, ST, CL>& s, Ctx& c, uint8_t* out, bool have_cand, const Trans& cand) {
    const InstDev& I = *c.I;
    float tr[3];
    cand_floats(s, c, tr, have_cand, cand);
    float* f = (float*)out;
    float* ops = f + 4;
    float* loc = ops + I.N;
    const int total = obs_bytes_oparray(I.J, I.N);
    wp_for_dyn(1, [&](int) {
        f[0] = tr[0]; f[1] = tr[1]; f[2] = tr[2]; f[3] = 0.f;
        for (int i = 16 + 4 * I.N + 4 * I.J; i < total; i++) out[i] = 0;
    });
    const int max_buffer_id = I.S + 3 * I.M + I.T - 1;
    Mask<JW> bad = wp_ballot<JW>(I.J, [&](int j) {
        int o0 = I.job_op_start[j], n = job_nops(I, j), k = s.j_k[j];
        bool zero_div = false;
        for (int i = 0; i < n; i++) {
            float v = i < k ? 1.f : 0.f;
            if (i == k && s.j_proc[j]) {
                int dur = s.j_end[j] - s.j_start[j];
                if (dur == 0) zero_div = true;
                else v = (float)((double)(s.time - s.j_start[j]) / (double)dur);
            }
            ops[o0 + i] = v;
        }
        loc[j] = (float)((double)I.buf_ref_id[s.j_loc[j]] / (double)max_buffer_id);
        return zero_div;
    });
    JSL_SYNCWARP();
    if (bad.any() && !c.raise) c.raise = JSL_ST_OTHER_INVALID;  // ZeroDivisionError
}

// TasselJsspObservation.make (observations.py:738-918): f32 [7][J] = allocatable | left_over_time |
// percent_finished | total_completion | time_until_next_machine_is_free | idle_since_last_op | cum_idle_time.
// Needs the per-op (start, end) log (cfg.record_ops is forced on for this observation kind).
template <int JW, bool ST, int CL>
JSL_D void write_obs_tassel(const EnvS<JW, ST, CL>& s, const Ctx& c, uint8_t* out) {
    const InstDev& I = *c.I;
    const int J = I.J;
    float* f = (float*)out;
    const double mat = (double)s.max_allowed_time;
    const int32_t* log = c.aux->op_log;
    const int now = s.time;
    const int total_bytes = obs_bytes_tassel(J);
    wp_for_dyn(1, [&](int) { for (int i = 7 * 4 * J; i < total_bytes; i++) out[i] = 0; });
    wp_for<JW>(J, [&](int j) {
        const int o0 = I.job_op_start[j], n = job_nops(I, j), k = s.j_k[j];
        const bool proc = s.j_proc[j] != 0;
        const int kidle = k + (proc ? 1 : 0);  // index of the first IDLE op
        f[0 * J + j] = (kidle < n && !proc) ? 1.f : 0.f;
        const int left = proc ? s.j_end[j] - now : 0;
        f[1 * J + j] = (float)left;
        f[2 * J + j] = (float)(n > 0 ? (double)k / (double)n : 0.0);
        long long remaining = left;
        for (int i = kidle; i < n; i++) remaining += op_duration<ST>(c, o0 + i);
        f[3 * J + j] = (float)((double)remaining / mat);
        int wait = 0;
        if (kidle < n) {
            int m = c.aux->op_mt[o0 + kidle] & 0xffff;
            if (s.m_state[m] == JSL_M_WORKING) wait = s.m_occ[m] - now;
        }
        f[4 * J + j] = (float)wait;
        const int last_end = k > 0 ? log[2 * (o0 + k - 1) + 1] : 0;
        f[5 * J + j] = (k > 0 && !proc) ? (float)((double)(now - last_end) / mat) : 0.f;
        long long total = 0;
        if (n > 0) {
            if (k > 0) total += log[2 * o0];                 // first op done: its logged start
            else if (proc) total += s.j_start[j];            // first op processing
        }
        for (int i = 0; i + 1 < n && i < k; i++) {           // op i DONE and op i+1 not IDLE
            if (i + 1 < k) total += log[2 * (o0 + i + 1)] - log[2 * (o0 + i) + 1];
            else if (proc) total += s.j_start[j] - log[2 * (o0 + i) + 1];
        }
        if (n > 0 && k == n && !proc) total += now - log[2 * (o0 + n - 1) + 1];
        f[6 * J + j] = (float)((double)total / mat);
    });
}

// ------------------------------------------------------------------------------------ canonical digest
// Layout documented in DESIGN.md / oracle/ref_harness.py.  Executed by one env's warp (diagnostics,
// parity tests, render()).  Returns the number of int32 produced (writes at most cap).
template <int JW, bool ST, int CL>
JSL_D int write_digest(const EnvS<JW, ST, CL>& s, const Ctx& c, int32_t* out, int cap) {
    const InstDev& I = *c.I;
    int n = 0;
    // executed redundantly by every lane with identical values (uniform stores)
#define JSL_PUT(v) do { if (n < cap) out[n] = (int32_t)(v); n++; } while (0)
    JSL_PUT(s.time); JSL_PUT(I.J); JSL_PUT(I.M); JSL_PUT(I.T); JSL_PUT(I.S);
    for (int j = 0; j < I.J; j++) {
        int k = s.j_k[j], nops = job_nops(I, j), o0 = I.job_op_start[j];
        JSL_PUT(k); JSL_PUT(s.j_proc[j]); JSL_PUT(s.j_loc[j]);
        for (int i = 0; i < nops; i++) {
            if (i < k) { JSL_PUT(2); JSL_PUT(c.aux->op_log ? c.aux->op_log[2 * (o0 + i)] : -1); JSL_PUT(c.aux->op_log ? c.aux->op_log[2 * (o0 + i) + 1] : -1); }
            else if (i == k && s.j_proc[j]) { JSL_PUT(1); JSL_PUT(s.j_start[j]); JSL_PUT(s.j_end[j]); }
            else { JSL_PUT(0); JSL_PUT(-1); JSL_PUT(-1); }
        }
    }
    auto put_store = [&](int b) {  // jobs of buffer b in stamp order
        int cnt = 0;
        for (int j = 0; j < I.J; j++) cnt += s.j_loc[j] == b;
        JSL_PUT(cnt);
        uint32_t last = 0; bool have = false;
        for (int i = 0; i < cnt; i++) {
            int best = -1;
            for (int j = 0; j < I.J; j++)
                if (s.j_loc[j] == b && (!have || s.j_seq[j] > last) && (best < 0 || s.j_seq[j] < s.j_seq[best])) best = j;
            JSL_PUT(best); last = s.j_seq[best]; have = true;
        }
    };
    for (int m = 0; m < I.M; m++) {
        JSL_PUT(s.m_state[m]); JSL_PUT(s.m_occ[m]); JSL_PUT(s.m_tool[m]);
        put_store(3 * m); put_store(3 * m + 2); put_store(3 * m + 1);
    }
    for (int t = 0; t < I.T; t++) {
        JSL_PUT(s.t_state[t]); JSL_PUT(s.t_occkind[t]);
        if (s.t_occkind[t] == OCC_TD) { JSL_PUT(s.t_occ[t]); JSL_PUT(s.t_tdbuf[t]); JSL_PUT(s.t_tdcomp[t]); JSL_PUT(s.t_tdjob[t]); }
        else if (s.t_occkind[t] == OCC_TIME) { JSL_PUT(s.t_occ[t]); JSL_PUT(-1); JSL_PUT(-1); JSL_PUT(-1); }
        else { JSL_PUT(-1); JSL_PUT(-1); JSL_PUT(-1); JSL_PUT(-1); }
        JSL_PUT(s.t_carry[t] == NONE8 ? -1 : s.t_carry[t]);
        JSL_PUT(s.t_job[t] == NONE8 ? -1 : s.t_job[t]);
        JSL_PUT(s.t_lockind[t]); JSL_PUT(s.t_loc0[t]);
        JSL_PUT(s.t_lockind[t] ? (int)s.t_loc1[t] : -1); JSL_PUT(s.t_lockind[t] ? (int)s.t_loc2[t] : -1);
    }
    for (int b = 3 * I.M + I.T; b < I.NB; b++) put_store(b);
    if (c.aux->outs) {
        const OutS<JW>& ou = *(const OutS<JW>*)c.aux->outs;
        for (int m = 0; m < I.M; m++) {
            int nm = I.has_m_out ? I.m_out_start[m + 1] - I.m_out_start[m] : 0;
            for (int k = 0; k < nm; k++) {
                int i = m * MAX_OUT + k;
                JSL_PUT(ou.mo_active[i]); JSL_PUT(ou.mo_a[i]); JSL_PUT(ou.mo_active[i] ? ou.mo_b[i] : -1);
            }
        }
        for (int t = 0; t < I.T; t++)
            for (int k = 0; k < (I.has_t_out ? I.n_t_out : 0); k++) {
                int i = t * MAX_OUT + k;
                JSL_PUT(ou.to_active[i]); JSL_PUT(ou.to_a[i]); JSL_PUT(ou.to_active[i] ? ou.to_b[i] : -1);
            }
    }
    int remaining = s.n_possible - s.skip;
    if (remaining < 0) remaining = 0;
    JSL_PUT(remaining);
    if (remaining > 0) {
        Poss<JW> p = cached_poss(s);
        for (int i = 0; i < remaining; i++) {
            Trans t = nth_possible(s, c, p, s.skip + i);
            JSL_PUT(t.kind); JSL_PUT(t.comp); JSL_PUT(t.new_state); JSL_PUT(t.job);
        }
    }
#undef JSL_PUT
    return n;
}

}  // namespace jsl
