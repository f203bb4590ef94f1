// jsl_kernels.cuh -- the sm_100a kernels around jsl_engine.cuh and their typed launchers.
//
// Kernels (one warp per env, 4 warps per CTA, state staged in shared memory):
//   k_reset   : env.reset for the masked envs, writes the HBM state image + first observation
//   k_step    : HBM state -> smem (128-bit coalesced), one env.step, obs/reward/flags out, smem -> HBM
//   k_rollout : persistent grid; each warp runs whole episodes with an in-kernel policy, state
//               never leaves shared memory between steps
//   k_digest  : canonical digest of one env (diagnostics / parity / render)
// Each lane-word count JW (1, 2, 4 -> up to 32, 64, 128 jobs/transports) is instantiated in its own
// translation unit (jsl_kernels_jw<N>.cu) so the three compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "jsl_engine.cuh"

namespace jsl {

#ifndef JSL_WARPS_PER_CTA
#define JSL_WARPS_PER_CTA 2
#endif
constexpr int WARPS_PER_CTA = JSL_WARPS_PER_CTA;
#ifndef JSL_ROLLOUT_MIN_CTAS
#define JSL_ROLLOUT_MIN_CTAS 32  // resident CTAs per SM the rollout kernel is register-budgeted for
#endif
#ifndef JSL_STEP_MIN_CTAS
#define JSL_STEP_MIN_CTAS 24
#endif
#ifndef JSL_ROLLOUT_MIN_CTAS_ST
#define JSL_ROLLOUT_MIN_CTAS_ST 32  // the stochastic rollout kernels (24 / 20 = 40 / 48 registers were measured slower)
#endif
#ifndef JSL_ROLLOUT_MIN_CTAS_GENERAL
#define JSL_ROLLOUT_MIN_CTAS_GENERAL 32  // same budget for the general (non-classic) rollout kernels
#endif

struct BatchDev {
    InstDev inst;
    CfgDev cfg;
    int64_t B;
    uint8_t* state;      // [B][state_stride]
    uint8_t* outs;       // [B][outs_stride] or nullptr
    int32_t* op_log;     // [B][N][2] or nullptr
    int32_t* leg_log;    // [B][leg_stride] transport-task log (jsl_engine.cuh) or nullptr
    int leg_stride;
    int state_stride, outs_stride;
    uint64_t seed;
    int64_t env_base;    // global index of env 0 (RANDOM policy stream id)
    int32_t* tcur;       // [B][n_tobj] current .time of every time object (stochastic instances) or nullptr
    const int32_t* tape; // [B][tape_stride] replayed update() results or nullptr
    int64_t tape_stride;
    int classic;         // specialisation level CL of jsl_engine.cuh: 0 general, 1 simple, 2 classic, 3 queues, 4 features
    unsigned long long* work_counter;  // next env index of the persistent rollout grid (dynamic work queue)
    uint16_t* tseed;     // [B][n_tobj] numpy seed of every stochastic time object
    const int32_t* start_seeds;  // [B][n_tobj] or nullptr
};

// index of the warp inside its CTA, obtained through a warp reduction so that the compiler knows it (and
// every shared-memory address derived from it) to be warp-uniform
__device__ __forceinline__ int warp_index() { return (int)__reduce_min_sync(0xffffffffu, threadIdx.x >> 5); }

// fixed-size variant: all 128-bit loads of a lane are issued before the first store (latency overlap)
template <int BYTES>
__device__ __forceinline__ void copy_fixed(void* dst, const void* src) {
    constexpr int N = BYTES / 16, K = (N + 31) / 32;
    const int4* s = (const int4*)src;
    int4* d = (int4*)dst;
    int4 v[K];
    const int lane = JSL_LANE;
#pragma unroll
    for (int k = 0; k < K; k++)
        if (lane + 32 * k < N) v[k] = s[lane + 32 * k];
#pragma unroll
    for (int k = 0; k < K; k++)
        if (lane + 32 * k < N) d[lane + 32 * k] = v[k];
    __syncwarp();
}

// the two halves of copy_fixed, for work that does not depend on the data to be placed between them
template <int BYTES>
struct Staged {
    static constexpr int N = BYTES / 16, K = (N + 31) / 32;
    int4 v[K];
    __device__ __forceinline__ void load(const void* src) {
        const int4* s = (const int4*)src;
        const int lane = JSL_LANE;
#pragma unroll
        for (int k = 0; k < K; k++)
            if (lane + 32 * k < N) v[k] = s[lane + 32 * k];
    }
    __device__ __forceinline__ void store(void* dst) {
        int4* d = (int4*)dst;
        const int lane = JSL_LANE;
#pragma unroll
        for (int k = 0; k < K; k++)
            if (lane + 32 * k < N) d[lane + 32 * k] = v[k];
        __syncwarp();
    }
};

__device__ __forceinline__ void copy_in(void* dst, const void* src, int bytes) {
    const int4* s = (const int4*)src;
    int4* d = (int4*)dst;
    for (int i = JSL_LANE; i < bytes / 16; i += 32) d[i] = s[i];
    __syncwarp();
}

// per-warp shared-memory slot: EnvS | EnvAux | OutS (only for instances with outages).  The offsets inside the
// slot are compile-time constants, so every address is "slot base + immediate"; only the slot stride of the
// general kernels depends on the batch (a bigger slot would cost them L1 capacity for nothing).
template <int JW>
struct Slot {
    static constexpr int AUX = (int)sizeof(EnvS<JW, false, 0>);
    static constexpr int OUTS = AUX + (int)sizeof(EnvAux);
    static constexpr int BYTES = OUTS;                              // without outage bookkeeping
    static constexpr int BYTES_OUTS = OUTS + (int)sizeof(OutS<JW>);  // with it
};
inline int slot_bytes(int JW, bool has_outs) {
    return JW == 1 ? (has_outs ? Slot<1>::BYTES_OUTS : Slot<1>::BYTES)
         : JW == 2 ? (has_outs ? Slot<2>::BYTES_OUTS : Slot<2>::BYTES)
                   : (has_outs ? Slot<4>::BYTES_OUTS : Slot<4>::BYTES);
}
template <int JW, int CL>
__device__ __forceinline__ int slot_stride(const BatchDev& bd) {
    return (!JSL_LVL_OUT(CL) || !bd.outs) ? Slot<JW>::BYTES : Slot<JW>::BYTES_OUTS;
}

// Per-env context.  STEP = the one-step kernels (k_reset / k_step / k_digest): the fields only the fused rollout
// reads (policy streams, action tape, step limit) are not set up.  `env` is 32-bit (B < 2^31, checked at batch
// creation) so that every address is one IMAD.WIDE.
template <int JW, bool ST, int CL, bool STEP>
__device__ __forceinline__ void make_ctx(Ctx& c, const BatchDev& bd, uint32_t env, uint8_t* slot) {
    const InstDev& I = bd.inst;
    const uint32_t ni = (uint32_t)I.n_inst;
    const uint32_t v = ni == 1u ? 0u : (env < ni ? env : env % ni);
    EnvAux* aux = (EnvAux*)(slot + Slot<JW>::AUX);
    c.I = &bd.inst;
    c.C = &bd.cfg;
    c.aux = aux;
    c.raise = 0;
    c.now = 0;
    c.log = LEG_LOG_BUILD && bd.leg_log != nullptr;
    __syncwarp();
    // every lane stores the same (warp-uniform) values: no divergent single-lane section
    const size_t vo = (size_t)v * (uint32_t)I.N;
    aux->op_mt = I.op_mt + vo;
    aux->op_dur = I.op_dur + vo;
    aux->job_work = I.job_work + (size_t)v * (uint32_t)I.J;
    aux->lb_mat = I.lb_mat + 2 * (size_t)v;
    aux->op_log = bd.op_log ? bd.op_log + (size_t)env * (2u * (uint32_t)I.N) : nullptr;
    aux->leg_log = (LEG_LOG_BUILD && bd.leg_log) ? bd.leg_log + (size_t)env * (uint32_t)bd.leg_stride : nullptr;
    aux->outs = (JSL_LVL_OUT(CL) && bd.outs) ? (void*)(slot + Slot<JW>::OUTS) : nullptr;
    if (ST) {
        aux->tcur = bd.tcur + (size_t)env * (uint32_t)I.n_tobj;
        aux->tseed = bd.tseed ? bd.tseed + (size_t)env * (uint32_t)I.n_tobj : nullptr;
        aux->start_seeds = bd.start_seeds ? bd.start_seeds + (size_t)env * (uint32_t)I.n_tobj : nullptr;
        aux->tape = bd.tape ? bd.tape + (size_t)env * bd.tape_stride : nullptr;
        aux->tape_len = (int)bd.tape_stride;
    }
    if (ST || !STEP) {  // Philox key / stream: stochastic draws and the RANDOM policy
        aux->seed = bd.seed;
        aux->env_id = (uint64_t)(bd.env_base + env);
    }
    aux->neg_inv_n = I.neg_inv_n;
    if (!STEP) aux->act_block = 0;
    __syncwarp();
}

// per-env scalars of a step: one 32-byte jsl_step_record (two 16-byte stores) or, for callers that did not ask for
// it, the separate arrays
// `fin` < 0: a plain step.  Otherwise the env ended in this call and was reset in the same launch (JSL_STEP_AUTO_RESET):
// fin = terminated | truncated << 8 | status << 16 of the finished step, fin_mk its makespan; they are reported together
// with the time, status and candidate of the fresh episode.
template <int JW, bool ST, int CL>
__device__ __forceinline__ void write_scalars(const EnvS<JW, ST, CL>& s, const jsl_step_out& out, uint32_t env,
                                              double reward, int status, int cd0, int cd1, int cd2, int cd3,
                                              int fin = -1, int fin_mk = -1) {
    if (JSL_LANE != 0) return;
    const int mk = fin >= 0 ? fin_mk : (s.terminated ? s.time : -1);
    const uint32_t term = fin >= 0 ? (uint32_t)(fin & 0xff) : (uint32_t)s.terminated;
    const uint32_t trunc = fin >= 0 ? (uint32_t)((fin >> 8) & 0xff) : (uint32_t)s.truncated;
    if (out.record) {
        int4 a, b;
        a.x = __double2loint(reward); a.y = __double2hiint(reward);
        a.z = s.time; a.w = mk;
        b.x = (int)(((uint32_t)cd0 & 0xffffu) | ((uint32_t)cd1 << 16));
        b.y = (int)(((uint32_t)cd2 & 0xffffu) | ((uint32_t)cd3 << 16));
        b.z = (int)(term | (trunc << 8) | ((uint32_t)status << 16) | (fin >= 0 ? ((uint32_t)(fin >> 16) & 0xffu) << 24 : 0u));
        b.w = 0;
        int4* q = (int4*)(out.record + env);
        q[0] = a; q[1] = b;
        return;
    }
    if (out.reward) out.reward[env] = reward;
    if (out.terminated) out.terminated[env] = (uint8_t)term;
    if (out.truncated) out.truncated[env] = (uint8_t)trunc;
    if (out.status) out.status[env] = (uint8_t)status;
    if (out.time) out.time[env] = s.time;
    if (out.makespan) out.makespan[env] = mk;
    if (out.cand) {
        int16_t* q = out.cand + (size_t)env * 4;
        q[0] = (int16_t)cd0; q[1] = (int16_t)cd1; q[2] = (int16_t)cd2; q[3] = (int16_t)cd3;
    }
}

template <int JW, bool ST, int CL>
__device__ __forceinline__ void write_step_out(const EnvS<JW, ST, CL>& s, Ctx& c, const jsl_step_out& out, uint32_t env,
                                               double reward, int obs_bytes, int fin = -1, int fin_mk = -1) {
    const int remaining = s.n_possible - s.skip;
    // the candidate the agent sees next: needed by the observation (current_transition) and by the record
    Trans cand;
    cand.kind = -1; cand.comp = -1; cand.new_state = 0; cand.job = -1;
    const bool have_cand = remaining > 0 && s.status <= JSL_ST_STEP_FAILED;
    if (have_cand) cand = current_candidate(s, c);
    if (out.obs && s.status <= JSL_ST_STEP_FAILED) {
        uint8_t* o = out.obs + (size_t)env * (uint32_t)obs_bytes;
        if (c.C->obs_kind == JSL_OBS_BINARY_ACTION) write_obs_binary(s, c, o, have_cand, cand);
        else if (c.C->obs_kind == JSL_OBS_OPERATION_ARRAY) write_obs_oparray(s, c, o, have_cand, cand);
        else if (c.C->obs_kind == JSL_OBS_TASSEL) write_obs_tassel(s, c, o);
        else if (c.C->obs_kind == JSL_OBS_STATE) copy_fixed<Image<JW, CL>::BYTES>(o, &s);  // raw state for torch-side factories
    }
    int status = s.status;
    if (c.raise && status <= JSL_ST_STEP_FAILED) status = c.raise;  // observation factory raised
    const bool show = remaining > 0 && status <= JSL_ST_TRUNCATED;
    write_scalars(s, out, env, reward, status, show ? cand.kind : -1, show ? cand.comp : -1, show ? cand.job : -1,
                  show ? (remaining > 32767 ? 32767 : remaining) : 0, fin, fin_mk);
}

template <int JW, bool ST, int CL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_reset(const __grid_constant__ BatchDev bd, const uint8_t* __restrict__ mask, jsl_step_out out, int obs_bytes) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = WARPS_PER_CTA == 1 ? 0 : warp_index();
    const uint32_t env = blockIdx.x * WARPS_PER_CTA + warp;
    if (env >= (uint32_t)bd.B) return;
    if (mask && !mask[env]) return;
    uint8_t* slot = smem + warp * slot_stride<JW, CL>(bd);
    EnvS<JW, ST, CL>& s = *(EnvS<JW, ST, CL>*)slot;
    void* outs_sm = bd.outs ? (void*)(slot + Slot<JW>::OUTS) : nullptr;
    Ctx c;
    make_ctx<JW, ST, CL, true>(c, bd, env, slot);
    run_env_op(s, c, OP_RESET);
    __syncwarp();
    write_step_out(s, c, out, env, 0.0, obs_bytes);
    __syncwarp();
    copy_fixed<Image<JW, CL>::BYTES>(bd.state + (size_t)env * Image<JW, CL>::BYTES, &s);
    if (bd.outs) copy_in(bd.outs + (size_t)env * (uint32_t)bd.outs_stride, outs_sm, bd.outs_stride);
}

// AUTO (jsl_step_out.flags & JSL_STEP_AUTO_RESET): an env whose episode ends in this step -- terminated, truncated or
// dead with an escaped exception -- is reset in the same launch, the way a vector env auto-resets: reward, flags and
// makespan of the finished step, observation / time / status / candidate of the fresh episode.  A separate
// instantiation: the plain kernel's code is what it is without the flag.
template <int JW, bool ST, int CL, bool AUTO = false>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, JSL_STEP_MIN_CTAS)
k_step(const __grid_constant__ BatchDev bd, const uint8_t* __restrict__ actions, jsl_step_out out, int obs_bytes) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = WARPS_PER_CTA == 1 ? 0 : warp_index();
    const uint32_t env = blockIdx.x * WARPS_PER_CTA + warp;
    if (env >= (uint32_t)bd.B) return;
    const int action = actions[env];  // issued ahead of the image copy, whose latency covers it
    uint8_t* slot = smem + warp * slot_stride<JW, CL>(bd);
    EnvS<JW, ST, CL>& s = *(EnvS<JW, ST, CL>*)slot;
    constexpr int IMG = Image<JW, CL>::BYTES;
    const bool has_outs = JSL_LVL_OUT(CL) && bd.outs != nullptr;
    void* outs_sm = (void*)(slot + Slot<JW>::OUTS);
    uint8_t* g_state = bd.state + (size_t)env * IMG;
    uint8_t* g_outs = has_outs ? bd.outs + (size_t)env * (uint32_t)bd.outs_stride : nullptr;
    // The image travels by bulk copy (UBLKCP: the TMA engine writes the 1.5 KB straight into the slot and signals the
    // warp's mbarrier) instead of 128-bit loads and shared-memory stores of the lanes: the kernel's busiest unit is the
    // LSU data pipe, and the two image copies were a fifth of its wavefronts.  The context is set up under the copy.
    static_assert(IMG % 16 == 0, "bulk copies move multiples of 16 bytes");
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&((EnvAux*)(slot + Slot<JW>::AUX))->bar);
    const uint32_t s_sm = (uint32_t)__cvta_generic_to_shared(slot);
    if (JSL_LANE == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(IMG) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s_sm), "l"(g_state), "r"(IMG), "r"(bar) : "memory");
    }
    __syncwarp();
    if (has_outs) copy_in(outs_sm, g_outs, bd.outs_stride);
    Ctx c;
    make_ctx<JW, ST, CL, true>(c, bd, env, slot);
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    __syncwarp();
    double r = 0.0;
    int op = action, fin = -1, fin_mk = -1;
#pragma unroll 1
    for (;;) {  // (one call site of the engine: the second pass exists only in the AUTO instantiation)
    const double r_op = run_env_op(s, c, op);
    __syncwarp();
    if (AUTO && op == OP_RESET) break;
    r = r_op;
    if (s.status >= JSL_ST_STEP_FAILED) {
        // success=False / an escaping exception: the reference keeps the OLD state (state.py:202-210);
        // re-read the HBM image and carry over only the bookkeeping scalars.
        int status = s.status, n_steps = s.n_steps, reward_noop = s.reward_noop, step_action = s.step_action,
            step_noop = s.step_noop;
        double ret = s.ret;
        __syncwarp();
        copy_in(&s, g_state, IMG);
        if (has_outs) copy_in(outs_sm, g_outs, bd.outs_stride);
        s.status = status;
        if (status == JSL_ST_STEP_FAILED) {
            s.truncated = 1; s.terminated = 0; s.obs_done = 1;
            s.n_steps = n_steps; s.reward_noop = reward_noop; s.step_action = step_action; s.step_noop = step_noop;
            s.ret = ret;
        }
        __syncwarp();
    }
    if (!AUTO || !(s.terminated || s.truncated || s.status >= JSL_ST_STEP_FAILED)) break;
    fin = (int)s.terminated | ((int)s.truncated << 8) | (s.status << 16);
    fin_mk = s.terminated ? s.time : -1;
    op = OP_RESET; c.raise = 0;
    __syncwarp();
    }
    write_step_out(s, c, out, env, r, obs_bytes, fin, fin_mk);
    // generic-proxy writes of all lanes -> visible to the async proxy, then one bulk store of the image; the slot must
    // stay allocated until the engine has read it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (JSL_LANE == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g_state), "r"(s_sm), "r"(IMG) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (has_outs) copy_in(g_outs, outs_sm, bd.outs_stride);
    if (JSL_LANE == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

#ifdef JSL_STEP_PERSISTENT
// EXPERIMENT (not in the shipped library): persistent step kernel.  One resident wave of CTAs, every warp
// walks envs gw, gw + W, ...; the image of the NEXT env is fetched with cp.async (LDGSTS, no registers) into the warp's
// second slot while the current env is stepped.
template <int JW, bool ST, int CL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, JSL_STEP_MIN_CTAS)
k_step_p(const __grid_constant__ BatchDev bd, const uint8_t* __restrict__ actions, jsl_step_out out, int obs_bytes) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = WARPS_PER_CTA == 1 ? 0 : warp_index();
    const uint32_t B = (uint32_t)bd.B, nw = gridDim.x * WARPS_PER_CTA;
    uint32_t env = blockIdx.x * WARPS_PER_CTA + warp;
    if (env >= B) return;
    constexpr int IMG = Image<JW, CL>::BYTES;
    const int stride = slot_stride<JW, CL>(bd);
    uint8_t* base = smem + warp * 2 * stride;
    const bool has_outs = JSL_LVL_OUT(CL) && bd.outs != nullptr;
    auto prefetch = [&](uint8_t* dst, uint32_t e) {
        const uint8_t* src = bd.state + (size_t)e * IMG;
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
#pragma unroll
        for (int k = 0; k < (IMG / 16 + 31) / 32; k++) {
            const int i = JSL_LANE + 32 * k;
            if (i < IMG / 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16 * i), "l"(src + 16 * i));
        }
    };
    int cur = 0;
    prefetch(base, env);
    asm volatile("cp.async.commit_group;");
    for (;;) {
        const uint32_t nxt = env + nw;
        if (nxt < B) prefetch(base + (1 - cur) * stride, nxt);
        asm volatile("cp.async.commit_group;");
        const int action = actions[env];
        uint8_t* slot = base + cur * stride;
        EnvS<JW, ST, CL>& s = *(EnvS<JW, ST, CL>*)slot;
        void* outs_sm = (void*)(slot + Slot<JW>::OUTS);
        uint8_t* g_state = bd.state + (size_t)env * IMG;
        uint8_t* g_outs = has_outs ? bd.outs + (size_t)env * (uint32_t)bd.outs_stride : nullptr;
        Ctx c;
        make_ctx<JW, ST, CL, true>(c, bd, env, slot);
        if (has_outs) copy_in(outs_sm, g_outs, bd.outs_stride);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        double r = run_env_op(s, c, action);
        __syncwarp();
        if (s.status >= JSL_ST_STEP_FAILED) {
            int status = s.status, n_steps = s.n_steps, reward_noop = s.reward_noop, step_action = s.step_action,
                step_noop = s.step_noop;
            double ret = s.ret;
            __syncwarp();
            copy_in(&s, g_state, IMG);
            if (has_outs) copy_in(outs_sm, g_outs, bd.outs_stride);
            s.status = status;
            if (status == JSL_ST_STEP_FAILED) {
                s.truncated = 1; s.terminated = 0; s.obs_done = 1;
                s.n_steps = n_steps; s.reward_noop = reward_noop; s.step_action = step_action; s.step_noop = step_noop;
                s.ret = ret;
            }
            __syncwarp();
        }
        write_step_out(s, c, out, env, r, obs_bytes);
        __syncwarp();
        copy_fixed<IMG>(g_state, &s);
        if (has_outs) copy_in(g_outs, outs_sm, bd.outs_stride);
        if (nxt >= B) break;
        env = nxt; cur ^= 1;
    }
}
#endif

template <int JW, bool ST, int CL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, ST ? JSL_ROLLOUT_MIN_CTAS_ST : CL ? JSL_ROLLOUT_MIN_CTAS : JSL_ROLLOUT_MIN_CTAS_GENERAL)
k_rollout(const __grid_constant__ BatchDev bd, int policy, const uint8_t* __restrict__ tape, int64_t tape_stride,
          int max_steps, int flags, jsl_rollout_out out) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = WARPS_PER_CTA == 1 ? 0 : warp_index();
    uint8_t* slot = smem + warp * slot_stride<JW, CL>(bd);
    asm volatile("" : "+l"(slot));          // keep the slot address in registers instead of re-deriving it
    __builtin_assume(__isShared(slot));
    EnvS<JW, ST, CL>& s = *(EnvS<JW, ST, CL>*)slot;
    void* outs_sm = bd.outs ? (void*)(slot + Slot<JW>::OUTS) : nullptr;
    // persistent warps pull env indices from a global counter: episodes differ in length, a static
    // grid-stride split leaves the tail of the launch half empty
    for (;;) {
        unsigned long long next = 0;
        if (JSL_LANE == 0) next = atomicAdd(bd.work_counter, 1ull);
        const int64_t env = (int64_t)__shfl_sync(0xffffffffu, next, 0);
        if (env >= bd.B) break;
        Ctx c;
        make_ctx<JW, ST, CL, false>(c, bd, (uint32_t)env, slot);
        bool pending_reset = (flags & JSL_RO_RESET_FIRST) != 0;
        if (!pending_reset) {
            copy_in(&s, bd.state + env * (int64_t)Image<JW, CL>::BYTES, Image<JW, CL>::BYTES);
            if (bd.outs) copy_in(outs_sm, bd.outs + env * (int64_t)bd.outs_stride, bd.outs_stride);
        }
        if (JSL_LANE == 0) {  // per-env loop constants live in the aux block, not in registers
            c.aux->env = env;
            c.aux->action_tape = tape ? tape + env * tape_stride : nullptr;
            const int64_t lim = (int64_t)(pending_reset ? 0 : s.n_steps) + max_steps;
            c.aux->step_limit = lim > 0x7fffffff ? 0x7fffffff : (int32_t)lim;
        }
        __syncwarp();
#pragma unroll 1
        for (;;) {
            int op;
            if (pending_reset) {
                op = OP_RESET;
            } else {
                if (s.n_steps >= c.aux->step_limit || s.status != JSL_ST_OK) break;
                if (policy == JSL_POLICY_TAPE) {
                    if (s.n_steps >= tape_stride) break;
                    op = c.aux->action_tape[s.n_steps];
                } else {
                    op = policy_action(s, c, policy);
                }
            }
            run_env_op(s, c, op);
            pending_reset = false;
            __syncwarp();
        }
        const int64_t e = c.aux->env;
        if (JSL_LANE == 0) {
            if (out.makespan) out.makespan[e] = s.time;
            if (out.n_steps) out.n_steps[e] = s.n_steps;
            if (out.ret) out.ret[e] = s.ret;
            if (out.status) out.status[e] = (uint8_t)s.status;
            if (out.no_op_count) out.no_op_count[e] = s.noop_count;
        }
        __syncwarp();
        if (!(flags & JSL_RO_NO_WRITEBACK)) {
            copy_in(bd.state + e * (int64_t)Image<JW, CL>::BYTES, &s, Image<JW, CL>::BYTES);
            if (bd.outs) copy_in(bd.outs + e * (int64_t)bd.outs_stride, outs_sm, bd.outs_stride);
        }
        __syncwarp();
    }
}

template <int JW, bool ST, int CL>
__global__ void __launch_bounds__(32)
k_digest(const __grid_constant__ BatchDev bd, int64_t env, int32_t* out, int cap, int32_t* n_out) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* slot = smem;
    EnvS<JW, ST, CL>& s = *(EnvS<JW, ST, CL>*)smem;
    void* outs_sm = bd.outs ? (void*)(smem + Slot<JW>::OUTS) : nullptr;
    copy_in(&s, bd.state + env * (int64_t)Image<JW, CL>::BYTES, Image<JW, CL>::BYTES);
    if (bd.outs) copy_in(outs_sm, bd.outs + env * (int64_t)bd.outs_stride, bd.outs_stride);
    Ctx c;
    make_ctx<JW, ST, CL, true>(c, bd, (uint32_t)env, slot);
    int n = write_digest(s, c, out, cap);
    if (JSL_LANE == 0) *n_out = n;
}

// ------------------------------------------------------------------------------------ launchers
struct LaunchCfg {
    int grid, smem;
    cudaStream_t stream;
    int ctas_per_sm = 0;  // resident CTAs per SM the launch is sized for (0: unknown) -> shared-memory carve-out hint
};

template <int JW, bool ST, int CL>
cudaError_t launch_reset(const BatchDev& bd, const LaunchCfg& lc, const uint8_t* mask, jsl_step_out out, int obs_bytes) {
    if (lc.smem > 48 * 1024) cudaFuncSetAttribute(k_reset<JW, ST, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, lc.smem);
    k_reset<JW, ST, CL><<<lc.grid, WARPS_PER_CTA * 32, lc.smem, lc.stream>>>(bd, mask, out, obs_bytes);
    return cudaGetLastError();
}
template <int JW, bool ST, int CL, bool AUTO>
cudaError_t launch_step_impl(const BatchDev& bd, const LaunchCfg& lc, const uint8_t* actions, jsl_step_out out, int obs_bytes) {
    if (lc.smem > 48 * 1024) cudaFuncSetAttribute(k_step<JW, ST, CL, AUTO>, cudaFuncAttributeMaxDynamicSharedMemorySize, lc.smem);
#ifdef JSL_STEP_PERSISTENT
    {
        static int sms = 0;
        if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
        const int smem2 = 2 * lc.smem;
        static bool once = false;
        if (!once) {
            cudaFuncSetAttribute(k_step_p<JW, ST, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
            cudaFuncSetAttribute(k_step_p<JW, ST, CL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            once = true;
        }
        int ctas = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_step_p<JW, ST, CL>, WARPS_PER_CTA * 32, smem2);
        const long long wave = (long long)sms * (ctas > 0 ? ctas : 1);
        k_step_p<JW, ST, CL><<<(int)(lc.grid < wave ? lc.grid : wave), WARPS_PER_CTA * 32, smem2, lc.stream>>>(bd, actions, out, obs_bytes);
        return cudaGetLastError();
    }
#endif
    static bool carved = false;  // per instantiation: room for JSL_STEP_MIN_CTAS resident CTAs' slots, the rest stays L1
    if (!carved) {
        int pct = (int)(((long long)lc.smem + 1024) * JSL_STEP_MIN_CTAS * 100 / (228 * 1024)) + 1;
        cudaFuncSetAttribute(k_step<JW, ST, CL, AUTO>, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
        carved = true;
    }
    k_step<JW, ST, CL, AUTO><<<lc.grid, WARPS_PER_CTA * 32, lc.smem, lc.stream>>>(bd, actions, out, obs_bytes);
    return cudaGetLastError();
}
template <int JW, bool ST, int CL>
cudaError_t launch_step(const BatchDev& bd, const LaunchCfg& lc, const uint8_t* actions, jsl_step_out out, int obs_bytes) {
    return (out.flags & JSL_STEP_AUTO_RESET) ? launch_step_impl<JW, ST, CL, true>(bd, lc, actions, out, obs_bytes)
                                             : launch_step_impl<JW, ST, CL, false>(bd, lc, actions, out, obs_bytes);
}
template <int JW, bool ST, int CL>
cudaError_t launch_rollout(const BatchDev& bd, const LaunchCfg& lc, int policy, const uint8_t* tape, int64_t tape_stride,
                           int max_steps, int flags, jsl_rollout_out out) {
    if (lc.smem > 48 * 1024) cudaFuncSetAttribute(k_rollout<JW, ST, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, lc.smem);
    static int carved = 0;  // per instantiation: shared memory for the resident CTAs' slots, the rest stays L1 (local memory)
    if (lc.ctas_per_sm > 0 && carved != lc.ctas_per_sm) {
        int pct = (int)(((long long)lc.smem + 1024) * lc.ctas_per_sm * 100 / (228 * 1024)) + 1;
        cudaFuncSetAttribute(k_rollout<JW, ST, CL>, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
        carved = lc.ctas_per_sm;
    }
    k_rollout<JW, ST, CL><<<lc.grid, WARPS_PER_CTA * 32, lc.smem, lc.stream>>>(bd, policy, tape, tape_stride, max_steps, flags, out);
    return cudaGetLastError();
}
template <int JW, bool ST, int CL>
cudaError_t launch_digest(const BatchDev& bd, const LaunchCfg& lc, int64_t env, int32_t* out, int cap, int32_t* n_out) {
    k_digest<JW, ST, CL><<<1, 32, lc.smem, lc.stream>>>(bd, env, out, cap, n_out);
    return cudaGetLastError();
}
template <int JW, bool ST, int CL>
int rollout_occupancy(int smem) {  // resident CTAs per SM for the persistent rollout grid
    int n = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_rollout<JW, ST, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_rollout<JW, ST, CL>, WARPS_PER_CTA * 32, smem) != cudaSuccess) return 8;
    return n > 0 ? n : 1;
}

// field table of the state image (jsl_batch_state_layout)
template <int JW>
int state_layout(int level, jsl_state_field* out, int cap) {
    using S = EnvS<JW>;
    const int limit = JSL_LVL_BUF(level) ? Image<JW, 0>::BYTES : Image<JW, 1>::BYTES;
    int n = 0;
    auto put = [&](const char* name, size_t off, int eb, int cnt, int flt) {
        if ((int)off >= limit) return;
        if (n < cap) {
            jsl_state_field& f = out[n];
            int k = 0;
            for (; name[k] && k < 23; k++) f.name[k] = name[k];
            for (; k < 24; k++) f.name[k] = 0;
            f.offset = (int32_t)off; f.elem_bytes = eb; f.count = cnt; f.is_float = flt;
        }
        n++;
    };
#define JSL_F(field, flt) put(#field, offsetof(S, field), (int)sizeof(((S*)0)->field), 1, flt)
#define JSL_A(field) put(#field, offsetof(S, field), (int)sizeof(((S*)0)->field[0]), (int)(sizeof(((S*)0)->field) / sizeof(((S*)0)->field[0])), 0)
    JSL_F(time, 0); JSL_F(seq_ctr, 0); JSL_F(n_possible, 0); JSL_F(skip, 0); JSL_F(joker, 0); JSL_F(step_noop, 0);
    JSL_F(step_action, 0); JSL_F(reward_noop, 0); JSL_F(n_steps, 0); JSL_F(noop_count, 0); JSL_F(max_done_end, 0);
    JSL_F(status, 0); JSL_F(terminated, 0); JSL_F(truncated, 0); JSL_F(last_action_empty, 0); JSL_F(obs_done, 0);
    JSL_F(all_out, 0); JSL_F(draw_ctr, 0); JSL_F(lower_bound, 0); JSL_F(ret, 1); JSL_F(inv_mat, 1); JSL_F(max_allowed_time, 0);
    JSL_A(c_p1); JSL_A(c_idle); JSL_A(c_lr); JSL_A(c_li);
    JSL_A(j_start); JSL_A(j_end); JSL_A(j_seq); JSL_A(j_loc); JSL_A(j_k); JSL_A(j_proc); JSL_A(j_agv); JSL_A(j_mach); JSL_A(j_exec);
    JSL_A(m_occ); JSL_A(m_done); JSL_A(m_state); JSL_A(m_tool); JSL_A(m_job);
    JSL_A(t_occ); JSL_A(t_loc1); JSL_A(t_state); JSL_A(t_job); JSL_A(t_carry); JSL_A(t_occkind); JSL_A(t_lockind); JSL_A(t_loc0); JSL_A(t_loc2);
    JSL_A(t_tdbuf); JSL_A(t_tdcomp); JSL_A(t_tdjob);
#undef JSL_F
#undef JSL_A
    return n;
}

// entry points defined by jsl_kernels_jw<N>.cu
#define JSL_DECLARE_JW(N)                                                                                             \
    cudaError_t launch_reset_jw##N(const BatchDev&, const LaunchCfg&, const uint8_t*, jsl_step_out, int);             \
    cudaError_t launch_step_jw##N(const BatchDev&, const LaunchCfg&, const uint8_t*, jsl_step_out, int);              \
    cudaError_t launch_rollout_jw##N(const BatchDev&, const LaunchCfg&, int, const uint8_t*, int64_t, int, int,      \
                                     jsl_rollout_out);                                                                \
    cudaError_t launch_digest_jw##N(const BatchDev&, const LaunchCfg&, int64_t, int32_t*, int, int32_t*);            \
    int rollout_occupancy_jw##N(const BatchDev& bd, int smem);                                                                            \
    int state_bytes_jw##N(int level);                                                                                 \
    int state_layout_jw##N(int level, jsl_state_field* out, int cap);                                                 \
    int outs_bytes_jw##N();
JSL_DECLARE_JW(1)
JSL_DECLARE_JW(2)
JSL_DECLARE_JW(4)

// one instantiation per (lane words, kind): stochastic (general | features) | classic (level 2) | simple (level 1) |
// queues (level 3) | features (level 4) | general
#define JSL_BY_KIND(bd, F, N, ...)                                                                \
    ((bd).tcur           ? ((bd).classic == 4 ? F<N, true, 4>(__VA_ARGS__) : F<N, true, 0>(__VA_ARGS__)) \
     : (bd).classic == 2 ? F<N, false, 2>(__VA_ARGS__)                                            \
     : (bd).classic == 1 ? F<N, false, 1>(__VA_ARGS__)                                            \
     : (bd).classic == 3 ? F<N, false, 3>(__VA_ARGS__)                                            \
     : (bd).classic == 4 ? F<N, false, 4>(__VA_ARGS__)                                            \
                         : F<N, false, 0>(__VA_ARGS__))
#define JSL_DEFINE_JW(N)                                                                                              \
    cudaError_t launch_reset_jw##N(const BatchDev& bd, const LaunchCfg& lc, const uint8_t* m, jsl_step_out o, int ob) { \
        return JSL_BY_KIND(bd, launch_reset, N, bd, lc, m, o, ob);                                                    \
    }                                                                                                                 \
    cudaError_t launch_step_jw##N(const BatchDev& bd, const LaunchCfg& lc, const uint8_t* a, jsl_step_out o, int ob) { \
        return JSL_BY_KIND(bd, launch_step, N, bd, lc, a, o, ob);                                                     \
    }                                                                                                                 \
    cudaError_t launch_rollout_jw##N(const BatchDev& bd, const LaunchCfg& lc, int p, const uint8_t* t, int64_t ts,    \
                                     int ms, int f, jsl_rollout_out o) {                                              \
        return JSL_BY_KIND(bd, launch_rollout, N, bd, lc, p, t, ts, ms, f, o);                                        \
    }                                                                                                                 \
    cudaError_t launch_digest_jw##N(const BatchDev& bd, const LaunchCfg& lc, int64_t e, int32_t* o, int c, int32_t* n) { \
        return JSL_BY_KIND(bd, launch_digest, N, bd, lc, e, o, c, n);                                                 \
    }                                                                                                                 \
    int rollout_occupancy_jw##N(const BatchDev& bd, int smem) { return JSL_BY_KIND(bd, rollout_occupancy, N, smem); } \
    int state_bytes_jw##N(int level) { return JSL_LVL_BUF(level) ? Image<N, 0>::BYTES : Image<N, 1>::BYTES; }          \
    int state_layout_jw##N(int level, jsl_state_field* out, int cap) { return state_layout<N>(level, out, cap); }      \
    int outs_bytes_jw##N() { return (int)sizeof(OutS<N>); }

}  // namespace jsl
