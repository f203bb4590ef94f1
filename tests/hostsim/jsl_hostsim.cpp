// jsl_hostsim.cpp -- CPU-side DEBUG build of the CUDA engine source (jobshoplab_b200/csrc/jsl_engine.cuh).
//
// TEST TOOLING ONLY.  There is no GPU in the build container, so this compiles the very same engine
// header with its three warp primitives written as loops (JSL_HOST_SIM) and lets the CPU test-suite
// drive the kernel logic against the oracle before GPU time is spent.  It is built and loaded only by
// tests/test_engine_hostsim.py; the product package never loads it and has no CPU execution path.
#define JSL_HOST_SIM 1
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cmath>
#include "../../jobshoplab_b200/csrc/jsl_engine.cuh"
#include "../../jobshoplab_b200/csrc/jsl_randomize.cuh"
#include "../../jobshoplab_b200/csrc/jsl_tobj.hpp"

using namespace jsl;

namespace {

struct Sim {
    InstDev I;
    CfgDev C;
    int JW;
    std::vector<int32_t> job_op_start, op_mt, op_dur, job_work, lb_mat, pre_cap, post_cap, setup, sbuf_cap, buf_ref_id,
        travel, m_out_start, m_out_freq, m_out_dur, t_out_freq, t_out_dur;
    std::vector<uint8_t> pre_type, post_type, sbuf_type, sbuf_role, agv_init_place, job_lex, mach_lex;
    std::vector<uint16_t> job_init_buf;
    std::vector<float> tr_comp, tr_job;
    std::vector<int32_t> op_log, leg_log;
    std::vector<uint8_t> state, backup, outs, outs_backup;
    int state_bytes, outs_bytes;
    uint64_t seed;
    double last_reward;
    TobjTables tobj;
    std::vector<int32_t> tcur, tape, start_seeds;
    std::vector<uint16_t> tseed;
    uint64_t env_id = 0;
    EnvAux aux;
    int classic = 0;  // specialisation level, as in jsl_abi.cu
};

template <class T, class U>
std::vector<T> conv(const U* p, size_t n) {
    std::vector<T> v(n);
    for (size_t i = 0; i < n; i++) v[i] = (T)(p ? p[i] : 0);
    return v;
}
bool all_eq(const std::vector<int32_t>& v, int x) { for (int a : v) if (a != x) return false; return true; }
template <class T> bool all_eq8(const std::vector<T>& v, int x) { for (auto a : v) if ((int)a != x) return false; return true; }

void lex_order(const char* pfx, int n, std::vector<uint8_t>& out) {
    std::vector<std::string> ids(n);
    for (int i = 0; i < n; i++) ids[i] = std::string(pfx) + "-" + std::to_string(i);
    out.resize(n);
    for (int i = 0; i < n; i++) out[i] = (uint8_t)i;
    for (int i = 1; i < n; i++) {
        uint8_t v = out[i]; int k = i - 1;
        while (k >= 0 && ids[out[k]] > ids[v]) { out[k + 1] = out[k]; k--; }
        out[k + 1] = v;
    }
}

void make_ctx(Sim* s, Ctx& c) {
    EnvAux& a = s->aux;
    c.I = &s->I; c.C = &s->C; a.op_mt = s->op_mt.data(); a.op_dur = s->op_dur.data(); a.job_work = s->job_work.data();
    c.aux = &a; c.raise = 0; c.now = 0;
    a.op_log = s->C.record_ops ? s->op_log.data() : nullptr;
    a.leg_log = s->C.record_ops ? s->leg_log.data() : nullptr;
    c.log = a.leg_log != nullptr;
    a.outs = s->outs_bytes ? (void*)s->outs.data() : nullptr;
    a.tcur = s->tobj.stochastic ? s->tcur.data() : nullptr;
    a.tseed = s->tobj.stochastic ? s->tseed.data() : nullptr;
    a.start_seeds = s->start_seeds.empty() ? nullptr : s->start_seeds.data();
    a.tape = s->tape.empty() ? nullptr : s->tape.data(); a.tape_len = (int)s->tape.size();
    a.seed = s->seed; a.env_id = s->env_id; a.act_block = 0;
    a.lb_mat = s->lb_mat.data(); a.neg_inv_n = -1.0 / (double)s->I.N;
}

// Only the first Image<JW, CL>::BYTES of the state travel to HBM between the launches of step mode: poison the rest
// after every operation, so that any dependence of the engine on scratch surviving a launch shows up here.
template <int JW, bool ST, int CL> void scrub(Sim* s) {
    std::memset(s->state.data() + Image<JW, CL>::BYTES, 0xA5, sizeof(EnvS<JW, ST, CL>) - Image<JW, CL>::BYTES);
}
template <int JW, bool ST, int CL> int do_reset(Sim* s) {
    EnvS<JW, ST, CL>& st = *(EnvS<JW, ST, CL>*)s->state.data();
    Ctx c; make_ctx(s, c);
    run_env_op(st, c, OP_RESET);
    s->last_reward = 0.0;
    scrub<JW, ST, CL>(s);
    return st.status;
}
template <int JW, bool ST, int CL> int do_step(Sim* s, int action) {
    EnvS<JW, ST, CL>& st = *(EnvS<JW, ST, CL>*)s->state.data();
    s->backup = s->state; s->outs_backup = s->outs;
    Ctx c; make_ctx(s, c);
    double r = run_env_op(st, c, action);
    if (st.status >= JSL_ST_STEP_FAILED) {  // same rollback as k_step
        int status = st.status, n_steps = st.n_steps, rn = st.reward_noop, sa = st.step_action, sn = st.step_noop;
        double ret = st.ret;
        s->state = s->backup; s->outs = s->outs_backup;
        EnvS<JW, ST, CL>& o = *(EnvS<JW, ST, CL>*)s->state.data();
        o.status = status;
        if (status == JSL_ST_STEP_FAILED) { o.truncated = 1; o.terminated = 0; o.obs_done = 1; o.n_steps = n_steps; o.reward_noop = rn; o.step_action = sa; o.step_noop = sn; o.ret = ret; }
    }
    s->last_reward = r;
    scrub<JW, ST, CL>(s);
    return ((EnvS<JW, ST, CL>*)s->state.data())->status;
}
template <int JW, bool ST, int CL> int do_digest(Sim* s, int32_t* out, int cap) {
    EnvS<JW, ST, CL>& st = *(EnvS<JW, ST, CL>*)s->state.data();
    Ctx c; make_ctx(s, c);
    return write_digest(st, c, out, cap);
}
template <int JW, bool ST, int CL> int do_obs(Sim* s, uint8_t* out) {
    EnvS<JW, ST, CL>& st = *(EnvS<JW, ST, CL>*)s->state.data();
    Ctx c; make_ctx(s, c);
    if (s->C.obs_kind == JSL_OBS_OPERATION_ARRAY) { { Trans none = {-1, -1, 0, -1}; write_obs_oparray(st, c, out, false, none); } return c.raise ? -c.raise : obs_bytes_oparray(s->I.J, s->I.N); }
    if (s->C.obs_kind == JSL_OBS_TASSEL) { write_obs_tassel(st, c, out); return obs_bytes_tassel(s->I.J); }
    { Trans none = {-1, -1, 0, -1}; write_obs_binary(st, c, out, false, none); }
    return obs_bytes_binary(s->I.J, s->I.M);
}
template <int JW, bool ST, int CL> int do_rollout(Sim* s, int policy, const uint8_t* tape, int n_tape, int max_steps, uint64_t env_idx) {
    EnvS<JW, ST, CL>& st = *(EnvS<JW, ST, CL>*)s->state.data();
    Ctx c; make_ctx(s, c);
    int n = 0;
    while (n < max_steps && st.status == JSL_ST_OK) {
        int a;
        if (policy == JSL_POLICY_TAPE) { if (st.n_steps >= n_tape) break; a = tape[st.n_steps]; }
        else { c.aux->env_id = env_idx; a = policy_action(st, c, policy); }
        run_env_op(st, c, a);
        n++;
    }
    scrub<JW, ST, CL>(s);
    return n;
}
template <int JW, bool ST, int CL> void get_scalars(Sim* s, int32_t* o) {
    EnvS<JW, ST, CL>& st = *(EnvS<JW, ST, CL>*)s->state.data();
    o[0] = st.time; o[1] = st.terminated; o[2] = st.truncated; o[3] = st.status; o[4] = st.n_steps; o[5] = st.noop_count;
    o[6] = st.n_possible - st.skip;
}
template <int JW, bool ST, int CL> double get_ret(Sim* s) { return ((EnvS<JW, ST, CL>*)s->state.data())->ret; }

}  // namespace

#define DISPATCH1(s, F, N, ...) ((s)->tobj.stochastic ? ((s)->classic == 4 ? F<N, true, 4>(__VA_ARGS__) : F<N, true, 0>(__VA_ARGS__)) : (s)->classic == 4 ? F<N, false, 4>(__VA_ARGS__) : (s)->classic == 2 ? F<N, false, 2>(__VA_ARGS__) : (s)->classic == 1 ? F<N, false, 1>(__VA_ARGS__) : (s)->classic == 3 ? F<N, false, 3>(__VA_ARGS__) : F<N, false, 0>(__VA_ARGS__))
#define DISPATCH(s, F, ...) ((s)->JW == 1 ? DISPATCH1(s, F, 1, __VA_ARGS__) : (s)->JW == 2 ? DISPATCH1(s, F, 2, __VA_ARGS__) : DISPATCH1(s, F, 4, __VA_ARGS__))

extern "C" {

void* jsh_create(const jsl_instance_desc* d, const jsl_env_cfg* cfg, uint64_t seed) {
    Sim* s = new Sim();
    InstDev& I = s->I;
    std::memset(&I, 0, sizeof(I));
    I.J = d->n_jobs; I.M = d->n_machines; I.T = d->n_transports; I.S = d->n_sbuf; I.NT = d->n_tools; I.N = d->n_ops;
    I.NB = 3 * I.M + I.T + I.S; I.NP = I.M + I.S; I.n_inst = 1;
    int jt = I.J > I.T ? I.J : I.T;
    if (I.M > 32 || jt > 128) { delete s; return nullptr; }
    s->JW = jt <= 32 ? 1 : jt <= 64 ? 2 : 4;
    s->job_op_start = conv<int32_t>(d->job_op_start, I.J + 1);
    s->op_mt.resize(I.N); s->op_dur = conv<int32_t>(d->op_duration, I.N);
    for (int o = 0; o < I.N; o++) s->op_mt[o] = d->op_machine[o] | (d->op_tool[o] << 16);
    s->job_work.assign(I.J, 0);
    for (int j = 0; j < I.J; j++) for (int o = s->job_op_start[j]; o < s->job_op_start[j + 1]; o++) s->job_work[j] += s->op_dur[o];
    s->lb_mat = {d->lower_bound, d->max_allowed_time};
    s->pre_type = conv<uint8_t>(d->pre_type, I.M); s->pre_cap = conv<int32_t>(d->pre_cap, I.M);
    s->post_type = conv<uint8_t>(d->post_type, I.M); s->post_cap = conv<int32_t>(d->post_cap, I.M);
    s->setup = conv<int32_t>(d->setup_time, (size_t)I.M * I.NT * I.NT);
    s->sbuf_type = conv<uint8_t>(d->sbuf_type, I.S); s->sbuf_cap = conv<int32_t>(d->sbuf_cap, I.S);
    s->sbuf_role = conv<uint8_t>(d->sbuf_role, I.S); s->buf_ref_id = conv<int32_t>(d->buf_ref_id, I.NB);
    s->travel = conv<int32_t>(d->travel_time, (size_t)I.NP * I.NP);
    s->agv_init_place = conv<uint8_t>(d->agv_init_place, I.T); s->job_init_buf = conv<uint16_t>(d->job_init_buf, I.J);
    s->m_out_start = conv<int32_t>(d->m_outage_start, I.M + 1);
    int n_mo = d->m_outage_start ? d->m_outage_start[I.M] : 0;
    s->m_out_freq = conv<int32_t>(d->m_outage_freq, n_mo); s->m_out_dur = conv<int32_t>(d->m_outage_dur, n_mo);
    s->t_out_freq = conv<int32_t>(d->t_outage_freq, d->n_t_outages); s->t_out_dur = conv<int32_t>(d->t_outage_dur, d->n_t_outages);
    lex_order("j", I.J, s->job_lex); lex_order("m", I.M, s->mach_lex);
    I.job_op_start = s->job_op_start.data(); I.op_mt = s->op_mt.data(); I.op_dur = s->op_dur.data();
    I.job_work = s->job_work.data(); I.lb_mat = s->lb_mat.data();
    I.pre_type = s->pre_type.data(); I.pre_cap = s->pre_cap.data(); I.post_type = s->post_type.data(); I.post_cap = s->post_cap.data();
    I.setup = all_eq(s->setup, 0) ? nullptr : s->setup.data();
    I.sbuf_type = s->sbuf_type.data(); I.sbuf_cap = s->sbuf_cap.data(); I.sbuf_role = s->sbuf_role.data();
    I.buf_ref_id = s->buf_ref_id.data();
    I.travel = all_eq(s->travel, 0) ? nullptr : s->travel.data();
    I.agv_init_place = s->agv_init_place.data(); I.job_init_buf = s->job_init_buf.data();
    I.m_out_start = s->m_out_start.data(); I.m_out_freq = s->m_out_freq.data(); I.m_out_dur = s->m_out_dur.data();
    I.t_out_freq = s->t_out_freq.data(); I.t_out_dur = s->t_out_dur.data();
    s->tr_comp.resize(I.M + I.T); s->tr_job.resize(I.J + 1);
    for (int i = 0; i < I.M + I.T; i++) s->tr_comp[i] = (float)((double)i / (double)(I.M + I.T));
    for (int i = 0; i <= I.J; i++) s->tr_job[i] = (float)((double)i / (double)I.J);
    I.tr_comp = s->tr_comp.data(); I.tr_job = s->tr_job.data();
    I.job_lex = s->job_lex.data(); I.mach_lex = s->mach_lex.data();
    I.n_t_out = d->n_t_outages; I.has_m_out = n_mo > 0; I.has_t_out = I.n_t_out > 0;
    I.out_buf = -1;
    for (int b = 0; b < I.S; b++) if (s->sbuf_role[b] == JSL_ROLE_OUTPUT) { I.out_buf = 3 * I.M + I.T + b; break; }
    I.default_tool = d->default_tool; I.start_time = d->start_time;
    I.sb0 = 3 * I.M + I.T; I.out_mask = 0;
    for (int b = 0; b < I.S && b < 32; b++) if (s->sbuf_role[b] == JSL_ROLE_OUTPUT) I.out_mask |= 1u << b;
    I.flex_pre = all_eq8(s->pre_type, JSL_BUF_FLEX);
    I.flex_post = all_eq8(s->post_type, JSL_BUF_FLEX) && all_eq8(s->sbuf_type, JSL_BUF_FLEX);
    I.caps_inf = all_eq(s->pre_cap, JSL_CAP_INF) && all_eq(s->post_cap, JSL_CAP_INF) && all_eq(s->sbuf_cap, JSL_CAP_INF);
    CfgDev& C = s->C;
    C.allow_early_transport = cfg->allow_early_transport; C.truncation_joker = cfg->truncation_joker;
    C.truncation_active = cfg->truncation_active; C.obs_kind = cfg->obs_kind; C.record_ops = cfg->record_ops;
    C.sparse_bias = cfg->sparse_bias; C.dense_bias = cfg->dense_bias; C.truncation_bias = cfg->truncation_bias;
    s->state_bytes = s->JW == 1 ? sizeof(EnvS<1, false, 0>) : s->JW == 2 ? sizeof(EnvS<2, false, 0>) : sizeof(EnvS<4, false, 0>);
    s->outs_bytes = (I.has_m_out || I.has_t_out) ? (s->JW == 1 ? sizeof(OutS<1>) : s->JW == 2 ? sizeof(OutS<2>) : sizeof(OutS<4>)) : 0;
    s->state.assign(s->state_bytes + 16, 0); s->outs.assign(s->outs_bytes + 16, 0);
    s->op_log.assign((size_t)2 * I.N + 2, -1);
    s->leg_log.assign((size_t)2 + 32 * s->JW + 5 * (I.N + 2 * I.J + 8) + 8, 0);
    s->seed = seed; s->last_reward = 0.0;
    s->tobj = make_tobj(*d);
    if (s->tobj.stochastic) {
        s->tcur.assign(s->tobj.n + 1, 0); s->tseed.assign(s->tobj.n + 1, 0);
        if (s->tobj.sample_depth > 0) { I.sample_table = s->tobj.sample_table.data(); I.sample_row = s->tobj.sample_row.data(); I.sample_depth = s->tobj.sample_depth; }
        I.tobj_kind = s->tobj.kind.data(); I.tobj_base = s->tobj.base.data(); I.tobj_param = s->tobj.param.data();
        I.tobj_init = s->tobj.init.data(); I.n_tobj = s->tobj.n; I.off_setup = s->tobj.off_setup;
        I.off_travel = s->tobj.off_travel; I.off_mof = s->tobj.off_mof; I.off_mod = s->tobj.off_mod;
        I.off_tof = s->tobj.off_tof; I.off_tod = s->tobj.off_tod;
    }
    {
        // same rule as jsl_abi.cu
        const bool plain = !I.has_m_out && !I.has_t_out && !I.setup && !s->tobj.stochastic && I.out_buf >= 0;
        const bool flexbuf = I.flex_pre && I.flex_post && I.caps_inf && I.out_buf >= 0;
        const bool simple = plain && flexbuf;
        s->classic = simple ? (I.travel ? 1 : 2) : plain ? 3 : flexbuf ? 4 : 0;
        if (const char* lvl = std::getenv("JSL_HOSTSIM_MAX_LEVEL")) {  // force a less specialised level (tests)
            const int lv = std::atoi(lvl);
            if (s->classic >= 3) { if (lv == 0) s->classic = 0; }
            else if (s->classic > lv) s->classic = lv < 0 ? 0 : lv;
        }
    }
    return s;
}
void jsh_destroy(void* h) { delete (Sim*)h; }
void jsh_set_fuse(int mask) { g_fuse = mask; }
void jsh_fused(long* o) { for (int i = 0; i < 5; i++) o[i] = g_fused[i]; }
void jsh_set_tape(void* h, const int32_t* t, int n) { Sim* s = (Sim*)h; s->tape.assign(t, t + n); }
void jsh_invariants(long* o) { o[0] = g_invariant_checks; o[1] = g_invariant_violations; }
void jsh_randomize_variant(uint64_t seed, uint64_t variant, int J, const int32_t* job_op_start, int min_d, int max_d,
                           int32_t* op_mt, int32_t* op_dur, int32_t* job_work, int32_t* lb_mat) {
    randomize_variant(seed, variant, J, job_op_start, min_d, max_d, op_mt, op_dur, job_work, lb_mat);
}
int jsh_leg_log(void* h, int32_t* out, int cap) {
    Sim* s = (Sim*)h;
    const int n = s->leg_log[0], hdr = 2 + 32 * s->JW;
    if (n > cap) return -1;
    std::memcpy(out, s->leg_log.data() + hdr, (size_t)n * 5 * sizeof(int32_t));
    return s->leg_log[1] ? -4 : n;
}
void jsh_fast_paths(long* o) { o[0] = g_fast_start; o[1] = g_fast_complete; o[2] = g_general_batches; }
void jsh_set_env_id(void* h, uint64_t id) { ((Sim*)h)->env_id = id; }
void jsh_set_start_seeds(void* h, const int32_t* t, int n) { Sim* s = (Sim*)h; s->start_seeds.assign(t, t + n); }
int jsh_draw(uint64_t seed, uint64_t env, uint32_t n, int kind, int base, float param) { return draw_time(seed, env, n, kind, base, param); }
int jsh_reset(void* h) { Sim* s = (Sim*)h; return DISPATCH(s, do_reset, s); }
int jsh_step(void* h, int a) { Sim* s = (Sim*)h; return DISPATCH(s, do_step, s, a); }
int jsh_digest(void* h, int32_t* out, int cap) { Sim* s = (Sim*)h; return DISPATCH(s, do_digest, s, out, cap); }
int jsh_obs(void* h, uint8_t* out) { Sim* s = (Sim*)h; return DISPATCH(s, do_obs, s, out); }
int jsh_rollout(void* h, int policy, const uint8_t* tape, int n_tape, int max_steps, uint64_t env_idx) {
    Sim* s = (Sim*)h; return DISPATCH(s, do_rollout, s, policy, tape, n_tape, max_steps, env_idx);
}
void jsh_scalars(void* h, int32_t* o) { Sim* s = (Sim*)h; DISPATCH(s, get_scalars, s, o); }
double jsh_reward(void* h) { return ((Sim*)h)->last_reward; }
double jsh_return(void* h) { Sim* s = (Sim*)h; return DISPATCH(s, get_ret, s); }
int jsh_state_bytes(void* h) { return ((Sim*)h)->state_bytes + ((Sim*)h)->outs_bytes; }

}  // extern "C"
