// jsl_abi.cu -- host side of the C ABI (include/jobshoplab_b200.h): instance / batch objects, device
// table upload and kernel dispatch.  The kernels live in jsl_kernels.cuh (one TU per lane-word count).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "jsl_kernels.cuh"
#include "jsl_randomize.cuh"
#include "jsl_tobj.hpp"

using namespace jsl;

namespace {

thread_local std::string g_err;
void set_err(const std::string& s) { g_err = s; }

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                 \
            return JSL_E_CUDA;                                                           \
        }                                                                                \
    } while (0)

// ------------------------------------------------------------------------------------ host side
template <class T>
T* dev_upload(const std::vector<T>& v, cudaError_t& err) {
    if (v.empty()) return nullptr;
    T* p = nullptr;
    err = cudaMalloc(&p, v.size() * sizeof(T));
    if (err != cudaSuccess) return nullptr;
    err = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return p;
}

}  // namespace

struct jsl_instance {
    int J, M, T, S, NT, N;
    std::vector<int32_t> job_op_start, op_machine, op_tool, op_duration, pre_type, pre_cap, post_type, post_cap,
        setup_time, sbuf_type, sbuf_cap, sbuf_role, buf_ref_id, travel_time, agv_init_place, job_init_buf,
        m_outage_start, m_outage_freq, m_outage_dur, t_outage_freq, t_outage_dur;
    int default_tool, start_time, lower_bound, max_allowed_time;
    bool stochastic;
    TobjTables tobj;
};

struct jsl_batch {
    BatchDev bd;
    int JW;
    int device;
    int obs_bytes;
    int digest_cap;
    std::vector<void*> allocs;
    int32_t* d_digest = nullptr;
    int32_t* d_digest_n = nullptr;
    int64_t launches = 0;
    int sm_count = 148;
    int smem_per_warp = 0;
    int rollout_ctas_per_sm = 0;
    bool ops_static = true;       // every operation duration is a static time object
    int32_t* d_bad = nullptr;     // number of out-of-range table entries seen by k_pack_variants (jsl_batch_check)
    int32_t* d_stage = nullptr;   // [2][n_inst*N] + [2][n_inst] staging for jsl_batch_load_variants
    bool stage_packed = false;    // the staged tables are u8 machines / u16 durations (jsl_batch_stage_variants_packed)
    int32_t* d_op_tool = nullptr; // [N]
};

namespace {

// packs staged (machine, duration) tables into the engine layout and recomputes the per-job work sums
// (uploaded tables are untrusted input: an operation machine outside [0, M) would index the per-machine state out
// of bounds inside the engine, so it is replaced by machine 0 and counted; jsl_batch_check reports it)
template <class TM, class TD>
__global__ void k_pack_variants(int64_t n_inst, int N, int J, int M, int32_t* bad, const int32_t* __restrict__ job_op_start,
                                const int32_t* __restrict__ op_tool, const TM* __restrict__ st_m,
                                const TD* __restrict__ st_d, const int32_t* __restrict__ st_lb,
                                const int32_t* __restrict__ st_mat, int32_t* op_mt, int32_t* op_dur,
                                int32_t* job_work, int32_t* lb_mat) {
    const int64_t total = n_inst * N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int o = (int)(i % N);
        int m = (int)st_m[i], d = (int)st_d[i];
        if (m < 0 || m >= M || d < 0) { atomicAdd(bad, 1); m = m < 0 || m >= M ? 0 : m; d = d < 0 ? 0 : d; }
        op_mt[i] = m | (op_tool[o] << 16);
        op_dur[i] = d;
    }
    const int64_t nj = n_inst * J;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nj; i += stride) {
        int64_t v = i / J;
        int j = (int)(i % J);
        int w = 0;
        for (int o = job_op_start[j]; o < job_op_start[j + 1]; o++) w += (int)st_d[v * N + o] < 0 ? 0 : (int)st_d[v * N + o];
        job_work[i] = w;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_inst; i += stride) {
        lb_mat[2 * i] = st_lb[i];
        lb_mat[2 * i + 1] = st_mat[i];
    }
}

// InstanceRandomizer for all variants: one thread per variant runs the sequential routine of jsl_randomize.cuh
__global__ void k_randomize(int64_t n_inst, int N, int J, const int32_t* __restrict__ job_op_start, uint64_t seed,
                            int64_t base, int min_d, int max_d, const uint8_t* __restrict__ mask, int32_t* op_mt,
                            int32_t* op_dur, int32_t* job_work, int32_t* lb_mat) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_inst; v += stride)
        if (!mask || mask[v])
            randomize_variant(seed, (uint64_t)(base + v), J, job_op_start, min_d, max_d, op_mt + v * N, op_dur + v * N,
                          job_work + v * J, lb_mat + 2 * v);
}

bool spec_is_static(const jsl_time_spec& sp, int n) {
    if (!sp.kind) return true;
    for (int i = 0; i < n; i++)
        if (sp.kind[i] != JSL_TIME_STATIC) return false;
    return true;
}

void lex_order(const char* pfx, int n, std::vector<uint8_t>& out) {
    std::vector<std::string> ids(n);
    for (int i = 0; i < n; i++) ids[i] = std::string(pfx) + "-" + std::to_string(i);
    out.resize(n);
    for (int i = 0; i < n; i++) out[i] = (uint8_t)i;
    for (int i = 1; i < n; i++) {
        uint8_t v = out[i];
        int k = i - 1;
        while (k >= 0 && ids[out[k]] > ids[v]) { out[k + 1] = out[k]; k--; }
        out[k + 1] = v;
    }
}

template <class T>
int upload(jsl_batch* b, const std::vector<T>& v, const T** dst) {
    cudaError_t err = cudaSuccess;
    T* p = dev_upload(v, err);
    if (err != cudaSuccess) { set_err(std::string("upload: ") + cudaGetErrorString(err)); return JSL_E_CUDA; }
    if (p) b->allocs.push_back(p);
    *dst = p;
    return JSL_OK;
}

template <class T, class U>
std::vector<T> narrow(const std::vector<U>& v) {
    std::vector<T> o(v.size());
    for (size_t i = 0; i < v.size(); i++) o[i] = (T)v[i];
    return o;
}

bool all_equal(const std::vector<int32_t>& v, int32_t x) {
    for (int32_t a : v) if (a != x) return false;
    return true;
}

// shared tail of jsl_batch_create / jsl_batch_create_variants
jsl_batch_t* build_batch(const jsl_instance* t, int64_t n_inst, const int32_t* op_machine, const int32_t* op_duration,
                         const int32_t* lbs, const int32_t* mats, int64_t B, int device, uint64_t seed,
                         const jsl_env_cfg* cfg) {
    if (B <= 0 || n_inst <= 0) { set_err("B and n_inst must be positive"); return nullptr; }
    if (B >= (1ll << 31) || n_inst >= (1ll << 31)) { set_err("unsupported: 2^31 or more envs / variants in one batch"); return nullptr; }
    if (t->M > 32) { set_err("unsupported: more than 32 machines"); return nullptr; }
    if (t->S > 32) { set_err("unsupported: more than 32 standalone buffers"); return nullptr; }
    int jt = t->J > t->T ? t->J : t->T;
    if (jt > 128 || t->J > 250 || t->T > 250) { set_err("unsupported: more than 128 jobs/transports"); return nullptr; }
    if (3 * t->M + t->T + t->S > 65000) { set_err("unsupported: too many buffers"); return nullptr; }
    for (int j = 0; j < t->J; j++)
        if (t->job_op_start[j + 1] - t->job_op_start[j] > 250) { set_err("unsupported: more than 250 ops per job"); return nullptr; }
    if (t->stochastic && n_inst != 1) {
        // several variants may share the stochastic setup / travel / outage objects as long as the operation durations
        // themselves are static (what InstanceRandomizer produces: DeterministicTimeConfig, manipulators.py:133-137)
        for (int o = 0; o < t->N; o++)
            if (t->tobj.kind[o] != JSL_TIME_STATIC) {
                set_err("unsupported: batches of several instances need static operation durations (stochastic setup / "
                        "travel / outage times are fine)");
                return nullptr;
            }
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); return nullptr; }
    jsl_batch* b = new jsl_batch();
    b->device = device;
    b->JW = jt <= 32 ? 1 : jt <= 64 ? 2 : 4;
    InstDev& I = b->bd.inst;
    std::memset(&b->bd, 0, sizeof(b->bd));
    I.J = t->J; I.M = t->M; I.T = t->T; I.S = t->S; I.NT = t->NT; I.N = t->N;
    I.NB = 3 * t->M + t->T + t->S; I.NP = t->M + t->S;
    I.n_inst = (int)n_inst;
    const int N = t->N, J = t->J;
    std::vector<int32_t> op_mt((size_t)n_inst * N), op_dur((size_t)n_inst * N), job_work((size_t)n_inst * J),
        lb_mat((size_t)n_inst * 2);
    for (int64_t v = 0; v < n_inst; v++) {
        const int32_t* om = op_machine + v * N;
        const int32_t* od = op_duration + v * N;
        for (int o = 0; o < N; o++) {
            if (om[o] < 0 || om[o] >= t->M) { set_err("operation machine out of range"); delete b; return nullptr; }
            op_mt[v * N + o] = om[o] | (t->op_tool[o] << 16);
            op_dur[v * N + o] = od[o];
        }
        for (int j = 0; j < J; j++) {
            int64_t w = 0;
            for (int o = t->job_op_start[j]; o < t->job_op_start[j + 1]; o++) w += od[o];
            job_work[v * J + j] = (int32_t)w;
        }
        lb_mat[2 * v] = lbs[v]; lb_mat[2 * v + 1] = mats[v];
    }
    int rc = 0;
    rc |= upload(b, t->job_op_start, &I.job_op_start);
    rc |= upload(b, op_mt, &I.op_mt);
    rc |= upload(b, op_dur, &I.op_dur);
    rc |= upload(b, job_work, &I.job_work);
    rc |= upload(b, lb_mat, &I.lb_mat);
    rc |= upload(b, narrow<uint8_t>(t->pre_type), &I.pre_type);
    rc |= upload(b, t->pre_cap, &I.pre_cap);
    rc |= upload(b, narrow<uint8_t>(t->post_type), &I.post_type);
    rc |= upload(b, t->post_cap, &I.post_cap);
    if (!all_equal(t->setup_time, 0)) rc |= upload(b, t->setup_time, &I.setup);
    rc |= upload(b, narrow<uint8_t>(t->sbuf_type), &I.sbuf_type);
    rc |= upload(b, t->sbuf_cap, &I.sbuf_cap);
    rc |= upload(b, narrow<uint8_t>(t->sbuf_role), &I.sbuf_role);
    rc |= upload(b, t->buf_ref_id, &I.buf_ref_id);
    if (!all_equal(t->travel_time, 0)) rc |= upload(b, t->travel_time, &I.travel);
    rc |= upload(b, narrow<uint8_t>(t->agv_init_place), &I.agv_init_place);
    rc |= upload(b, narrow<uint16_t>(t->job_init_buf), &I.job_init_buf);
    rc |= upload(b, t->m_outage_start, &I.m_out_start);
    rc |= upload(b, t->m_outage_freq, &I.m_out_freq);
    rc |= upload(b, t->m_outage_dur, &I.m_out_dur);
    rc |= upload(b, t->t_outage_freq, &I.t_out_freq);
    rc |= upload(b, t->t_outage_dur, &I.t_out_dur);
    std::vector<uint8_t> jl, ml;
    lex_order("j", t->J, jl);
    lex_order("m", t->M, ml);
    { const int32_t* tmp = nullptr; rc |= upload(b, t->op_tool, &tmp); b->d_op_tool = (int32_t*)tmp; }
    if (t->stochastic) {
        for (int o = 0; o < t->N; o++) b->ops_static = b->ops_static && t->tobj.kind[o] == JSL_TIME_STATIC;
        rc |= upload(b, t->tobj.kind, &I.tobj_kind);
        rc |= upload(b, t->tobj.base, &I.tobj_base);
        rc |= upload(b, t->tobj.param, &I.tobj_param);
        rc |= upload(b, t->tobj.init, &I.tobj_init);
        I.n_tobj = t->tobj.n; I.off_setup = t->tobj.off_setup; I.off_travel = t->tobj.off_travel;
        I.off_mof = t->tobj.off_mof; I.off_mod = t->tobj.off_mod; I.off_tof = t->tobj.off_tof; I.off_tod = t->tobj.off_tod;
        if (t->tobj.sample_depth > 0) {
            rc |= upload(b, t->tobj.sample_table, &I.sample_table);
            rc |= upload(b, t->tobj.sample_row, &I.sample_row);
            I.sample_depth = t->tobj.sample_depth;
        }
    }
    {
        std::vector<float> tc(t->M + t->T), tj(t->J + 1);
        for (int i = 0; i < t->M + t->T; i++) tc[i] = (float)((double)i / (double)(t->M + t->T));
        for (int i = 0; i <= t->J; i++) tj[i] = (float)((double)i / (double)t->J);
        rc |= upload(b, tc, &I.tr_comp);
        rc |= upload(b, tj, &I.tr_job);
    }
    rc |= upload(b, jl, &I.job_lex);
    rc |= upload(b, ml, &I.mach_lex);
    if (rc) { jsl_batch_destroy(b); return nullptr; }
    I.n_t_out = (int)t->t_outage_freq.size();
    I.has_m_out = !t->m_outage_freq.empty();
    I.has_t_out = I.n_t_out > 0;
    for (int m = 0; m < t->M; m++)
        if (t->m_outage_start[m + 1] - t->m_outage_start[m] > MAX_OUT) { set_err("unsupported: more than 2 outages per machine"); jsl_batch_destroy(b); return nullptr; }
    if (I.n_t_out > MAX_OUT) { set_err("unsupported: more than 2 transport outages"); jsl_batch_destroy(b); return nullptr; }
    I.out_buf = -1;
    for (int s = 0; s < t->S; s++)
        if (t->sbuf_role[s] == JSL_ROLE_OUTPUT) { I.out_buf = 3 * t->M + t->T + s; break; }
    I.default_tool = t->default_tool;
    I.start_time = t->start_time;
    I.sb0 = 3 * t->M + t->T;
    I.neg_inv_n = -1.0 / (double)I.N;
    I.out_mask = 0;
    for (int sb = 0; sb < t->S; sb++)
        if (t->sbuf_role[sb] == JSL_ROLE_OUTPUT) I.out_mask |= 1u << sb;
    I.flex_pre = all_equal(t->pre_type, JSL_BUF_FLEX);
    I.flex_post = all_equal(t->post_type, JSL_BUF_FLEX) && all_equal(t->sbuf_type, JSL_BUF_FLEX);
    I.caps_inf = all_equal(t->pre_cap, JSL_CAP_INF) && all_equal(t->post_cap, JSL_CAP_INF) && all_equal(t->sbuf_cap, JSL_CAP_INF);
    // specialisation level (jsl_engine.cuh): which features the instance needs compiled in
    const bool plain = !I.has_m_out && !I.has_t_out && !I.setup && !t->stochastic && I.out_buf >= 0;
    const bool flexbuf = I.flex_pre && I.flex_post && I.caps_inf && I.out_buf >= 0;
    const bool simple = plain && flexbuf;
    b->bd.classic = simple ? (I.travel ? 1 : 2) : plain ? 3 : flexbuf ? 4 : 0;
    if (const char* cap = std::getenv("JSL_MAX_LEVEL")) {  // diagnostics: force a less specialised kernel
        const int lv = std::atoi(cap);
        if (b->bd.classic >= 3) { if (lv == 0) b->bd.classic = 0; }
        else if (b->bd.classic > lv) b->bd.classic = lv < 0 ? 0 : lv;
    }
    CfgDev& C = b->bd.cfg;
    C.allow_early_transport = cfg->allow_early_transport; C.truncation_joker = cfg->truncation_joker;
    C.truncation_active = cfg->truncation_active; C.obs_kind = cfg->obs_kind; C.record_ops = cfg->record_ops;
    C.sparse_bias = cfg->sparse_bias; C.dense_bias = cfg->dense_bias; C.truncation_bias = cfg->truncation_bias;
    b->bd.B = B; b->bd.seed = seed; b->bd.env_base = 0;
    const int lvl = b->bd.classic;
    int ss = b->JW == 1 ? state_bytes_jw1(lvl) : b->JW == 2 ? state_bytes_jw2(lvl) : state_bytes_jw4(lvl);
    int os = b->JW == 1 ? outs_bytes_jw1() : b->JW == 2 ? outs_bytes_jw2() : outs_bytes_jw4();
    b->bd.state_stride = ss;
    b->bd.outs_stride = (I.has_m_out || I.has_t_out) ? os : 0;
    b->smem_per_warp = slot_bytes(b->JW, b->bd.outs_stride != 0);
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)B * ss) != cudaSuccess) { set_err("cudaMalloc(state) failed"); jsl_batch_destroy(b); return nullptr; }
    b->allocs.push_back(p); b->bd.state = (uint8_t*)p;
    cudaMemset(p, 0, (size_t)B * ss);
    if (b->bd.outs_stride) {
        if (cudaMalloc(&p, (size_t)B * os) != cudaSuccess) { set_err("cudaMalloc(outs) failed"); jsl_batch_destroy(b); return nullptr; }
        b->allocs.push_back(p); b->bd.outs = (uint8_t*)p;
    }
    if (t->stochastic) {
        if (cudaMalloc(&p, (size_t)B * t->tobj.n * sizeof(int32_t)) != cudaSuccess) { set_err("cudaMalloc(tcur) failed"); jsl_batch_destroy(b); return nullptr; }
        b->allocs.push_back(p); b->bd.tcur = (int32_t*)p;
        if (cudaMalloc(&p, (size_t)B * t->tobj.n * sizeof(uint16_t)) != cudaSuccess) { set_err("cudaMalloc(tseed) failed"); jsl_batch_destroy(b); return nullptr; }
        b->allocs.push_back(p); b->bd.tseed = (uint16_t*)p;
        cudaMemset(p, 0, (size_t)B * t->tobj.n * sizeof(uint16_t));
    }
    if (cfg->record_ops || cfg->obs_kind == JSL_OBS_TASSEL) {
        C.record_ops = 1;
        if (cudaMalloc(&p, (size_t)B * N * 2 * sizeof(int32_t)) != cudaSuccess) { set_err("cudaMalloc(op_log) failed"); jsl_batch_destroy(b); return nullptr; }
        b->allocs.push_back(p); b->bd.op_log = (int32_t*)p;
    }
    if (cfg->record_ops && LEG_LOG_BUILD) {  // (not for the op log the Tassel observation needs on big batches)
        b->bd.leg_stride = 2 + 32 * b->JW + 5 * (N + 2 * I.J + 8);
        if (cudaMalloc(&p, (size_t)B * b->bd.leg_stride * sizeof(int32_t)) != cudaSuccess) { set_err("cudaMalloc(leg_log) failed"); jsl_batch_destroy(b); return nullptr; }
        b->allocs.push_back(p); b->bd.leg_log = (int32_t*)p;
        cudaMemset(p, 0, (size_t)B * b->bd.leg_stride * sizeof(int32_t));
    }
    b->obs_bytes = cfg->obs_kind == JSL_OBS_BINARY_ACTION ? obs_bytes_binary(t->J, t->M)
                   : cfg->obs_kind == JSL_OBS_OPERATION_ARRAY ? obs_bytes_oparray(t->J, t->N)
                   : cfg->obs_kind == JSL_OBS_TASSEL ? obs_bytes_tassel(t->J)
                   : cfg->obs_kind == JSL_OBS_STATE ? ss : 0;
    b->digest_cap = 64 + 3 * t->J + 3 * N + t->M * 6 + 2 * t->J + 11 * t->T + t->S +
                    3 * ((int)t->m_outage_freq.size() + t->T * I.n_t_out) + 4 * (t->J + t->T * t->J);
    if (cudaMalloc(&p, (size_t)(b->digest_cap + 4) * sizeof(int32_t)) != cudaSuccess) { set_err("cudaMalloc(digest) failed"); jsl_batch_destroy(b); return nullptr; }
    b->allocs.push_back(p); b->d_digest = (int32_t*)p; b->d_digest_n = b->d_digest + b->digest_cap;
    if (cudaMalloc(&p, 2 * sizeof(unsigned long long)) != cudaSuccess) { set_err("cudaMalloc(counter) failed"); jsl_batch_destroy(b); return nullptr; }
    b->allocs.push_back(p); b->bd.work_counter = (unsigned long long*)p;
    b->d_bad = (int32_t*)((unsigned long long*)p + 1);
    cudaMemset(p, 0, 2 * sizeof(unsigned long long));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) b->sm_count = prop.multiProcessorCount;
    return b;
}

}  // namespace

extern "C" {

const char* jsl_last_error(void) { return g_err.c_str(); }
int jsl_abi_version(void) { return JSL_ABI_VERSION; }

jsl_instance_t* jsl_instance_create(const jsl_instance_desc* d) {
    if (!d || d->abi_version != JSL_ABI_VERSION) { set_err("bad descriptor / ABI version"); return nullptr; }
    if (d->n_jobs <= 0 || d->n_machines <= 0 || d->n_ops <= 0 || d->n_tools <= 0) { set_err("empty instance"); return nullptr; }
    jsl_instance* t = new jsl_instance();
    t->J = d->n_jobs; t->M = d->n_machines; t->T = d->n_transports; t->S = d->n_sbuf; t->NT = d->n_tools; t->N = d->n_ops;
    const int NB = 3 * t->M + t->T + t->S, NP = t->M + t->S;
    auto cp = [](std::vector<int32_t>& v, const int32_t* p, size_t n) { v.assign(n, 0); if (p && n) std::memcpy(v.data(), p, n * sizeof(int32_t)); };
    cp(t->job_op_start, d->job_op_start, t->J + 1);
    cp(t->op_machine, d->op_machine, t->N); cp(t->op_tool, d->op_tool, t->N); cp(t->op_duration, d->op_duration, t->N);
    cp(t->pre_type, d->pre_type, t->M); cp(t->pre_cap, d->pre_cap, t->M);
    cp(t->post_type, d->post_type, t->M); cp(t->post_cap, d->post_cap, t->M);
    cp(t->setup_time, d->setup_time, (size_t)t->M * t->NT * t->NT);
    cp(t->sbuf_type, d->sbuf_type, t->S); cp(t->sbuf_cap, d->sbuf_cap, t->S); cp(t->sbuf_role, d->sbuf_role, t->S);
    cp(t->buf_ref_id, d->buf_ref_id, NB);
    cp(t->travel_time, d->travel_time, (size_t)NP * NP);
    cp(t->agv_init_place, d->agv_init_place, t->T); cp(t->job_init_buf, d->job_init_buf, t->J);
    cp(t->m_outage_start, d->m_outage_start, t->M + 1);
    const int n_mo = d->m_outage_start ? d->m_outage_start[t->M] : 0;
    cp(t->m_outage_freq, d->m_outage_freq, n_mo); cp(t->m_outage_dur, d->m_outage_dur, n_mo);
    cp(t->t_outage_freq, d->t_outage_freq, d->n_t_outages); cp(t->t_outage_dur, d->t_outage_dur, d->n_t_outages);
    t->default_tool = d->default_tool; t->start_time = d->start_time;
    t->lower_bound = d->lower_bound; t->max_allowed_time = d->max_allowed_time;
    t->tobj = make_tobj(*d);
    t->stochastic = !(spec_is_static(d->op_spec, t->N) && spec_is_static(d->setup_spec, t->M * t->NT * t->NT) &&
                      spec_is_static(d->travel_spec, NP * NP) && spec_is_static(d->m_ofreq_spec, n_mo) &&
                      spec_is_static(d->m_odur_spec, n_mo) && spec_is_static(d->t_ofreq_spec, d->n_t_outages) &&
                      spec_is_static(d->t_odur_spec, d->n_t_outages));
    if (t->job_op_start[t->J] != t->N) { set_err("job_op_start does not cover n_ops"); delete t; return nullptr; }
    for (int j = 0; j < t->J; j++)
        if (t->job_init_buf[j] < 0 || t->job_init_buf[j] >= NB) { set_err("job_init_buf out of range"); delete t; return nullptr; }
    for (int a = 0; a < t->T; a++)
        if (t->agv_init_place[a] < 0 || t->agv_init_place[a] >= NP) { set_err("agv_init_place out of range"); delete t; return nullptr; }
    return t;
}

void jsl_instance_destroy(jsl_instance_t* t) { delete t; }

jsl_batch_t* jsl_batch_create(jsl_instance_t* const* insts, int n_inst, int64_t B, int device, uint64_t seed,
                              const jsl_env_cfg* cfg) {
    if (!insts || n_inst <= 0 || !cfg) { set_err("bad arguments"); return nullptr; }
    const jsl_instance* t = insts[0];
    std::vector<int32_t> om((size_t)n_inst * t->N), od((size_t)n_inst * t->N), lb(n_inst), mat(n_inst);
    for (int i = 0; i < n_inst; i++) {
        const jsl_instance* u = insts[i];
        bool same = u->J == t->J && u->M == t->M && u->T == t->T && u->S == t->S && u->NT == t->NT && u->N == t->N &&
                    u->job_op_start == t->job_op_start && u->op_tool == t->op_tool && u->pre_type == t->pre_type &&
                    u->pre_cap == t->pre_cap && u->post_type == t->post_type && u->post_cap == t->post_cap &&
                    u->setup_time == t->setup_time && u->sbuf_type == t->sbuf_type && u->sbuf_cap == t->sbuf_cap &&
                    u->sbuf_role == t->sbuf_role && u->travel_time == t->travel_time &&
                    u->agv_init_place == t->agv_init_place && u->job_init_buf == t->job_init_buf &&
                    u->m_outage_start == t->m_outage_start && u->m_outage_freq == t->m_outage_freq &&
                    u->m_outage_dur == t->m_outage_dur && u->t_outage_freq == t->t_outage_freq &&
                    u->t_outage_dur == t->t_outage_dur && u->default_tool == t->default_tool &&
                    u->start_time == t->start_time && u->stochastic == t->stochastic &&
                    (!t->stochastic || (u->tobj.kind == t->tobj.kind && u->tobj.param == t->tobj.param &&
                                        std::equal(u->tobj.base.begin() + t->N, u->tobj.base.end(), t->tobj.base.begin() + t->N) &&
                                        u->tobj.sample_row.size() == t->tobj.sample_row.size() &&
                                        std::equal(u->tobj.sample_row.begin() + t->N, u->tobj.sample_row.end(), t->tobj.sample_row.begin() + t->N) &&
                                        u->tobj.sample_table == t->tobj.sample_table));
        if (!same && i > 0) { set_err("instances of one batch must share everything but the operation tables"); return nullptr; }
        std::memcpy(&om[(size_t)i * t->N], u->op_machine.data(), t->N * sizeof(int32_t));
        std::memcpy(&od[(size_t)i * t->N], u->op_duration.data(), t->N * sizeof(int32_t));
        lb[i] = u->lower_bound; mat[i] = u->max_allowed_time;
    }
    return build_batch(t, n_inst, om.data(), od.data(), lb.data(), mat.data(), B, device, seed, cfg);
}

jsl_batch_t* jsl_batch_create_variants(const jsl_instance_t* tmpl, int64_t n_inst, const int32_t* op_machine,
                                       const int32_t* op_duration, const int32_t* lower_bound,
                                       const int32_t* max_allowed_time, int64_t B, int device, uint64_t seed,
                                       const jsl_env_cfg* cfg) {
    if (!tmpl || !op_machine || !op_duration || !lower_bound || !max_allowed_time || !cfg) { set_err("bad arguments"); return nullptr; }
    return build_batch(tmpl, n_inst, op_machine, op_duration, lower_bound, max_allowed_time, B, device, seed, cfg);
}

// stage: host -> device staging buffers (overlaps a running rollout: the live tables are not touched)
int jsl_batch_stage_variants(jsl_batch_t* b, const int32_t* op_machine, const int32_t* op_duration,
                             const int32_t* lower_bound, const int32_t* max_allowed_time, void* stream) {
    if (!b || !op_machine || !op_duration || !lower_bound || !max_allowed_time) return JSL_E_INVALID;
    CK(cudaSetDevice(b->device));
    const InstDev& I = b->bd.inst;
    const size_t n = (size_t)I.n_inst * I.N, ni = (size_t)I.n_inst;
    if (!b->d_stage) {
        void* p = nullptr;
        CK(cudaMalloc(&p, (2 * n + 2 * ni) * sizeof(int32_t)));
        b->allocs.push_back(p);
        b->d_stage = (int32_t*)p;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *sm = b->d_stage, *sd = sm + n, *sl = sd + n, *sa = sl + ni;
    CK(cudaMemcpyAsync(sm, op_machine, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sd, op_duration, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sl, lower_bound, ni * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sa, max_allowed_time, ni * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    b->stage_packed = false;
    return JSL_OK;
}

// the same with narrow host tables (machines < 256, durations < 65536): 3 instead of 8 bytes per operation over PCIe
int jsl_batch_stage_variants_packed(jsl_batch_t* b, const uint8_t* op_machine, const uint16_t* op_duration,
                                    const int32_t* lower_bound, const int32_t* max_allowed_time, void* stream) {
    if (!b || !op_machine || !op_duration || !lower_bound || !max_allowed_time) return JSL_E_INVALID;
    CK(cudaSetDevice(b->device));
    const InstDev& I = b->bd.inst;
    const size_t n = (size_t)I.n_inst * I.N, ni = (size_t)I.n_inst;
    if (!b->d_stage) {
        void* p = nullptr;
        CK(cudaMalloc(&p, (2 * n + 2 * ni) * sizeof(int32_t)));
        b->allocs.push_back(p);
        b->d_stage = (int32_t*)p;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *sm = b->d_stage, *sd = sm + n, *sl = sd + n, *sa = sl + ni;  // (same staging area, narrower use)
    CK(cudaMemcpyAsync(sm, op_machine, n * sizeof(uint8_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sd, op_duration, n * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sl, lower_bound, ni * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sa, max_allowed_time, ni * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    b->stage_packed = true;
    return JSL_OK;
}

// commit: pack the staged tables into the live ones (must be ordered after the last kernel that reads them)
int jsl_batch_commit_variants(jsl_batch_t* b, void* stream) {
    if (!b) return JSL_E_INVALID;
    if (!b->d_stage) { set_err("nothing staged"); return JSL_E_INVALID; }
    CK(cudaSetDevice(b->device));
    const InstDev& I = b->bd.inst;
    const size_t n = (size_t)I.n_inst * I.N, ni = (size_t)I.n_inst;
    int32_t *sm = b->d_stage, *sd = sm + n, *sl = sd + n, *sa = sl + ni;
    if (b->stage_packed)
        k_pack_variants<uint8_t, uint16_t><<<b->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(
            I.n_inst, I.N, I.J, I.M, b->d_bad, I.job_op_start, b->d_op_tool, (const uint8_t*)sm, (const uint16_t*)sd, sl, sa,
            (int32_t*)I.op_mt, (int32_t*)I.op_dur, (int32_t*)I.job_work, (int32_t*)I.lb_mat);
    else
        k_pack_variants<int32_t, int32_t><<<b->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(
            I.n_inst, I.N, I.J, I.M, b->d_bad, I.job_op_start, b->d_op_tool, sm, sd, sl, sa,
            (int32_t*)I.op_mt, (int32_t*)I.op_dur, (int32_t*)I.job_work, (int32_t*)I.lb_mat);
    b->launches++;
    CK(cudaGetLastError());
    return JSL_OK;
}

int jsl_batch_check(jsl_batch_t* b) {
    if (!b) return JSL_E_INVALID;
    CK(cudaSetDevice(b->device));
    CK(cudaDeviceSynchronize());
    int32_t bad = 0;
    CK(cudaMemcpy(&bad, b->d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
    if (bad) {
        set_err(std::to_string(bad) + " uploaded table entries were out of range (operation machine outside [0, n_machines) "
                "or negative duration) and were replaced by 0");
        return JSL_E_INVALID;
    }
    return JSL_OK;
}

int jsl_batch_load_variants(jsl_batch_t* b, const int32_t* op_machine, const int32_t* op_duration,
                            const int32_t* lower_bound, const int32_t* max_allowed_time, void* stream) {
    int rc = jsl_batch_stage_variants(b, op_machine, op_duration, lower_bound, max_allowed_time, stream);
    return rc != JSL_OK ? rc : jsl_batch_commit_variants(b, stream);
}

int jsl_batch_randomize(jsl_batch_t* b, uint64_t seed, int min_duration, int max_duration, const uint8_t* mask,
                        void* stream) {
    if (!b) return JSL_E_INVALID;
    if (max_duration <= min_duration || min_duration < 0) { set_err("empty range for randrange()"); return JSL_E_INVALID; }
    if (b->bd.tcur && !b->ops_static) { set_err("batches with stochastic operation durations cannot be randomised"); return JSL_E_UNSUPPORTED; }
    CK(cudaSetDevice(b->device));
    const InstDev& I = b->bd.inst;
    const int threads = 128;
    int64_t blocks = (I.n_inst + threads - 1) / threads;
    if (blocks > (int64_t)b->sm_count * 64) blocks = (int64_t)b->sm_count * 64;
    k_randomize<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(I.n_inst, I.N, I.J, I.job_op_start, seed, b->bd.env_base,
                                                                   min_duration, max_duration, mask, (int32_t*)I.op_mt,
                                                                   (int32_t*)I.op_dur, (int32_t*)I.job_work, (int32_t*)I.lb_mat);
    b->launches++;
    CK(cudaGetLastError());
    return JSL_OK;
}

int jsl_batch_get_variants(jsl_batch_t* b, int64_t first, int64_t count, int32_t* op_machine, int32_t* op_tool,
                           int32_t* op_duration, int32_t* lower_bound, int32_t* max_allowed_time) {
    if (!b) return JSL_E_INVALID;
    const InstDev& I = b->bd.inst;
    if (first < 0 || count < 0 || first + count > I.n_inst) return JSL_E_INVALID;
    CK(cudaSetDevice(b->device));
    CK(cudaDeviceSynchronize());
    const size_t n = (size_t)count * I.N;
    if (op_machine || op_tool) {
        std::vector<int32_t> mt(n);
        CK(cudaMemcpy(mt.data(), I.op_mt + first * I.N, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) {
            if (op_machine) op_machine[i] = mt[i] & 0xffff;
            if (op_tool) op_tool[i] = (mt[i] >> 16) & 0xffff;
        }
    }
    if (op_duration) CK(cudaMemcpy(op_duration, I.op_dur + first * I.N, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (lower_bound || max_allowed_time) {
        std::vector<int32_t> lm(2 * (size_t)count);
        CK(cudaMemcpy(lm.data(), I.lb_mat + 2 * first, lm.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < count; i++) {
            if (lower_bound) lower_bound[i] = lm[2 * i];
            if (max_allowed_time) max_allowed_time[i] = lm[2 * i + 1];
        }
    }
    return JSL_OK;
}

void jsl_batch_destroy(jsl_batch_t* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    for (void* p : b->allocs) cudaFree(p);
    delete b;
}

int jsl_batch_shape(const jsl_batch_t* b, jsl_shape* out) {
    if (!b || !out) return JSL_E_INVALID;
    const InstDev& I = b->bd.inst;
    out->n_jobs = I.J; out->n_machines = I.M; out->n_transports = I.T; out->n_sbuf = I.S; out->n_ops = I.N;
    out->n_tools = I.NT; out->obs_bytes = b->obs_bytes; out->state_bytes = b->bd.state_stride + b->bd.outs_stride;
    out->digest_cap = b->digest_cap;
    return JSL_OK;
}

int jsl_batch_state_layout(const jsl_batch_t* b, jsl_state_field* out, int cap) {
    if (!b || (!out && cap > 0)) return JSL_E_INVALID;
    return b->JW == 1 ? state_layout_jw1(b->bd.classic, out, cap) : b->JW == 2 ? state_layout_jw2(b->bd.classic, out, cap)
                                                                               : state_layout_jw4(b->bd.classic, out, cap);
}

int jsl_batch_set_sample_tape(jsl_batch_t* b, const int32_t* dev_tape, int64_t stride) {
    if (!b || (dev_tape && stride <= 0)) return JSL_E_INVALID;
    b->bd.tape = dev_tape;
    b->bd.tape_stride = dev_tape ? stride : 0;
    return JSL_OK;
}

int jsl_batch_set_start_seeds(jsl_batch_t* b, const int32_t* dev_seeds) {
    if (!b) return JSL_E_INVALID;
    b->bd.start_seeds = dev_seeds;
    return JSL_OK;
}

int jsl_batch_set_env_base(jsl_batch_t* b, int64_t env_base) {
    if (!b) return JSL_E_INVALID;
    b->bd.env_base = env_base;
    return JSL_OK;
}

#define JSL_DISPATCH(b, WHAT, lc, ...)                                                   \
    do {                                                                                 \
        cudaError_t e_ = (b)->JW == 1   ? launch_##WHAT##_jw1((b)->bd, lc, __VA_ARGS__)  \
                         : (b)->JW == 2 ? launch_##WHAT##_jw2((b)->bd, lc, __VA_ARGS__)  \
                                        : launch_##WHAT##_jw4((b)->bd, lc, __VA_ARGS__); \
        (b)->launches++;                                                                 \
        if (e_ != cudaSuccess) {                                                         \
            set_err(std::string(#WHAT " launch: ") + cudaGetErrorString(e_));            \
            return JSL_E_CUDA;                                                           \
        }                                                                                \
    } while (0)

int jsl_reset(jsl_batch_t* b, const uint8_t* reset_mask, jsl_step_out out, void* stream) {
    if (!b) return JSL_E_INVALID;
    CK(cudaSetDevice(b->device));
    LaunchCfg lc = {(int)((b->bd.B + WARPS_PER_CTA - 1) / WARPS_PER_CTA), WARPS_PER_CTA * b->smem_per_warp,
                    (cudaStream_t)stream};
    JSL_DISPATCH(b, reset, lc, reset_mask, out, b->obs_bytes);
    return JSL_OK;
}

int jsl_step(jsl_batch_t* b, const uint8_t* actions, jsl_step_out out, void* stream) {
    if (!b || !actions) return JSL_E_INVALID;
    if (out.flags & ~JSL_STEP_AUTO_RESET) { set_err("jsl_step: reserved bits set in jsl_step_out.flags"); return JSL_E_INVALID; }
    CK(cudaSetDevice(b->device));
    LaunchCfg lc = {(int)((b->bd.B + WARPS_PER_CTA - 1) / WARPS_PER_CTA), WARPS_PER_CTA * b->smem_per_warp,
                    (cudaStream_t)stream};
    JSL_DISPATCH(b, step, lc, actions, out, b->obs_bytes);
    return JSL_OK;
}

int jsl_rollout(jsl_batch_t* b, int policy, const uint8_t* action_tape, int64_t tape_stride, int max_steps, int flags,
                jsl_rollout_out out, void* stream) {
    if (!b) return JSL_E_INVALID;
    if (policy < JSL_POLICY_FIFO || policy > JSL_POLICY_TAPE) { set_err("unknown policy"); return JSL_E_INVALID; }
    if (policy == JSL_POLICY_TAPE && !action_tape) { set_err("tape policy needs an action tape"); return JSL_E_INVALID; }
    CK(cudaSetDevice(b->device));
    int smem = WARPS_PER_CTA * b->smem_per_warp;
    if (b->rollout_ctas_per_sm == 0) {
        b->rollout_ctas_per_sm = b->JW == 1 ? rollout_occupancy_jw1(b->bd, smem)
                                 : b->JW == 2 ? rollout_occupancy_jw2(b->bd, smem) : rollout_occupancy_jw4(b->bd, smem);
        if (const char* cap = std::getenv("JSL_ROLLOUT_CTAS_PER_SM")) {  // diagnostics: fewer resident CTAs (more L1 each)
            const int n = std::atoi(cap);
            if (n > 0 && n < b->rollout_ctas_per_sm) b->rollout_ctas_per_sm = n;
        }
    }
    int64_t need = (b->bd.B + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    int64_t persistent = (int64_t)b->sm_count * b->rollout_ctas_per_sm;  // one full wave of resident CTAs
    LaunchCfg lc = {(int)(need < persistent ? need : persistent), smem, (cudaStream_t)stream, b->rollout_ctas_per_sm};
    CK(cudaMemsetAsync(b->bd.work_counter, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    JSL_DISPATCH(b, rollout, lc, policy, action_tape, tape_stride, max_steps, flags, out);
    return JSL_OK;
}

int jsl_get_state(jsl_batch_t* b, int64_t idx, int32_t* host_out, int cap) {
    if (!b || !host_out || idx < 0 || idx >= b->bd.B) return JSL_E_INVALID;
    CK(cudaSetDevice(b->device));
    LaunchCfg lc = {1, b->smem_per_warp, (cudaStream_t)0};
    JSL_DISPATCH(b, digest, lc, idx, b->d_digest, b->digest_cap, b->d_digest_n);
    CK(cudaDeviceSynchronize());
    int32_t n = 0;
    CK(cudaMemcpy(&n, b->d_digest_n, sizeof(int32_t), cudaMemcpyDeviceToHost));
    int m = n < cap ? n : cap;
    if (m > b->digest_cap) m = b->digest_cap;
    CK(cudaMemcpy(host_out, b->d_digest, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return n;
}

int jsl_get_leg_log(jsl_batch_t* b, int64_t idx, int32_t* host_out, int cap) {
    if (!b || !host_out || idx < 0 || idx >= b->bd.B) return JSL_E_INVALID;
    if (!LEG_LOG_BUILD) { set_err("this build has no transport-task log: use libjsl_b200_log.so"); return JSL_E_UNSUPPORTED; }
    if (!b->bd.leg_log) { set_err("batch was created without record_ops"); return JSL_E_INVALID; }
    CK(cudaSetDevice(b->device));
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> raw(b->bd.leg_stride);
    CK(cudaMemcpy(raw.data(), b->bd.leg_log + idx * (int64_t)b->bd.leg_stride, raw.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    const int n = raw[0], hdr = 2 + 32 * b->JW;
    if (n > cap) return JSL_E_INVALID;
    std::memcpy(host_out, raw.data() + hdr, (size_t)n * 5 * sizeof(int32_t));
    return raw[1] ? JSL_E_NOMEM : n;
}

int jsl_get_op_log(jsl_batch_t* b, int64_t idx, int32_t* host_out) {
    if (!b || !host_out || idx < 0 || idx >= b->bd.B) return JSL_E_INVALID;
    if (!b->bd.op_log) { set_err("batch was created without record_ops"); return JSL_E_INVALID; }
    CK(cudaSetDevice(b->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(host_out, b->bd.op_log + idx * 2 * (int64_t)b->bd.inst.N, (size_t)b->bd.inst.N * 2 * sizeof(int32_t),
                  cudaMemcpyDeviceToHost));
    return JSL_OK;
}

int64_t jsl_launch_count(const jsl_batch_t* b) { return b ? b->launches : 0; }

}  // extern "C"
