/*
 * jobshoplab_b200 -- C ABI of the B200-native batched JobShopLab env-step path.
 *
 * Every entry point replaces a seam of the (pure Python) reference; the reference
 * file:line each one stands in for is cited next to it.  Plain pointers and sizes only,
 * no torch types.  Conventions: return 0 on success / negative JSL_E_* code, never throw
 * across the ABI; the caller owns every buffer; device work is asynchronous on the
 * stream passed in (a cudaStream_t cast to void*; NULL = legacy default stream); no host
 * sync inside jsl_step / jsl_rollout.  One stream at a time per jsl_batch_t.
 *
 * The same POD descriptor structs are consumed by the CPU oracle (oracle/jsl_oracle.c),
 * which is TEST INFRASTRUCTURE and is never linked into or called by this library.
 */
#ifndef JOBSHOPLAB_B200_H
#define JOBSHOPLAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JSL_ABI_VERSION 2

/* buffer types -- jobshoplab/types/instance_config_types.py:30-36 (BufferTypeConfig) */
#define JSL_BUF_FIFO 0
#define JSL_BUF_LIFO 1
#define JSL_BUF_DUMMY 2
#define JSL_BUF_FLEX 3
/* buffer roles -- instance_config_types.py:147-158 (BufferRoleConfig) */
#define JSL_ROLE_INPUT 0
#define JSL_ROLE_OUTPUT 1
#define JSL_ROLE_COMPONENT 2
#define JSL_ROLE_COMPENSATION 3
/* time behaviours -- jobshoplab/types/stochasticy_models.py:44-121, mapper.py:661-695 */
#define JSL_TIME_STATIC 0
#define JSL_TIME_POISSON 1
#define JSL_TIME_UNIFORM 2
#define JSL_TIME_GAUSSIAN 3
#define JSL_TIME_GAMMA 4
/* "no capacity limit" (reference: sys.maxsize) and "no entry in the mapping" */
#define JSL_CAP_INF 0x7fffffff
#define JSL_NO_TIME (-1)

/* machine states -- types/state_types.py:167-175 ; transport states :178-188 */
#define JSL_M_IDLE 0
#define JSL_M_SETUP 1
#define JSL_M_WORKING 2
#define JSL_M_OUTAGE 3
#define JSL_T_IDLE 0
#define JSL_T_WORKING 1
#define JSL_T_PICKUP 2
#define JSL_T_TRANSIT 3
#define JSL_T_OUTAGE 4
#define JSL_T_WAITINGPICKUP 5

/* Per-env status: the reference's exceptions / outcomes as data. */
#define JSL_ST_OK 0
#define JSL_ST_DONE 1                /* terminated: all jobs in an output buffer (core_utils.py:44) */
#define JSL_ST_TRUNCATED 2           /* truncation joker < 0 (middleware.py:433) */
#define JSL_ST_STEP_FAILED 3         /* StateMachineResult.success == False (state.py:202-210) -> truncated */
#define JSL_ST_BUFFER_FULL 4         /* BufferFullError (buffer_type_utils.py:90) */
#define JSL_ST_UNSUCCESSFUL 5        /* UnsuccessfulStateMachineResult (middleware.py:330,344) */
#define JSL_ST_NO_POSSIBLE 6         /* InvalidValue "No possible transitions" (actions.py:141) */
#define JSL_ST_ENV_DONE 7            /* EnvDone: step() after done (env.py:441) */
#define JSL_ST_ACTION_OUT_OF_SPACE 8 /* ActionOutOfActionSpace (actions.py:148) */
#define JSL_ST_OTHER_INVALID 9       /* any other raise (missing travel/setup time, not-implemented branch) */

/* error codes */
#define JSL_OK 0
#define JSL_E_INVALID (-1)
#define JSL_E_UNSUPPORTED (-2)
#define JSL_E_CUDA (-3)
#define JSL_E_NOMEM (-4)

/* policies for jsl_rollout -- jobshoplab/utils/dipatching_benchmark.py:17-67 */
#define JSL_POLICY_FIFO 0   /* always accept */
#define JSL_POLICY_SPT 1    /* dispatch_rules_uitls.py:57-79 */
#define JSL_POLICY_MWKR 2   /* dispatch_rules_uitls.py:82-111 ("srpt": least TOTAL job work) */
#define JSL_POLICY_RANDOM 3 /* Philox4x32-10(key=seed, ctr=(step,0,env_lo,env_hi)).x & 1 */
#define JSL_POLICY_TAPE 4   /* actions read from a u8 tape [B][tape_stride] */

/* observation kinds -- jobshoplab/env/factories/observations.py */
#define JSL_OBS_NONE 0
#define JSL_OBS_BINARY_ACTION 1 /* BinaryActionObservationFactory :488-551 (default) */
#define JSL_OBS_OPERATION_ARRAY 2 /* BinaryOperationArrayObservation :631-693 */
#define JSL_OBS_TASSEL 3          /* TasselJsspObservation :696-918 (implies record_ops) */
#define JSL_OBS_STATE 4           /* the env's raw state image (jsl_batch_state_layout names its fields): the input of
                                     user-defined observation / reward factories (observations.py:31-76,
                                     rewards.py:11-58 ABCs) written on the torch side */

/*
 * Lowered instance: what jobshoplab.compiler (compiler.py:84, mapper.py:1201,1387) produces,
 * as structure-of-arrays int32.  Index spaces:
 *   buffer idx : pre(m)=3m, post(m)=3m+1, machine buffer(m)=3m+2, agv(t)=3M+t, standalone s=3M+T+s
 *   place      : machine m -> m, standalone buffer s -> M+s          (travel matrix axis)
 *   time object: every DeterministicTimeConfig|StochasticTimeConfig of the instance; *_kind/_param
 *                arrays are NULL for a fully deterministic instance.
 */
/* Time behaviour of a family of time objects.  All three NULL = static (deterministic).
 * kind: JSL_TIME_*; base: base_time; param: offset (uniform) / std (gaussian) / scale (gamma). */
typedef struct jsl_time_spec {
    const int32_t* kind;
    const int32_t* base;
    const float* param;
} jsl_time_spec;

typedef struct jsl_instance_desc {
    int32_t abi_version;
    int32_t n_jobs, n_machines, n_transports, n_sbuf, n_tools, n_ops;
    const int32_t* job_op_start; /* [J+1] ragged op offsets */
    const int32_t* op_machine;   /* [N] */
    const int32_t* op_tool;      /* [N] */
    const int32_t* op_duration;  /* [N] duration.time at compile time */
    const int32_t* pre_type;     /* [M] JSL_BUF_* */
    const int32_t* pre_cap;      /* [M] */
    const int32_t* post_type;    /* [M] */
    const int32_t* post_cap;     /* [M] */
    const int32_t* setup_time;   /* [M*nt*nt] (old tool, new tool) .time; JSL_NO_TIME = missing */
    int32_t default_tool;        /* index of "tl-0" (mapper.py:396) */
    const int32_t* sbuf_type;    /* [S] */
    const int32_t* sbuf_cap;     /* [S] */
    const int32_t* sbuf_role;    /* [S] JSL_ROLE_* */
    const int32_t* buf_ref_id;   /* [3M+T+S] numeric part of the reference's "b-<k>" id per buffer idx */
    const int32_t* travel_time;  /* [(M+S)^2] row=from place, col=to place, .time; JSL_NO_TIME = missing */
    const int32_t* agv_init_place; /* [T] */
    const int32_t* job_init_buf; /* [J] buffer idx */
    int32_t start_time;
    /* outages (mapper.py:989-1008): machine outages per machine, transport outages shared by all AGVs */
    const int32_t* m_outage_start; /* [M+1] offsets into the m_outage_* arrays */
    const int32_t* m_outage_freq;  /* [n] frequency.time */
    const int32_t* m_outage_dur;   /* [n] duration.time */
    int32_t n_t_outages;
    const int32_t* t_outage_freq;  /* [n_t_outages] */
    const int32_t* t_outage_dur;   /* [n_t_outages] */
    /* stochastic behaviour of each family (index-aligned with the arrays above) */
    jsl_time_spec op_spec, setup_spec, travel_spec, m_ofreq_spec, m_odur_spec, t_ofreq_spec, t_odur_spec;
    /* sample tables of the stochastic time objects (jobshoplab_b200/stochastic.py): value of a time object
     * for numpy seed k = sample_table[sample_row[obj]][k]; obj in flat order ops | setup | travel | machine
     * outage freq | dur | transport outage freq | dur.  NULL for deterministic instances. */
    const int32_t* sample_table; /* [n_sample_rows][sample_depth] */
    const int32_t* sample_row;   /* [n time objects], -1 = static */
    int32_t n_sample_rows, sample_depth;
    /* constants computed once on the host (utils.py:310-352) */
    int32_t lower_bound;
    int32_t max_allowed_time;
} jsl_instance_desc;

/* Behavioural configuration: state_machine / middleware / reward_factory groups of
 * jobshoplab/types/config_types.py:135-145 that reach the step path. */
typedef struct jsl_env_cfg {
    int32_t allow_early_transport; /* state.py:325, default 1 */
    int32_t truncation_joker;      /* middleware.py:260, default 5 */
    int32_t truncation_active;     /* middleware.py:262, default 0 */
    int32_t obs_kind;              /* JSL_OBS_* */
    double sparse_bias;            /* rewards.py:90-153, defaults 1.0 / 0.001 / -1.0 */
    double dense_bias;
    double truncation_bias;
    int32_t record_ops;            /* keep per-op (start,end) log for render()/history (8(f)1) */
    int32_t reserved;
} jsl_env_cfg;

typedef struct jsl_instance jsl_instance_t;
typedef struct jsl_batch jsl_batch_t;

/* Sizes (bytes / elements) a caller needs to allocate step outputs. */
typedef struct jsl_shape {
    int32_t n_jobs, n_machines, n_transports, n_sbuf, n_ops, n_tools;
    int32_t obs_bytes;     /* packed observation record per env */
    int32_t state_bytes;   /* HBM bytes of one env's mutable state */
    int32_t digest_cap;    /* upper bound on jsl_get_state digest length (int32 elements) */
} jsl_shape;

/* Per-env scalars of one env.step packed into ONE 32-byte record (two 16-byte stores per env instead of eight
 * scattered ones): what env.step returns besides the observation (env.py:454) plus the candidate the agent sees
 * next.  Written when jsl_step_out.record is set -- the separate arrays are then not written. */
typedef struct jsl_step_record {
    double reward;       /* RewardFactory.make (rewards.py:145) */
    int32_t time;        /* state.time.time */
    int32_t makespan;    /* info["makespan"]: time if terminated else -1 (env.py:404) */
    int16_t cand[4];     /* next possible_transitions[0]: kind(0 m/1 t), comp, job, n_remaining */
    uint8_t terminated;  /* env.py:446 */
    uint8_t truncated;   /* env.py:447-450 */
    uint8_t status;      /* JSL_ST_* */
    uint8_t ended;       /* JSL_STEP_AUTO_RESET: 0, or the JSL_ST_* (>= JSL_ST_DONE) the episode ended with in this call --
                            then reward / terminated / truncated / makespan describe that finished step and time / status /
                            cand (and the observation) the first state of the fresh episode.  Always 0 without the flag. */
    int32_t reserved1;
} jsl_step_record;

/* One field of the state image an env keeps in HBM (= the JSL_OBS_STATE observation record). */
typedef struct jsl_state_field {
    char name[24];
    int32_t offset;     /* bytes from the start of the record */
    int32_t elem_bytes; /* 1, 2, 4 or 8 */
    int32_t count;      /* elements (arrays are padded to 32 per lane word; only the first n_jobs / n_machines /
                           n_transports entries are meaningful) */
    int32_t is_float;   /* 1: IEEE floating point (f64), 0: integer */
} jsl_state_field;

/* Device output pointers of one jsl_step; any pointer may be NULL (= not wanted). */
typedef struct jsl_step_out {
    uint8_t* obs;        /* [B][obs_bytes]   ObservationFactory.make (observations.py:515) */
    double* reward;      /* [B]              RewardFactory.make (rewards.py:145) */
    uint8_t* terminated; /* [B]              env.py:446 */
    uint8_t* truncated;  /* [B]              env.py:447-450 */
    uint8_t* status;     /* [B]              JSL_ST_* */
    int32_t* time;       /* [B]              state.time.time */
    int32_t* makespan;   /* [B]              info["makespan"] (time if terminated else -1), env.py:404 */
    int16_t* cand;       /* [B][4]           next possible_transitions[0]: kind(0 m/1 t), comp, job, n_remaining */
    jsl_step_record* record; /* [B]          all of the above scalars in one record per env (preferred: when set, the
                                             six scalar arrays and cand are ignored) */
    uint32_t flags;      /* JSL_STEP_* bits for jsl_step (other bits: reserved, must be 0; jsl_reset ignores them) */
    uint32_t reserved;
} jsl_step_out;

/* jsl_step_out.flags: reset every env whose episode ends in this step -- terminated, truncated, or dead with an escaped
 * exception (status >= JSL_ST_STEP_FAILED) -- inside the same launch, as a vector env with auto-reset does (what the
 * reference would do with env.reset() right after the final env.step, env.py:491).  No host round trip, no second
 * launch.  The operation tables in force are the ones the batch holds at that moment (no per-reset jsl_batch_randomize:
 * callers that redraw instances at every reset keep resetting explicitly). */
#define JSL_STEP_AUTO_RESET 1u

/* Device output pointers of one jsl_rollout; any may be NULL. */
typedef struct jsl_rollout_out {
    int32_t* makespan;   /* [B] final state.time.time */
    int32_t* n_steps;    /* [B] env.step calls made */
    double* ret;         /* [B] sum of rewards */
    uint8_t* status;     /* [B] JSL_ST_* at the end */
    int32_t* no_op_count;/* [B] info["no_op_count"] */
} jsl_rollout_out;

const char* jsl_last_error(void);
int jsl_abi_version(void);

/* Compiler.compile() result -> device-ready instance (compiler.py:84-106).  Host arrays are
 * copied; returns NULL (+ jsl_last_error) on failure. */
jsl_instance_t* jsl_instance_create(const jsl_instance_desc* d);
void jsl_instance_destroy(jsl_instance_t*);

/* B envs over n_inst same-shaped instances (env b uses instance b % n_inst), on CUDA device
 * `device`.  Replaces JobShopLabEnv.__init__ wiring (env.py:388-402) for a whole batch.  The instances share
 * everything but the operation tables; with stochastic time objects, several instances need static operation
 * durations (stochastic setup / travel / outage times are shared).
 * The batch is served by the most specialised kernel set the instance allows (classic / simple / queues /
 * features / general) -- all of them bit-exact with the reference; the environment variable
 * JSL_MAX_LEVEL=0|1, read here, forces the less specialised ones (diagnostics, cross-level parity tests). */
jsl_batch_t* jsl_batch_create(jsl_instance_t* const* insts, int n_inst, int64_t B, int device,
                              uint64_t seed, const jsl_env_cfg* cfg);
/* Synthetic same-shape batches: n_inst variants of `tmpl` that differ only in
 * the operation tables.  op_machine / op_duration: host int32 [n_inst][n_ops]; lower_bound /
 * max_allowed_time: host int32 [n_inst] (utils.py:310-352, computed by the caller). */
jsl_batch_t* jsl_batch_create_variants(const jsl_instance_t* tmpl, int64_t n_inst, const int32_t* op_machine,
                                       const int32_t* op_duration, const int32_t* lower_bound,
                                       const int32_t* max_allowed_time, int64_t B, int device, uint64_t seed,
                                       const jsl_env_cfg* cfg);
/* Replace the operation tables of all n_inst variants of a batch from HOST arrays (pinned memory makes
 * the copies asynchronous): the per-pass input of a synthetic-batch rollout.  Layout as in
 * jsl_batch_create_variants; packed on the device by a small kernel on `stream`. */
int jsl_batch_load_variants(jsl_batch_t*, const int32_t* op_machine, const int32_t* op_duration,
                            const int32_t* lower_bound, const int32_t* max_allowed_time, void* stream);
/* The two halves of jsl_batch_load_variants, for pipelining: `stage` only copies the host arrays into staging
 * buffers on `stream` (it may overlap a rollout that still reads the live tables, e.g. on a copy stream);
 * `commit` packs the staged tables into the live ones on `stream` (order it after the kernels that read them and
 * after the staging copies).  One staging area per batch: stage again only after the commit has run. */
int jsl_batch_stage_variants(jsl_batch_t*, const int32_t* op_machine, const int32_t* op_duration,
                             const int32_t* lower_bound, const int32_t* max_allowed_time, void* stream);
/* `stage` with narrow host tables -- uint8 machines, uint16 durations (3 instead of 8 bytes per operation over PCIe);
 * the caller guarantees that the values fit.  Committed by the same jsl_batch_commit_variants. */
int jsl_batch_stage_variants_packed(jsl_batch_t*, const uint8_t* op_machine, const uint16_t* op_duration,
                                    const int32_t* lower_bound, const int32_t* max_allowed_time, void* stream);
int jsl_batch_commit_variants(jsl_batch_t*, void* stream);
/* Uploaded tables are validated on the device while they are packed (the host arrays of the staged path are never
 * walked on the CPU): an operation machine outside [0, n_machines) or a negative duration is replaced by 0 and
 * counted.  jsl_batch_check synchronises and returns JSL_E_INVALID (+ jsl_last_error) if any such entry was seen
 * since the batch was created, JSL_OK otherwise.  (jsl_batch_create* validate their host arrays directly.)
 * A jsl_rollout that ends on a validation error or an escaping exception leaves the env with that status; its state
 * image is then unspecified (jsl_step restores the pre-step state instead, as the reference does, state.py:202-210). */
int jsl_batch_check(jsl_batch_t*);
/* compiler/manipulators.py:99-150 InstanceRandomizer.manipulate for every variant of the batch, on the device:
 * per job a uniformly random permutation of its operations (machine and tool travel together) and durations
 * uniform in [min_duration, max_duration) (random.randrange); per-job work, the lower bound and
 * max_allowed_time are recomputed (utils/utils.py:310-352).  Counter-based generator keyed by `seed` and the
 * global variant index (env_base + v), see jsl_randomize.cuh; the reference draws from Python's global RNGs,
 * so the distribution -- not sample values -- is the contract.  `mask` (dev u8 [n_inst], nullable) selects the
 * variants to redraw (the reference recompiles, hence re-randomises, on every env.reset).  Takes effect at the
 * next jsl_reset of the envs concerned.
 * JSL_E_INVALID for an empty range (the reference raises ValueError), JSL_E_UNSUPPORTED for batches whose operation
 * durations are stochastic (stochastic setup / travel / outage times are fine: the manipulator makes durations static). */
int jsl_batch_randomize(jsl_batch_t*, uint64_t seed, int min_duration, int max_duration, const uint8_t* mask,
                        void* stream);
/* Operation tables of variants [first, first+count) copied to HOST arrays ([count][n_ops] machine, tool,
 * duration; [count] lower bound, max_allowed_time; any pointer may be NULL).  Synchronises. */
int jsl_batch_get_variants(jsl_batch_t*, int64_t first, int64_t count, int32_t* op_machine, int32_t* op_tool,
                           int32_t* op_duration, int32_t* lower_bound, int32_t* max_allowed_time);
void jsl_batch_destroy(jsl_batch_t*);
int jsl_batch_shape(const jsl_batch_t*, jsl_shape* out);
/* Field table of the state image of this batch (layout of the JSL_OBS_STATE record and of jsl_shape.state_bytes):
 * writes up to `cap` entries, returns the number of fields. */
int jsl_batch_state_layout(const jsl_batch_t*, jsl_state_field* out, int cap);
/* Stochastic instances (stochasticy_models.py): by default every StochasticTimeConfig.update() draws from
 * the reference's own sample function (first variate of numpy default_rng(start seed + n), tabulated).  For bit-exact replay against
 * the reference, hand in the recorded update() results instead: dev int32 [B][stride], consumed in call
 * order (NULL switches back to Philox).  Takes effect at the next jsl_reset. */
int jsl_batch_set_sample_tape(jsl_batch_t*, const int32_t* dev_tape, int64_t stride);
/* numpy start seeds (stochasticy_models.py:12) of every time object, dev int32 [B][n time objects] in the
 * flat order above (-1 / ignored for static objects): makes a free-running episode reproduce the reference
 * run that had these seeds.  NULL (default): seeds are drawn uniformly from 0..30 per env from Philox. */
int jsl_batch_set_start_seeds(jsl_batch_t*, const int32_t* dev_seeds);
/* Global index of env 0 of this batch: the RANDOM policy draws from Philox stream (env_base + b), so a
 * job sharded over several GPUs reproduces the single-GPU action streams. */
int jsl_batch_set_env_base(jsl_batch_t*, int64_t env_base);

/* env.reset (env.py:456-545) + middleware.reset (middleware.py:264-290) for every env whose
 * reset_mask byte is non-zero (NULL = all).  Writes the first observation / candidate to `out`. */
int jsl_reset(jsl_batch_t*, const uint8_t* reset_mask /*dev [B], nullable*/, jsl_step_out out, void* stream);

/* env.step (env.py:427-454) for all B envs: actions dev u8[B] in {0,1}. */
int jsl_step(jsl_batch_t*, const uint8_t* actions, jsl_step_out out, void* stream);

/* `while not env.done: env.step(policy(env))` fused on device (dipatching_benchmark.py:131-137).
 * Runs at most max_steps further env.steps per env from the CURRENT state (call jsl_reset first).
 * action_tape: dev u8 [B][tape_stride] for JSL_POLICY_TAPE, else NULL. */
#define JSL_RO_RESET_FIRST 1   /* fuse env.reset into the launch (state is not read from HBM) */
#define JSL_RO_NO_WRITEBACK 2  /* do not store the final state image (results only) */
int jsl_rollout(jsl_batch_t*, int policy, const uint8_t* action_tape, int64_t tape_stride,
                int max_steps, int flags, jsl_rollout_out out, void* stream);

/* Canonical int32 digest of env `idx` (state + possible transitions; layout in DESIGN.md),
 * copied to host: parity tests, render(), history reconstruction.  Synchronises.
 * Returns the number of int32 written (<= cap) or a negative error. */
int jsl_get_state(jsl_batch_t*, int64_t idx, int32_t* host_out, int cap);

/* Per-op (start,end) log of env idx, host int32 [n_ops][2]; needs cfg.record_ops. */
int jsl_get_op_log(jsl_batch_t*, int64_t idx, int32_t* host_out);

/* Transport tasks of env idx in creation order, what env.render() draws for the AGVs
 * (rendering/gant_dashboard.py:299-359): host int32 [n][5] = agv (< 0: a task that ran entirely inside
 * env.reset, which keeps no history), job, start, end (-1: no time),
 * route = place | pickup buffer idx << 8 | drop place << 24.  Needs cfg.record_ops and the logging build of the
 * library (libjsl_b200_log.so: same sources with -DJSL_LEG_LOG, same ABI; the throughput build compiles the
 * logging out and answers JSL_E_UNSUPPORTED).  Returns n (<= cap) or a negative error (JSL_E_NOMEM: overflow). */
int jsl_get_leg_log(jsl_batch_t*, int64_t idx, int32_t* host_out, int cap);

/* How many kernels this batch has launched so far (bench.py "gpu_launches"). */
int64_t jsl_launch_count(const jsl_batch_t*);

#ifdef __cplusplus
}
#endif
#endif /* JOBSHOPLAB_B200_H */
