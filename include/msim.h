/*
 * msim.h -- C ABI of libmsim.so, the B200 (sm_100a) batched factory-replication engine.
 *
 * This is the drop-in boundary for ONE hot path of kleindaniell/manufacturing-simulation ("manusim"):
 *     ExperimentRunner._run_single_simulation            src/manusim/experiment.py:42-76
 *       -> FactorySimulation.reset_simulation(seed, ..)  src/manusim/factory_sim.py:804-808
 *       -> FactorySimulation.run_simulation()            src/manusim/factory_sim.py:810-814
 *            -> simpy.Environment.run(until)             (simpy==4.1.1, third party)
 *       -> products/resources/general_metrics()          src/manusim/factory_sim.py:734-789
 * executed for MANY independent replications at once.  The reference is pure Python and has no FFI; the
 * entry points below are what a ctypes binding placed behind ``FactorySimulation`` binds.
 *
 * Conventions:
 *   - plain C structs of pointers + sizes, no C++/torch types across the boundary;
 *   - every function returns 0 on success or a negative msim_error code; text via msim_last_error();
 *   - the caller owns all device/host buffers it passes; a handle owns only its compiled tables + scratch;
 *   - a handle is not thread-safe; one handle per (host thread, GPU); msim_run is asynchronous on the
 *     given CUDA stream, msim_run_host synchronises before returning;
 *   - per-replication model errors (SimPy ValueError cases, pool overflow) are reported in status[],
 *     not as a call failure.
 */
#ifndef MSIM_H
#define MSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSIM_VERSION 201 /* 0.2.1 */

/* distribution kinds -- DistributionGenerator.random_number, src/manusim/engine/utils.py:42-67 */
enum { MSIM_DIST_NONE = -1, MSIM_DIST_CONSTANT = 0, MSIM_DIST_UNIFORM = 1, MSIM_DIST_GAMMA = 2,
       MSIM_DIST_ERLANG = 3, MSIM_DIST_EXPO = 4, MSIM_DIST_NORMAL = 5 };

/* delivery modes -- FactorySimulation._delivery_orders, src/manusim/factory_sim.py:467-549 */
enum { MSIM_DELIVERY_AS_READY = 0, MSIM_DELIVERY_ON_DUE = 1, MSIM_DELIVERY_INSTANTLY = 2 };

/* dispatch / release policies (compile-time device functions):
 *   FIFO          FactorySimulation.order_selection + scheduler   factory_sim.py:463-465, 180-193
 *   MAX_PRIORITY  NewSimulation.order_selection                   examples/simple_scheduler/simulation.py:31-44
 *   DBR           DBRSimulation.scheduler/order_selection         examples/toc_dbr/simulation.py:200-328
 *   EDD           extension (not in the reference): earliest ProductionOrder.duedate first (orders.py:11,
 *                 factory_sim.py:188), FIFO among equal due dates; same machinery as MAX_PRIORITY */
enum { MSIM_POLICY_FIFO = 0, MSIM_POLICY_MAX_PRIORITY = 1, MSIM_POLICY_DBR = 2, MSIM_POLICY_EDD = 3 };

/* variate sources (BASELINE.json north_star "Random variates") */
enum { MSIM_RNG_PHILOX = 0, MSIM_RNG_REPLAY = 1 };

/* per-replication status[] values */
enum { MSIM_OK = 0,
       MSIM_ST_NEGATIVE_DELAY = 1,   /* SimPy Timeout(delay<0) -> ValueError */
       MSIM_ST_AMOUNT_NONPOSITIVE = 2, /* SimPy Container.put/get(amount<=0) -> ValueError */
       MSIM_ST_PO_POOL_OVERFLOW = 3, MSIM_ST_DO_POOL_OVERFLOW = 4, MSIM_ST_FIFO_OVERFLOW = 5,
       MSIM_ST_TAPE_EXHAUSTED = 6, MSIM_ST_WIP_GET_BLOCKED = 7, MSIM_ST_NOT_RUN = 8 };

/* call-level error codes */
enum { MSIM_E_INVALID = -1, MSIM_E_CUDA = -2, MSIM_E_UNSUPPORTED = -3, MSIM_E_NOMEM = -4 };

typedef struct msim_dist {
    int32_t kind;      /* MSIM_DIST_* */
    int32_t _pad;
    double raw0, raw1; /* params[0], params[1] as configured (mean, std) */
    double d0, d1;     /* derived in float64 exactly as utils.py:8-18: uniform (a,b); gamma/erlang (k,theta);
                          expo (mean,0); normal (mu,sigma); constant (value,0) */
} msim_dist_t;

/* Flat structure-of-arrays scenario tables (all HOST pointers; msim_create copies them).
 * Replaces the dict walks of stores.py:22-30 (routing) and factory_sim.py:144-152,249-260,404-410,425-430. */
typedef struct msim_tables {
    int32_t n_resources;             /* R <= 32 */
    int32_t n_products;              /* P <= 32 */
    int32_t n_steps;                 /* S = total route steps <= 4096 */
    const int32_t* route_start;      /* [P+1] steps of product p are [route_start[p], route_start[p+1]) */
    const int32_t* step_resource;    /* [S] */
    const msim_dist_t* step_time;    /* [S] processing_time */
    const msim_dist_t* res_setup;    /* [R] */
    const msim_dist_t* res_tbf;      /* [R] kind NONE => never fails (tbf = inf, factory_sim.py:258-260) */
    const msim_dist_t* res_ttr;      /* [R] */
    const msim_dist_t* prod_freq;    /* [P] demand.freq */
    const msim_dist_t* prod_qty;     /* [P] demand.quantity */
    const msim_dist_t* prod_due;     /* [P] demand.duedate */
    const double* prod_shipping_buffer; /* [P] (DBR), may be NULL */
    double run_until;                /* python numbers in the reference's config (int or float; the clock reaches them
                                        as python-typed instants) */
    double warmup;
    int32_t delivery_mode;           /* MSIM_DELIVERY_* */
    int32_t policy;                  /* MSIM_POLICY_* */
    int32_t log_queues;              /* factory_sim.py:53 */
    /* DBR parameters (examples/toc_dbr/simulation.py:29-32, 146-171, 201-204) */
    int32_t dbr_ccr;                 /* constraint resource index */
    int32_t dbr_repair;              /* oracle patch P3: feed resource_finished[CCR] */
    double dbr_interval, dbr_cb_target, dbr_release_limit, dbr_ccr_limit, dbr_min_lot, dbr_ccr_setup;
    const double* dbr_ccr_pt;        /* [P] sum of mean processing time on the CCR */
    /* capacities of the per-replication shared-memory pools (0 => defaults) */
    int32_t max_production_orders;   /* <= 250 */
    int32_t max_demand_orders;       /* <= 250 */
    int32_t fifo_capacity;           /* zero-delay event ring, power of two */
} msim_tables_t;

/* One streamed log record: 128 bits, written with one st.global.v4 per record.
 * (time,value) are the float32 pair Logger.log stores (src/manusim/engine/logs.py:64); ring identifies
 * (variable,key): product metric m, product p -> m*P+p ; resource metric m, resource r -> 11P+m*R+r ;
 * general metric g -> 11P+4R+g (enum orders of src/manusim/metrics.py:43-67).  seq = emission index. */
typedef struct msim_record {
    float time;
    float value;
    uint32_t ring;
    uint32_t seq;
} msim_record_t;

typedef struct msim_run_args {
    int64_t n_reps;                  /* replications in this call (this rank's shard) */
    int64_t rep_offset;              /* global index of replication 0 of this call (Philox key / sharding) */
    int32_t rng_mode;                /* MSIM_RNG_* */
    int32_t _pad;
    uint64_t philox_seed;            /* experiment-level key; replication key = f(philox_seed, global index) */
    const uint64_t* rep_seeds;       /* optional DEVICE [n_reps]: explicit per-replication Philox keys */
    /* replay tapes (DEVICE): values in call order per stream, rep i uses [off[i], off[i+1]) */
    const float* tape_A; const int64_t* tape_A_off;   /* rnd_product scalar draws  factory_sim.py:155-162 */
    const float* tape_B; const int64_t* tape_B_off;   /* rnd_process scalar draws  factory_sim.py:410 */
    const double* tape_C; const int64_t* tape_C_off;  /* rnd_process batch sums    factory_sim.py:436-438 */
    const float* tape_D; const int64_t* tape_D_off;   /* rnd_breakdown draws       factory_sim.py:258-260,275,288 */
    /* optional tape RECORDING (DEVICE, Philox mode): capacity per replication, same layout as above but
       with fixed stride rec_cap; counts in rec_count[4*n_reps] */
    float* rec_A; float* rec_B; double* rec_C; float* rec_D; int64_t rec_cap; uint32_t* rec_count;
    /* outputs (DEVICE, caller-allocated) */
    double* acc_sum;                 /* [n_reps][K] sum of post-warm-up record values per ring */
    uint32_t* acc_cnt;               /* [n_reps][K] number of post-warm-up records per ring */
    uint64_t* n_events;              /* [n_reps] SimPy-equivalent events (= Environment.step() count) */
    uint32_t* n_timeouts;            /* [n_reps] logical events (calendar pops), may be NULL */
    int32_t* status;                 /* [n_reps] MSIM_OK / MSIM_ST_* */
    msim_record_t* log;              /* [n_log_reps][log_cap] record rings, may be NULL */
    uint32_t* log_count;             /* [n_log_reps] total records emitted (ring wraps at log_cap) */
    int64_t log_cap;
    int64_t n_log_reps;              /* replications [0,n_log_reps) stream their records */
    void* stream;                    /* cudaStream_t (NULL = default stream) */
} msim_run_args_t;

/* Host-buffer twin used for the end-to-end path: all pointers are HOST memory (pinned or pageable) unless their
 * name ends in _dev; the call stages tapes/seeds H2D, runs, copies the outputs D2H and synchronises. */
typedef struct msim_host_args {
    int64_t n_reps;
    int64_t rep_offset;
    int32_t rng_mode;
    int32_t _pad;
    uint64_t philox_seed;
    const uint64_t* rep_seeds;
    const float* tape_A; const int64_t* tape_A_off;
    const float* tape_B; const int64_t* tape_B_off;
    const double* tape_C; const int64_t* tape_C_off;
    const float* tape_D; const int64_t* tape_D_off;
    double* acc_sum; uint32_t* acc_cnt; uint64_t* n_events; uint32_t* n_timeouts; int32_t* status;
    msim_record_t* log; uint32_t* log_count; int64_t log_cap; int64_t n_log_reps;
    double* kernel_ms;               /* optional out: device time of the simulation kernel */
    /* optional DEVICE buffers (caller-allocated): record rings that stay in HBM -- replications [0, n_log_reps_dev)
       stream their records exactly as in msim_run, nothing of them is copied back -- and the per-ring experiment
       statistics of this call, accumulated into stats_dev [K][3] like msim_reduce_metrics (what the ranks allreduce).
       `log` (host) and `log_dev` are mutually exclusive. */
    msim_record_t* log_dev; uint32_t* log_count_dev; int64_t log_cap_dev; int64_t n_log_reps_dev;
    double* stats_dev;
} msim_host_args_t;

typedef struct msim_info {
    int32_t n_rings;                 /* K = 11 P + 4 R + 3 */
    int32_t smem_per_replication;    /* bytes of shared memory one replication (one warp) owns */
    int32_t smem_tables;             /* bytes of scenario tables (carried in the kernel parameters / constant bank) */
    int32_t warps_per_block;
    int32_t blocks_per_sm;
    int32_t n_sms;
    int32_t regs_per_thread;
    int32_t max_production_orders, max_demand_orders, fifo_capacity;
    int32_t const_layout;            /* 1: the slab / tables have the standard constant layout, so runs without queue records
                                      * and tape recording launch the fast kernels with compile-time offsets; 0: general kernels */
} msim_info_t;

typedef struct msim_handle_s* msim_handle_t;

int msim_version(void);
const char* msim_last_error(void);

/* compile scenario tables for the current CUDA device */
int msim_create(const msim_tables_t* tables, msim_handle_t* out);
int msim_destroy(msim_handle_t h);
int msim_query(msim_handle_t h, msim_info_t* info);

/* run n_reps replications; device buffers; asynchronous on args->stream.  A handle may be launched on several streams at
 * once (every launch takes its own work counter; up to 64 launches of one handle in flight). */
int msim_run(msim_handle_t h, const msim_run_args_t* args);
/* same, host buffers, synchronous (H2D + kernel + D2H inside the call) */
int msim_run_host(msim_handle_t h, const msim_host_args_t* args);

/* Per-ring experiment statistics over replications (DEVICE buffers): for each ring the per-replication metric
 * value x (factory_sim.py:748-767: sum, sum/(run_until-warmup), or mean=sum/count; NaN runs dropped as in
 * metrics.py:863) is reduced to out[ring] = {sum x, sum x^2, n}.  out is [K][3] float64 and is ACCUMULATED
 * into (zero it first); the result is what the ranks allreduce over NCCL. */
int msim_reduce_metrics(msim_handle_t h, const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status,
                        int64_t n_reps, double* out, void* stream);

/* Per-VARIABLE statistics, the default of ExperimentMetrics.calculate_experiment_stats (metrics.py:809-816): per
 * replication the per-key values of a variable are averaged over its keys (NaN-skipping) before the reduction over
 * replications.  out is [18][3] float64 {sum, sum^2, n} in enum order (11 product, 4 resource, 3 general metrics),
 * ACCUMULATED into. */
int msim_reduce_variables(msim_handle_t h, const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status,
                          int64_t n_reps, double* out, void* stream);

/* Per-replication metric values (DEVICE buffers): out[rep][ring] is the number FactorySimulation.*_metrics() puts in
 * that run's metrics_*.csv (factory_sim.py:748-789: sum, sum / (run_until - warmup), or mean = sum / count, NaN when
 * no record was logged); all rings of a replication whose status != 0 are NaN.  out is [n_reps][K] float64.  This is
 * the `value` column of ExperimentMetrics' runs_metrics.csv (metrics.py:21,497-605) without per-run files. */
int msim_metric_values(msim_handle_t h, const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status,
                       int64_t n_reps, double* out, void* stream);

/* test hook: Philox4x32-10 blocks computed on the device; ctr [n][4], key [n][2], out [n][4] (HOST buffers) */
int msim_test_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int64_t n);
/* test hook: device variates; draws n values of `dist` with Philox key `seed` on stream id `stream_id`,
 * scalar (batch_count<0) or batch sums of batch_count samples; out is HOST double[n] */
int msim_test_variates(const msim_dist_t* dist, uint64_t seed, int32_t stream_id, int32_t batch_count,
                       double* out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* MSIM_H */
