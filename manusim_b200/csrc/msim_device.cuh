// Shared declarations between the kernels (msim_kernel.cu, msim_reduce.cu) and the host C ABI (msim_api.cu).
#pragma once
#include <stdint.h>
#include "../../include/msim.h"

namespace msim {

constexpr int kWarp = 32;
constexpr int kStageCap = 32;      // staged records per replication (one 128-bit record per lane on flush)
constexpr int kStageFlushAt = 25;  // a single event emits at most 7 records
constexpr uint32_t kNone = 0xFFu;
constexpr int kTabMax = 12288;     // scenario tables travel in the kernel parameters (constant bank), not shared memory

// device-side distribution entry (scenario table in the kernel parameters)
struct DDist {
    double raw0;   // params[0] in float64 (constant batch = count * raw0, utils.py:83-84)
    float d0, d1;  // derived parameters rounded to float32 for the Philox samplers
    int32_t kind;
    float mt_d, mt_c;  // gamma/erlang: Marsaglia-Tsang d = k' - 1/3, c = 1/sqrt(9 d) with k' = k < 1 ? k + 1 : k
    int32_t pad;
};

// per-replication scalars that live in shared memory (only lane 0 touches them between warp syncs)
struct Scal {
    float tot_wip;          // _cached_total_wip           factory_sim.py:71,215,315
    float fgc_sum;          // sum(_cached_total_fg)        factory_sim.py:578,589
    uint32_t mach_wait;     // bit r: _production_system(r) blocked on resource_input[r].get()
    uint32_t trans_wait;    // bit r: _transportation(r) blocked on resource_output[r].get()
    uint32_t deliv_wait;    // bit p: _delivery_orders(p) blocked on outbound[p].get()
    uint32_t log_count;     // records emitted and flushed so far
    uint32_t draw[4];       // per-stream draw index (A,B,C,D)
    uint32_t tlen[4];       // replay: tape lengths
    const void* tbase[4];   // replay: tape base pointers of this replication
    // lane-0 loop state that is touched rarely (see Sim in msim_kernel.cu)
    double x_now_d;
    uint32_t x_uqh, x_uqt, x_logc, x_tseq, x_same_left, x_pending, x_rb_need, x_now_py, x_sel_req;
    int32_t x_status;
    uint32_t k0, k1;        // Philox replication key
    uint32_t rb_base[3];    // draw index held by slot 0 of the precomputed (x,u) buffer of streams A,B,C
    uint32_t n_timeouts;
    uint8_t sched_wait;     // scheduler blocked on inbound.get()
    uint8_t inb_head, inb_tail;   // inbound_demand_orders items (linked through do_next)
    uint8_t po_free, do_free;     // free lists
    uint8_t pad[3];
    // DBR (examples/toc_dbr/simulation.py)
    double tick_d;          // exact python-typed time of the pending rope tick
    float cb_level;         // constraint_buffer_level (:52, :285, :188)
    uint8_t fwd_wait;       // _process_demandOrders blocked on inbound.get()
    uint8_t drain_wait;     // _update_constraint_buffer blocked on resource_finished[CCR].get()
    uint8_t fin_n;          // items in resource_finished[CCR]
    uint8_t fin[5];
};

// byte offsets of every per-replication array inside the warp's shared-memory slab and of every table
// inside the scenario table blob (KParams::tblob).  Computed once per handle on the host.
struct Lay {
    int32_t R, P, S, NT, MP, MD, NQ, UQ, K;
    // per-replication slab.  Entity state is array-of-structures (strides below): r_utf (= 0), p_fg, po_rel, do_arr
    // and do_prod hold the start of the resource / product / production-order / demand-order (three floats; two link
    // bytes) struct arrays.  What the whole warp reads is structure-of-arrays, one slot per lane.
    int32_t stage, nq, uq, tm_time, tm_eid, tm_ev, rb;
    int32_t r_utf, p_fg, po_rel, do_arr, do_prod;
    int32_t r_qq, r_lastq;                                   // `queue` records only (log_queues)
    int32_t po_tt, po_eid;                                   // DBR: process_order delay timers
    int32_t do_tt, do_eid;                                   // onDue: delivery timers
    int32_t r_tpy, r_start_d, r_utf_d, r_flags, po_relpy;   // DBR: python-typed clock excursions (Appendix B.2)
    int32_t scal, slab_bytes;
    // scenario tables (offsets into KParams::tblob)
    int32_t t_route_start, t_step_res, t_step_dist, t_setup, t_tbf, t_ttr, t_freq, t_qty, t_due, t_shipbuf, t_ccrpt,
        t_bytes;
};

struct KParams {
    alignas(16) unsigned char tblob[kTabMax];   // scenario tables (routes, distributions): read with LDC, cost no shared memory
    Lay lay;
    unsigned long long* work_counter;
    float run_until, warmup;       // float32 views (heap comparisons happen at float32 precision, Appendix B.1)
    int32_t delivery_mode, log_queues;
    // DBR
    int32_t dbr_ccr, dbr_repair;
    double dbr_interval, dbr_cb_target, dbr_release_limit, dbr_ccr_limit, dbr_min_lot, dbr_ccr_setup;
    double warmup_d;               // exact (python-typed) warm-up instant
    int32_t warps_per_block, rng_streams;   // bit s: stream s (A,B,C) has at least one non-constant distribution
    msim_run_args_t a;             // device pointers
};

// Constant part of the layout.  msim_create lays every slab out in this order -- resource structs (room for kRmax), record
// staging, NORMAL fifo, URGENT fifo, calendar, (x,u) buffer, scalars, product structs (room for kPmax), production-order
// pool, then everything whose place depends on the scenario -- and the scenario tables as setup, route starts, step
// resources (room for kSmax), step distributions, then the rest.  When the
// scenario has the standard shape (R <= 24, P <= 16, <= 256 route steps, NORMAL fifo of 128, the policy's URGENT fifo) all
// of these offsets are the compile-time constants below, and the lean kernels (OPTS = 0) use them as immediates instead of
// loading `Lay` fields from the constant bank and adding them per access (a fifth of a hot handler's instructions).  Other
// shapes run the general kernels, which always read `Lay`.
struct CL {
    static constexpr int kRmax = 24, kPmax = 16, kSmax = 256;     // room reserved for resources / products / route steps
    static constexpr int NT = 32, NQ = 128;
    __host__ __device__ static constexpr int UQ(int policy) { return policy == MSIM_POLICY_DBR ? 64 : 16; }
    static constexpr int r_utf = 0;
    static constexpr int stage_ = kRmax * 28;                                            // 672 (16-byte aligned)
    static constexpr int nq_ = stage_ + kStageCap * 16;                                  // 1184
    static constexpr int uq_ = nq_ + NQ * 4;                                             // 1696
    __host__ __device__ static constexpr int stage(int) { return stage_; }
    __host__ __device__ static constexpr int nq(int) { return nq_; }
    __host__ __device__ static constexpr int uq(int) { return uq_; }
    __host__ __device__ static constexpr int tm_time(int policy) { return uq_ + UQ(policy) * 4; }
    __host__ __device__ static constexpr int tm_eid(int policy) { return tm_time(policy) + NT * 4; }
    __host__ __device__ static constexpr int tm_ev(int policy) { return tm_eid(policy) + NT * 4; }
    __host__ __device__ static constexpr int rb(int policy) { return (tm_ev(policy) + NT * 4 + 7) / 8 * 8; }
    __host__ __device__ static constexpr int scal(int policy) { return (rb(policy) + 2 * 16 * 8 + 7) / 8 * 8; }
    __host__ __device__ static constexpr int p_fg(int policy) { return (scal(policy) + (int)sizeof(Scal) + 3) / 4 * 4; }
    __host__ __device__ static constexpr int po_rel(int policy) { return p_fg(policy) + kPmax * 24; }
    // scenario tables
    __host__ __device__ static constexpr int t_setup(int) { return 0; }
    __host__ __device__ static constexpr int t_route_start(int) { return kRmax * (int)sizeof(DDist); }
    __host__ __device__ static constexpr int t_step_res(int) { return (kRmax * (int)sizeof(DDist) + (kPmax + 1) * 2 + 7) / 8 * 8; }
    __host__ __device__ static constexpr int t_step_dist(int) { return t_step_res(0) + kSmax; }
};

// struct strides of the entity arrays inside the slab, in bytes (field offsets: msim_kernel.cu)
constexpr int kAosResStride = 28, kAosPrdStride = 24, kAosDoStride = 12, kAosDoLinkStride = 2;
__host__ __device__ constexpr int aos_po_stride(int policy) { return policy == MSIM_POLICY_FIFO ? 12 : 16; }   // + prio (PRIO/EDD) or ccr (DBR)

// launchers implemented in msim_kernel.cu.  `shape` = launch bounds / register budget of the lean kernels (0: 256 x 4
// CTAs, 64 registers -- the default and the only shape of the general kernels; 1: 192 x 6, 56; 2: 256 x 5, 48); the
// launcher picks the lean kernel itself when the run uses neither queue records nor tape recording.
int launch_sim(const KParams& p, int policy, int rng_mode, int shape, bool const_layout, int grid, int block, size_t smem,
               void* stream);
int sim_kernel_attrs(int policy, int rng_mode, int shape, int delivery_mode, size_t smem, int* regs, int block, int* blocks_per_sm);
int sim_shape_max_warps(int shape);
int sim_shape_compiled(int shape, int delivery_mode);   // the shape a request for `shape` is served with in this build
// launchers implemented in msim_reduce.cu
int launch_reduce(const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status, int64_t n_reps, int K, int P,
                  int R, double span, double* out, void* stream);
int launch_reduce_variables(const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status, int64_t n_reps, int K,
                            int P, int R, double span, double* out, void* stream);
int launch_metric_values(const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status, int64_t n_reps, int K, int P,
                         int R, double span, double* out, void* stream);
int launch_test_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int64_t n);
int launch_test_variates(DDist d, uint64_t seed, int stream_id, int batch_count, double* out, int64_t n);
DDist make_ddist(const msim_dist_t& d);

} // namespace msim
