// msim_api.cu -- host side of the C ABI declared in include/msim.h.
//
// msim_create compiles the scenario tables into (a) one table blob carried in the kernel parameters (constant
// bank) and (b) the byte layout of the per-replication shared-memory slab; it then sizes the launch as a
// persistent grid of  n_SMs x resident-CTAs-per-SM  blocks whose warps pull replication indices from a global
// counter (replications have very different lengths, so static striping would leave a long tail).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>
#include "msim_device.cuh"

using namespace msim;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(expr)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) return fail(MSIM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct msim_handle_s {
    KParams kp;                 // layout + scalar parameters (args filled per run)
    int policy;
    int n_sms;
    int warps_per_block[2];     // per rng mode (the Philox and replay kernels differ in register use)
    int blocks_per_sm[2];
    int regs[2];
    int shape;                  // launch bounds / register budget of the lean kernels (msim_kernel.cu)
    bool const_layout;          // the scenario has the standard shape: the lean kernels' constant layout applies
    size_t smem_bytes[2];
    unsigned long long* d_counter;      // kCounters work counters: every launch takes the next one, so launches of one
    std::atomic<unsigned> run_seq{0};   // handle on different streams never share a counter
    double span;                // run_until - warmup
    // scratch for msim_run_host
    void* d_scratch; size_t scratch_bytes;
    cudaEvent_t ev0, ev1;
};

constexpr unsigned kCounters = 64;

static int align_up(int x, int a) { return (x + a - 1) / a * a; }

static DDist to_ddist(const msim_dist_t& d) {
    DDist o;
    o.raw0 = d.raw0;
    o.kind = d.kind;
    o.pad = 0;
    o.d0 = (float)d.d0;
    o.d1 = (float)d.d1;
    o.mt_d = 0.f; o.mt_c = 0.f;
    if (d.kind == MSIM_DIST_GAMMA || d.kind == MSIM_DIST_ERLANG) {
        float k = o.d0;
        float kk = k < 1.0f ? k + 1.0f : k;
        o.mt_d = kk - (1.0f / 3.0f);
        o.mt_c = 1.0f / sqrtf(9.0f * o.mt_d);
    }
    return o;
}
DDist msim::make_ddist(const msim_dist_t& d) { return to_ddist(d); }

static const char* check_dist(const msim_dist_t& d, bool allow_none) {
    if (d.kind == MSIM_DIST_NONE) return allow_none ? nullptr : "missing distribution";
    if (d.kind < 0 || d.kind > MSIM_DIST_NORMAL) return "Unknowh distribution type";   // utils.py:65 (sic)
    return nullptr;
}

extern "C" int msim_version(void) { return MSIM_VERSION; }
extern "C" const char* msim_last_error(void) { return g_err.c_str(); }

extern "C" int msim_create(const msim_tables_t* t, msim_handle_t* out) {
    if (!t || !out) return fail(MSIM_E_INVALID, "null argument");
    const int R = t->n_resources, P = t->n_products, S = t->n_steps;
    if (R < 1 || R > 32 || P < 1 || P > 32) return fail(MSIM_E_UNSUPPORTED, "need 1 <= R <= 32 and 1 <= P <= 32");
    if (S < P || S > 4096) return fail(MSIM_E_INVALID, "bad n_steps");
    if (!(t->run_until > 0.0) || !(t->run_until > t->warmup) || t->warmup < 0.0)
        return fail(MSIM_E_INVALID, "need 0 <= warmup < run_until");
    if (t->delivery_mode < 0 || t->delivery_mode > 2) return fail(MSIM_E_INVALID, "bad delivery_mode");
    if (t->policy < MSIM_POLICY_FIFO || t->policy > MSIM_POLICY_EDD) return fail(MSIM_E_INVALID, "bad policy");
    if (t->policy == MSIM_POLICY_DBR) {
        if (!t->prod_shipping_buffer || !t->dbr_ccr_pt) return fail(MSIM_E_INVALID, "DBR needs shipping buffers and CCR times");
        if (t->dbr_ccr < 0 || t->dbr_ccr >= R || !(t->dbr_interval > 0.0)) return fail(MSIM_E_INVALID, "bad DBR parameters");
        for (int p = 0; p < P; ++p)
            if (!(t->prod_shipping_buffer[p] > 0.0))   // FG.put(0) raises in _update_finished_goods (toc_dbr :71-73)
                return fail(MSIM_E_INVALID, "DBR needs shipping_buffer > 0 for every product (SimPy Container.put(0) raises)");
    }
    for (int p = 0; p < P; ++p) {
        int n = t->route_start[p + 1] - t->route_start[p];
        if (n < 1 || n > 250) return fail(MSIM_E_INVALID, "every product needs 1..250 route steps");
        const char* e;
        if ((e = check_dist(t->prod_freq[p], false)) || (e = check_dist(t->prod_qty[p], false)) ||
            (e = check_dist(t->prod_due[p], false)))
            return fail(MSIM_E_INVALID, e);
    }
    if (t->route_start[0] != 0 || t->route_start[P] != S) return fail(MSIM_E_INVALID, "route_start must span [0,S]");
    for (int s = 0; s < S; ++s) {
        if (t->step_resource[s] < 0 || t->step_resource[s] >= R) return fail(MSIM_E_INVALID, "step_resource out of range");
        if (const char* e = check_dist(t->step_time[s], false)) return fail(MSIM_E_INVALID, e);
    }
    for (int r = 0; r < R; ++r) {
        const char* e;
        if ((e = check_dist(t->res_setup[r], false)) || (e = check_dist(t->res_tbf[r], true)))
            return fail(MSIM_E_INVALID, e);
        if (t->res_tbf[r].kind != MSIM_DIST_NONE && (e = check_dist(t->res_ttr[r], false)))
            return fail(MSIM_E_INVALID, "ttr is required when tbf is configured");
    }

    msim_handle_s* h = new msim_handle_s();
    memset(&h->kp, 0, sizeof(h->kp));
    h->policy = t->policy;
    h->d_scratch = nullptr; h->scratch_bytes = 0;
    Lay& L = h->kp.lay;
    L.R = R; L.P = P; L.S = S;
    L.NT = align_up(R + P + 2, 32);
    L.MP = t->max_production_orders > 0 ? t->max_production_orders : 96;
    L.MD = t->max_demand_orders > 0 ? t->max_demand_orders : 96;
    L.NQ = t->fifo_capacity > 0 ? t->fifo_capacity : 128;   // every live process owns at most one pending zero-delay event
    L.UQ = t->policy == MSIM_POLICY_DBR ? 64 : 16;          // a rope tick spawns up to P process_order Initialize events
    if (L.MP > 250 || L.MD > 250 || (L.NQ & (L.NQ - 1))) { delete h; return fail(MSIM_E_INVALID, "bad pool capacities"); }
    L.K = 11 * P + 4 * R + 3;
    const bool dbr = t->policy == MSIM_POLICY_DBR;
    const bool on_due = t->delivery_mode == MSIM_DELIVERY_ON_DUE;

    // ---- per-replication slab layout
    int off = 0;
    auto take = [&](int bytes, int align) { off = align_up(off, align); int o = off; off += bytes; return o; };
    // Constant part first, in the order of msim_device.cuh `CL` (resource structs with room for 32, record staging, NORMAL
    // fifo, URGENT fifo, calendar, (x,u) buffer, scalars, production-order pool): for the standard shape (calendar of 32
    // slots, NORMAL fifo of 128) these offsets equal the compile-time constants the lean kernels use as immediates; every
    // kernel can still read them from `Lay`.  Then everything whose place depends on the scenario.
    L.r_utf = take((R > CL::kRmax ? R : CL::kRmax) * kAosResStride, 4);
    L.stage = take(kStageCap * 16, 16);
    L.nq = take(L.NQ * 4, 4);
    L.uq = take(L.UQ * 4, 4);
    L.tm_time = take(L.NT * 4, 4);
    L.tm_eid = take(L.NT * 4, 4);
    L.tm_ev = take(L.NT * 4, 4);
    L.rb = take(2 * 16 * 8, 8);                    // precomputed (x,u) pairs of streams B,C (philox.cuh kXuBuf)
    L.scal = take((int)sizeof(Scal), 8);
    L.p_fg = take((P > CL::kPmax ? P : CL::kPmax) * kAosPrdStride, 4);
    L.po_rel = take(L.MP * aos_po_stride(t->policy), 4);
    L.do_arr = take(L.MD * kAosDoStride, 4);
    L.do_prod = take(L.MD * kAosDoLinkStride, 2);
    L.r_tpy = dbr ? take(R * 8, 8) : 0; L.r_start_d = dbr ? take(R * 8, 8) : 0; L.r_utf_d = dbr ? take(R * 8, 8) : 0;
    L.r_qq = t->log_queues ? take(R * 4, 4) : 0; L.r_lastq = t->log_queues ? take(R * 4, 4) : 0;   // `queue` records only
    L.po_tt = dbr ? take(L.MP * 4, 4) : 0; L.po_eid = dbr ? take(L.MP * 4, 4) : 0;
    L.do_tt = on_due ? take(L.MD * 4, 4) : 0; L.do_eid = on_due ? take(L.MD * 4, 4) : 0;   // delivery timers: onDue only
    L.r_flags = dbr ? take(R, 1) : 0; L.po_relpy = dbr ? take(L.MP, 1) : 0;
    L.slab_bytes = align_up(off, 16);

    // ---- scenario table blob (carried in the kernel parameters)
    std::vector<unsigned char> blob;
    auto put = [&](const void* src, size_t bytes, size_t align) {
        size_t o = (blob.size() + align - 1) / align * align;
        blob.resize(o + bytes);
        memcpy(blob.data() + o, src, bytes);
        return (int)o;
    };
    {
        std::vector<uint16_t> rs(P + 1);
        for (int p = 0; p <= P; ++p) rs[p] = (uint16_t)t->route_start[p];
        std::vector<uint8_t> sr(S);
        for (int s = 0; s < S; ++s) sr[s] = (uint8_t)t->step_resource[s];
        std::vector<DDist> sd(S), su(R), tb(R), tr(R), fq(P), qt(P), du(P);
        for (int s = 0; s < S; ++s) sd[s] = to_ddist(t->step_time[s]);
        for (int r = 0; r < R; ++r) {
            su[r] = to_ddist(t->res_setup[r]);
            tb[r] = to_ddist(t->res_tbf[r]);
            msim_dist_t none; memset(&none, 0, sizeof(none)); none.kind = MSIM_DIST_CONSTANT;
            tr[r] = to_ddist(t->res_tbf[r].kind == MSIM_DIST_NONE ? none : t->res_ttr[r]);
        }
        for (int p = 0; p < P; ++p) {
            fq[p] = to_ddist(t->prod_freq[p]); qt[p] = to_ddist(t->prod_qty[p]); du[p] = to_ddist(t->prod_due[p]);
        }
        std::vector<double> sbuf(P, 0.0), cpt(P, 0.0);
        for (int p = 0; p < P; ++p) {
            if (t->prod_shipping_buffer) sbuf[p] = t->prod_shipping_buffer[p];
            if (t->dbr_ccr_pt) cpt[p] = t->dbr_ccr_pt[p];
        }
        // constant part first (msim_device.cuh `CL`): setup distributions with room for 32 resources, route starts with
        // room for 32 products, then the step distributions; the rest follows
        if (R < CL::kRmax) su.resize(CL::kRmax, to_ddist(msim_dist_t{MSIM_DIST_CONSTANT, 0, 0.0, 0.0, 0.0, 0.0}));
        if (P < CL::kPmax) rs.resize(CL::kPmax + 1, rs.back());
        if (S < CL::kSmax) sr.resize(CL::kSmax, 0);
        L.t_setup = put(su.data(), su.size() * sizeof(DDist), 8);
        L.t_route_start = put(rs.data(), rs.size() * 2, 2);
        L.t_step_res = put(sr.data(), sr.size(), 8);
        L.t_step_dist = put(sd.data(), sd.size() * sizeof(DDist), 8);
        L.t_tbf = put(tb.data(), tb.size() * sizeof(DDist), 8);
        L.t_ttr = put(tr.data(), tr.size() * sizeof(DDist), 8);
        L.t_freq = put(fq.data(), fq.size() * sizeof(DDist), 8);
        L.t_qty = put(qt.data(), qt.size() * sizeof(DDist), 8);
        L.t_due = put(du.data(), du.size() * sizeof(DDist), 8);
        if (L.t_qty != L.t_freq + P * (int)sizeof(DDist) || L.t_due != L.t_qty + P * (int)sizeof(DDist)) {
            delete h;
            return fail(MSIM_E_INVALID, "internal: demand tables must be contiguous");
        }
        L.t_shipbuf = put(sbuf.data(), sbuf.size() * 8, 8);
        L.t_ccrpt = put(cpt.data(), cpt.size() * 8, 8);
        blob.resize((blob.size() + 15) / 16 * 16);
        L.t_bytes = (int)blob.size();
        if (blob.size() > (size_t)kTabMax) {
            delete h;
            return fail(MSIM_E_UNSUPPORTED, "scenario tables exceed the 12 KB kernel-parameter budget (too many route steps)");
        }
        memcpy(h->kp.tblob, blob.data(), blob.size());
    }

    // do the offsets equal the constants the lean kernels were compiled with?  (the standard shape; anything else runs
    // the general kernels, which read `Lay`)
    {
        const int pol = t->policy;
        h->const_layout = L.NT == CL::NT && L.NQ == CL::NQ && L.UQ == CL::UQ(pol) && L.stage == CL::stage(pol) &&
                          L.nq == CL::nq(pol) && L.uq == CL::uq(pol) && L.tm_time == CL::tm_time(pol) &&
                          L.tm_eid == CL::tm_eid(pol) && L.tm_ev == CL::tm_ev(pol) && L.rb == CL::rb(pol) &&
                          L.scal == CL::scal(pol) && L.p_fg == CL::p_fg(pol) && L.po_rel == CL::po_rel(pol) &&
                          L.t_setup == CL::t_setup(pol) && L.t_route_start == CL::t_route_start(pol) &&
                          L.t_step_res == CL::t_step_res(pol) && L.t_step_dist == CL::t_step_dist(pol);
    }

    KParams& kp = h->kp;
    kp.run_until = (float)t->run_until;
    kp.warmup = (float)t->warmup;
    kp.warmup_d = t->warmup;
    {
        int m = 0;
        for (int p = 0; p < P; ++p)
            if (t->prod_freq[p].kind > 0 || t->prod_qty[p].kind > 0 || t->prod_due[p].kind > 0) m |= 1;
        for (int r = 0; r < R; ++r) if (t->res_setup[r].kind > 0) m |= 2;
        for (int s_ = 0; s_ < S; ++s_) if (t->step_time[s_].kind > 0) m |= 4;
        kp.rng_streams = m;
    }
    kp.delivery_mode = t->delivery_mode;
    kp.log_queues = t->log_queues ? 1 : 0;
    kp.dbr_ccr = t->dbr_ccr; kp.dbr_repair = t->dbr_repair;
    kp.dbr_interval = t->dbr_interval; kp.dbr_cb_target = t->dbr_cb_target; kp.dbr_release_limit = t->dbr_release_limit;
    kp.dbr_ccr_limit = t->dbr_ccr_limit; kp.dbr_min_lot = t->dbr_min_lot; kp.dbr_ccr_setup = t->dbr_ccr_setup;
    h->span = t->run_until - t->warmup;

    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        delete h;
        return fail(MSIM_E_CUDA, "no CUDA device: libmsim has no CPU fallback");
    }
    h->n_sms = prop.multiProcessorCount;
    // warps per block: maximise resident replications (= warps) per SM.  Shared memory, the per-CTA copy of the
    // tables and the register file all bound it, so ask the occupancy calculator for every CTA width.
    const size_t smem_blk_max = prop.sharedMemPerBlockOptin;           // 227 KB
    // Launch shape (msim_kernel.cu) = register budget against resident warps.  Measured on a B200
    // (profiles/r02_launch_shape_sweep_*.jsonl): 56 registers x 36 warps per SM is the best trade for every workload
    // that shared memory lets reach it; a handle that shared memory holds at <= 32 warps takes the 64-register build.
    // Diagnostics for sweeps (tools/occupancy_sweep.sh): MSIM_SHAPE = 0 / 1 / 2 forces a shape, MSIM_MAX_CTAS_PER_SM
    // caps the resident CTAs per SM, MSIM_WARPS_PER_CTA fixes the CTA width.
    int cta_cap = 0, w_fixed = 0, shape_forced = -1;
    if (const char* e = getenv("MSIM_MAX_CTAS_PER_SM")) cta_cap = atoi(e);
    if (const char* e = getenv("MSIM_WARPS_PER_CTA")) w_fixed = atoi(e);
    if (const char* e = getenv("MSIM_SHAPE")) shape_forced = atoi(e);
    if (shape_forced > 2) { delete h; return fail(MSIM_E_INVALID, "MSIM_SHAPE must be 0, 1 or 2"); }
    struct Fit { int w, blocks, total, regs; };
    int rc_fit = 0;
    auto fit = [&](int shape, int rng) {
        Fit f{0, 0, 0, 0};
        for (int w = 1; w <= sim_shape_max_warps(shape); ++w) {
            if (w_fixed > 0 && w != w_fixed) continue;
            size_t need = (size_t)w * L.slab_bytes;
            if (need > smem_blk_max) break;
            int nblk = 0, regs = 0;
            int rc = sim_kernel_attrs(h->policy, rng, shape, t->delivery_mode, need, &regs, w * 32, &nblk);
            if (rc != 0) { rc_fit = rc; return f; }
            f.regs = regs;
            if (cta_cap > 0 && nblk > cta_cap) nblk = cta_cap;
            if (nblk * w > f.total || (nblk * w == f.total && w > f.w)) { f.total = nblk * w; f.w = w; f.blocks = nblk; }
        }
        return f;
    };
    int shape = shape_forced >= 0 ? shape_forced : 1;
    if (shape_forced < 0 && sim_shape_compiled(0, t->delivery_mode) == 0) {
        Fit f1 = fit(1, MSIM_RNG_PHILOX), f0 = fit(0, MSIM_RNG_PHILOX);
        if (rc_fit == 0 && f0.total >= f1.total) shape = 0;      // no residency to gain from the smaller register budget
    }
    shape = sim_shape_compiled(shape, t->delivery_mode);
    h->shape = shape;
    for (int rng = 0; rng < 2 && rc_fit == 0; ++rng) {
        Fit f = fit(shape, rng);
        if (rc_fit != 0) break;
        if (f.total == 0) { delete h; return fail(MSIM_E_UNSUPPORTED, "scenario does not fit shared memory"); }
        h->regs[rng] = f.regs;
        h->warps_per_block[rng] = f.w;
        h->blocks_per_sm[rng] = f.blocks;
        h->smem_bytes[rng] = (size_t)f.w * L.slab_bytes;
    }
    if (rc_fit != 0) {
        delete h;
        return fail(rc_fit, rc_fit == MSIM_E_UNSUPPORTED ? "policy not available in this build" : "kernel attribute query failed");
    }
    if (cudaMalloc(&h->d_counter, 8 * kCounters) != cudaSuccess) {
        delete h;
        return fail(MSIM_E_NOMEM, "cudaMalloc failed");
    }
    kp.work_counter = h->d_counter;
    h->ev0 = nullptr; h->ev1 = nullptr;
    if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
        msim_destroy(h);
        return fail(MSIM_E_CUDA, "cudaEventCreate failed");
    }
    *out = h;
    return 0;
}

extern "C" int msim_destroy(msim_handle_t h) {
    if (!h) return 0;
    cudaFree(h->d_counter);
    if (h->d_scratch) cudaFree(h->d_scratch);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return 0;
}

extern "C" int msim_query(msim_handle_t h, msim_info_t* info) {
    if (!h || !info) return fail(MSIM_E_INVALID, "null argument");
    const Lay& L = h->kp.lay;
    info->n_rings = L.K;
    info->smem_per_replication = L.slab_bytes;
    info->smem_tables = L.t_bytes;
    info->warps_per_block = h->warps_per_block[0];
    info->blocks_per_sm = h->blocks_per_sm[0];
    info->n_sms = h->n_sms;
    info->regs_per_thread = h->regs[0];
    info->max_production_orders = L.MP;
    info->max_demand_orders = L.MD;
    info->fifo_capacity = L.NQ;
    info->const_layout = h->const_layout;
    return 0;
}

extern "C" int msim_run(msim_handle_t h, const msim_run_args_t* a) {
    if (!h || !a) return fail(MSIM_E_INVALID, "null argument");
    if (a->n_reps <= 0) return 0;
    if (!a->acc_sum || !a->acc_cnt || !a->n_events || !a->status) return fail(MSIM_E_INVALID, "missing output buffer");
    if (a->rng_mode != MSIM_RNG_PHILOX && a->rng_mode != MSIM_RNG_REPLAY) return fail(MSIM_E_INVALID, "bad rng_mode");
    if (a->rng_mode == MSIM_RNG_REPLAY &&
        (!a->tape_A || !a->tape_B || !a->tape_C || !a->tape_D || !a->tape_A_off || !a->tape_B_off || !a->tape_C_off ||
         !a->tape_D_off))
        return fail(MSIM_E_INVALID, "replay mode needs the four tapes and their offsets");
    if (a->log && (a->log_cap <= 0 || !a->log_count)) return fail(MSIM_E_INVALID, "log ring needs log_cap > 0 and log_count");
    KParams kp = h->kp;
    kp.a = *a;
    if (!a->log) { kp.a.n_log_reps = 0; }
    cudaStream_t st = (cudaStream_t)a->stream;
    kp.work_counter = h->d_counter + h->run_seq.fetch_add(1u) % kCounters;
    CU(cudaMemsetAsync(kp.work_counter, 0, 8, st));
    const int rng = a->rng_mode;
    const int wpb = h->warps_per_block[rng];
    kp.warps_per_block = wpb;
    long long blocks_needed = (a->n_reps + wpb - 1) / wpb;
    long long grid = (long long)h->n_sms * h->blocks_per_sm[rng];
    if (grid > blocks_needed) grid = blocks_needed;
    int rc = launch_sim(kp, h->policy, rng, h->shape, h->const_layout, (int)grid, wpb * 32, h->smem_bytes[rng], st);
    if (rc != 0) return fail(rc, "simulation kernel launch failed (grid/block/shared-memory configuration)");
    return 0;
}

extern "C" int msim_reduce_metrics(msim_handle_t h, const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status,
                                   int64_t n_reps, double* out, void* stream) {
    if (!h || !acc_sum || !acc_cnt || !status || !out) return fail(MSIM_E_INVALID, "null argument");
    const Lay& L = h->kp.lay;
    int rc = launch_reduce(acc_sum, acc_cnt, status, n_reps, L.K, L.P, L.R, h->span, out, stream);
    if (rc != 0) return fail(rc, "reduce launch failed");
    return 0;
}

extern "C" int msim_reduce_variables(msim_handle_t h, const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status,
                                     int64_t n_reps, double* out, void* stream) {
    if (!h || !acc_sum || !acc_cnt || !status || !out) return fail(MSIM_E_INVALID, "null argument");
    const Lay& L = h->kp.lay;
    int rc = launch_reduce_variables(acc_sum, acc_cnt, status, n_reps, L.K, L.P, L.R, h->span, out, stream);
    if (rc != 0) return fail(rc, "reduce launch failed");
    return 0;
}

extern "C" int msim_metric_values(msim_handle_t h, const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status,
                                  int64_t n_reps, double* out, void* stream) {
    if (!h || !acc_sum || !acc_cnt || !status || !out) return fail(MSIM_E_INVALID, "null argument");
    if (n_reps <= 0) return 0;
    const Lay& L = h->kp.lay;
    int rc = launch_metric_values(acc_sum, acc_cnt, status, n_reps, L.K, L.P, L.R, h->span, out, stream);
    if (rc != 0) return fail(rc, "metric-values launch failed");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer path: stage inputs, run, copy results back (the e2e leg of bench.py and the ctypes default)
extern "C" int msim_run_host(msim_handle_t h, const msim_host_args_t* a) {
    if (!h || !a) return fail(MSIM_E_INVALID, "null argument");
    if (a->n_reps <= 0) return 0;
    const Lay& L = h->kp.lay;
    const int64_t n = a->n_reps;
    const bool replay = a->rng_mode == MSIM_RNG_REPLAY;
    int64_t la = 0, lb = 0, lc = 0, ld = 0;
    if (replay) {
        if (!a->tape_A_off || !a->tape_B_off || !a->tape_C_off || !a->tape_D_off)
            return fail(MSIM_E_INVALID, "replay mode needs tape offsets");
        la = a->tape_A_off[n]; lb = a->tape_B_off[n]; lc = a->tape_C_off[n]; ld = a->tape_D_off[n];
    }
    const bool logs = a->log && a->n_log_reps > 0 && a->log_cap > 0;
    const bool logs_dev = a->log_dev && a->n_log_reps_dev > 0 && a->log_cap_dev > 0;
    if (logs && logs_dev) return fail(MSIM_E_INVALID, "host log rings and device log rings are mutually exclusive");
    if (logs_dev && !a->log_count_dev) return fail(MSIM_E_INVALID, "device log rings need log_count_dev");
    const int64_t nlog = logs ? (a->n_log_reps < n ? a->n_log_reps : n) : 0;
    // carve one scratch allocation
    size_t off = 0;
    auto carve = [&](size_t bytes) { off = (off + 255) / 256 * 256; size_t o = off; off += bytes; return o; };
    size_t o_sum = carve((size_t)n * L.K * 8), o_cnt = carve((size_t)n * L.K * 4), o_ev = carve((size_t)n * 8),
           o_to = carve((size_t)n * 4), o_st = carve((size_t)n * 4), o_seed = carve((size_t)n * 8),
           o_ta = carve((size_t)la * 4), o_tb = carve((size_t)lb * 4), o_tc = carve((size_t)lc * 8), o_td = carve((size_t)ld * 4),
           o_oa = carve((size_t)(n + 1) * 8), o_ob = carve((size_t)(n + 1) * 8), o_oc = carve((size_t)(n + 1) * 8),
           o_od = carve((size_t)(n + 1) * 8), o_log = carve((size_t)nlog * (size_t)a->log_cap * 16),
           o_lc = carve((size_t)(nlog ? nlog : 1) * 4);
    size_t total = carve(0);
    if (total > h->scratch_bytes) {
        if (h->d_scratch) cudaFree(h->d_scratch);
        h->d_scratch = nullptr; h->scratch_bytes = 0;
        if (cudaMalloc(&h->d_scratch, total) != cudaSuccess) return fail(MSIM_E_NOMEM, "scratch cudaMalloc failed");
        h->scratch_bytes = total;
    }
    unsigned char* base = (unsigned char*)h->d_scratch;
    cudaStream_t st = 0;
    msim_run_args_t r;
    memset(&r, 0, sizeof(r));
    r.n_reps = n; r.rep_offset = a->rep_offset; r.rng_mode = a->rng_mode; r.philox_seed = a->philox_seed;
    if (a->rep_seeds) {
        CU(cudaMemcpyAsync(base + o_seed, a->rep_seeds, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        r.rep_seeds = (const uint64_t*)(base + o_seed);
    }
    if (replay) {
        CU(cudaMemcpyAsync(base + o_ta, a->tape_A, (size_t)la * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(base + o_tb, a->tape_B, (size_t)lb * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(base + o_tc, a->tape_C, (size_t)lc * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(base + o_td, a->tape_D, (size_t)ld * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(base + o_oa, a->tape_A_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(base + o_ob, a->tape_B_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(base + o_oc, a->tape_C_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(base + o_od, a->tape_D_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        r.tape_A = (const float*)(base + o_ta); r.tape_B = (const float*)(base + o_tb);
        r.tape_C = (const double*)(base + o_tc); r.tape_D = (const float*)(base + o_td);
        r.tape_A_off = (const int64_t*)(base + o_oa); r.tape_B_off = (const int64_t*)(base + o_ob);
        r.tape_C_off = (const int64_t*)(base + o_oc); r.tape_D_off = (const int64_t*)(base + o_od);
    }
    r.acc_sum = (double*)(base + o_sum); r.acc_cnt = (uint32_t*)(base + o_cnt);
    r.n_events = (uint64_t*)(base + o_ev); r.n_timeouts = (uint32_t*)(base + o_to); r.status = (int32_t*)(base + o_st);
    if (logs) {
        r.log = (msim_record_t*)(base + o_log); r.log_count = (uint32_t*)(base + o_lc);
        r.log_cap = a->log_cap; r.n_log_reps = nlog;
        CU(cudaMemsetAsync(base + o_lc, 0, (size_t)nlog * 4, st));
    } else if (logs_dev) {                       // records stream into the caller's HBM rings and stay there
        r.log = a->log_dev; r.log_count = a->log_count_dev;
        r.log_cap = a->log_cap_dev; r.n_log_reps = a->n_log_reps_dev < n ? a->n_log_reps_dev : n;
    }
    r.stream = st;
    CU(cudaEventRecord(h->ev0, st));
    int rc = msim_run(h, &r);
    if (rc != 0) return rc;
    CU(cudaEventRecord(h->ev1, st));
    if (a->stats_dev) {
        rc = msim_reduce_metrics(h, r.acc_sum, r.acc_cnt, r.status, n, a->stats_dev, st);
        if (rc != 0) return rc;
    }
    if (a->acc_sum) CU(cudaMemcpyAsync(a->acc_sum, r.acc_sum, (size_t)n * L.K * 8, cudaMemcpyDeviceToHost, st));
    if (a->acc_cnt) CU(cudaMemcpyAsync(a->acc_cnt, r.acc_cnt, (size_t)n * L.K * 4, cudaMemcpyDeviceToHost, st));
    if (a->n_events) CU(cudaMemcpyAsync(a->n_events, r.n_events, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    if (a->n_timeouts) CU(cudaMemcpyAsync(a->n_timeouts, r.n_timeouts, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (a->status) CU(cudaMemcpyAsync(a->status, r.status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (logs) {
        CU(cudaMemcpyAsync(a->log, r.log, (size_t)nlog * (size_t)a->log_cap * 16, cudaMemcpyDeviceToHost, st));
        if (a->log_count) CU(cudaMemcpyAsync(a->log_count, r.log_count, (size_t)nlog * 4, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    if (a->kernel_ms) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        *a->kernel_ms = ms;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int msim_test_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int64_t n) {
    if (n <= 0) return 0;
    uint32_t *dc = nullptr, *dk = nullptr, *dout = nullptr;
    CU(cudaMalloc(&dc, n * 16)); CU(cudaMalloc(&dk, n * 8)); CU(cudaMalloc(&dout, n * 16));
    CU(cudaMemcpy(dc, ctr, n * 16, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dk, key, n * 8, cudaMemcpyHostToDevice));
    int rc = launch_test_philox(dc, dk, dout, n);
    if (rc == 0) CU(cudaMemcpy(out, dout, n * 16, cudaMemcpyDeviceToHost));
    cudaFree(dc); cudaFree(dk); cudaFree(dout);
    return rc == 0 ? 0 : fail(rc, "philox test launch failed");
}

extern "C" int msim_test_variates(const msim_dist_t* dist, uint64_t seed, int32_t stream_id, int32_t batch_count,
                                  double* out, int64_t n) {
    if (!dist || !out || n <= 0) return fail(MSIM_E_INVALID, "bad argument");
    double* dout = nullptr;
    CU(cudaMalloc(&dout, n * 8));
    int rc = launch_test_variates(to_ddist(*dist), seed, stream_id, batch_count, dout, n);
    if (rc == 0) CU(cudaMemcpy(out, dout, n * 8, cudaMemcpyDeviceToHost));
    cudaFree(dout);
    return rc == 0 ? 0 : fail(rc, "variates test launch failed");
}
