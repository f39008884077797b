// msim_kernel.cu -- the batched factory-replication kernel for B200 (sm_100a).
//
// One WARP owns one replication.  Everything the reference keeps in SimPy objects lives in a per-warp slab of
// shared memory (layout: msim_device.cuh `Lay`, sized by msim_api.cu):
//   * event calendar  : one (time f32, eid) slot per static process that can sleep on a Timeout
//                       (R machines, P demand generators, warm-up, DBR tick) + per-order slots for onDue
//                       delivery timers / DBR rope delays; the next event is picked by a warp-wide min-reduction
//                       over (time, eid) -- SimPy's heap order (time, priority, eid) restricted to Timeouts;
//   * zero-delay events: SimPy schedules every put/get/request completion as a heap event at `now`.  Because
//                       eids only grow, those behave as two FIFOs (URGENT = process Initialize, NORMAL = the
//                       rest).  Lane 0 drains them literally, one micro-event at a
//                       time, so the (time, priority, eid) order -- and with it every tie-break -- is
//                       reproduced by construction, and the count of pops equals Environment.step() calls;
//   * stores           : resource_input / resource_output as linked lists through the order pool,
//                       processing/transport as single slots, FG/WIP containers as float levels + a FIFO of
//                       blocked getters (head-of-line blocking, SimPy Container semantics);
//   * log staging      : 32 x 128-bit records, flushed by the whole warp with one st.global.v4 per lane into
//                       the replication's HBM ring; the same flush folds the records into per-ring
//                       (sum f64, count) accumulators (warm-up filtering happens at emission, as in the
//                       reference).
// Entity state that only lane 0 touches is an array of small structs per entity kind -- resource 28 B (at slab
// offset 0), product 24 B, production order 12 / 16 B, demand order 12 B + 2 link bytes -- so a handler pays one
// multiply-add per ENTITY and reaches its 3-8 fields with immediate offsets; what the whole warp reads (calendar,
// per-order timers, event FIFOs, log staging, variate buffer) is structure-of-arrays, one slot per lane.
// Scheduling rules (FIFO / max-priority / DBR / EDD) are template parameters, so are the variate source
// (Philox4x32-10 keyed by replication/stream/draw index, or replay of host-recorded tapes) and the option flags
// (queue records, tape recording).  Scenario tables (routes, distributions) travel in the kernel parameters and are
// read with LDC; they use no shared memory.  Lane 0's loop state is split: what every micro-event touches stays in
// registers, the rest lives in the slab (`Scal`).
// The kernel is instruction-fetch bound (profiles/README.md): code size and the density of the hot path decide its
// speed, which is why cold paths are out of line and options are compile-time.
//
// Reference anchors are given next to each handler as file:line of /root/reference/src/manusim/factory_sim.py.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdlib.h>
#include "msim_device.cuh"
#include "philox.cuh"

// ---- build switches (tools/ab_probe.py builds -D variants next to libmsim.so and times them on identical
// trajectories; every variant is first checked bit for bit on the CPU-fiber emulation).  Measured on a B200 in round 2
// (profiles/r02_ab_probe.jsonl) and folded in: the shared out-of-line "push the successor and leave" exit of the
// chained handlers (+10 %), the OPTS = 0 kernels (+15..23 %); measured and dropped: per-ring-group log flush (-5 %),
// cached earliest onDue / rope timer (-5 % / 0: the extra code costs more than the scan it saves), counting calendar
// pops at the end (0).  (Byte-identical SASS, hence never timed: __builtin_expect on the error / option branches, an
// equality chain on the hottest micro-events in front of the handler switch.)
#ifndef MSIM_ALL_SHAPES    // 1: also compile the 64- and 48-register shapes of the lean kernels (sweeps); default: 56 only
#define MSIM_ALL_SHAPES 0
#endif
#ifndef MSIM_X_F32CELL     // 1 = NEGATIVE CONTROL for tests/test_emu_kernel.py: order the DBR calendar by float32 time alone
#define MSIM_X_F32CELL 0   //     (the round-1 behaviour; golden dbr_f64_order must fail with it)
#endif
// Option flags as a template parameter: OPTS bit 0 = the scenario may log `queue` records (config log_queues), bit 1 =
// the run may record variate tapes (rec_cap > 0).  OPTS = 3 is the general kernel; OPTS = 0 drops both branches from
// every handler (-8 % code) and is what a run that uses neither option launches.
// SHAPE = the launch bounds, i.e. the register budget against the resident warps (profiles/r02_launch_shape_sweep_c.jsonl:
// at EQUAL residency the 64-register build is 31 % faster than the 48-register one, whose spills sit on lane 0's
// critical path; 64 registers at 32 warps per SM beat 48 registers at 40 by 15 %):
//   0: __launch_bounds__(256, 4) -> 64 registers, <= 32 warps per SM   (-DMSIM_ALL_SHAPES=1 only)
//   1: __launch_bounds__(192, 6) -> 56 registers, <= 36 warps per SM   (the build: best on every measured workload)
//   2: __launch_bounds__(256, 5) -> 48 registers, <= 40 warps per SM   (-DMSIM_ALL_SHAPES=1 only; the round-1 shape)
#define LOGQ (((OPTS) & 1) && p.log_queues)
#define RECORDING (((OPTS) & 2) && p.a.rec_cap > 0)
#define NOT_RECORDING (!((OPTS) & 2) || p.a.rec_cap == 0)
#define PUSH_AND_LEAVE(ev) { tail_ev = (ev); goto L_TAIL_PUSH; }
namespace msim {

#define FULLMASK 0xFFFFFFFFu

// struct layouts inside the slab (byte offsets; msim_api.cu sizes the slab with the same strides, msim_device.cuh)
enum { kRes_utf = 0, kRes_tbf = 4, kRes_setup = 8, kRes_start = 12, kRes_busy = 16, kRes_cur = 20, kRes_tcur = 21,
       kRes_inh = 22, kRes_int = 23, kRes_outh = 24, kRes_outt = 25 };
enum { kPrd_fg = 0, kPrd_wip = 4, kPrd_fgc = 8, kPrd_dqty = 12, kPrd_ddue = 16, kPrd_obh = 20, kPrd_obt = 21,
       kPrd_fgqh = 22, kPrd_fgqt = 23 };
enum { kPo_rel = 0, kPo_qty = 4, kPo_prod = 8, kPo_done = 9, kPo_next = 10, kPo_loc = 11, kPo_prio = 12, kPo_ccr = 12 };
// demand orders are two struct arrays (three floats; two link bytes): one 16-byte struct would waste 2 bytes per
// order, i.e. a CTA's worth of occupancy at the pool sizes the 20-resource job shop needs
enum { kDo_arr = 0, kDo_due = 4, kDo_qty = 8, kDo_prod = 0, kDo_next = 1 };
static_assert(kAosResStride >= 26 && kAosPrdStride >= 24 && kAosDoStride == 12 && kAosDoLinkStride == 2,
              "struct strides (msim_device.cuh)");
#define kDoBase_arr p.lay.do_arr
#define kDoBase_due p.lay.do_arr
#define kDoBase_qty p.lay.do_arr
#define kDoBase_prod p.lay.do_prod
#define kDoBase_next p.lay.do_prod
#define kDoStr_arr kAosDoStride
#define kDoStr_due kAosDoStride
#define kDoStr_qty kAosDoStride
#define kDoStr_prod kAosDoLinkStride
#define kDoStr_next kAosDoLinkStride

// entity fields: resource structs start at slab offset 0; the product / order struct arrays start where the layout
// table keeps the first field of the SoA slab (p_fg, po_rel, do_arr; do_prod for the demand-order link bytes)
#define ARR(T, field) (reinterpret_cast<T*>(sb + p.lay.field))
#define TAB(T, field) (reinterpret_cast<const T*>(p.tblob + p.lay.field))
// fields of the constant part of the layout (msim_device.cuh `CL`): immediates in the lean kernels (CLAY), `Lay` otherwise
#define LC(field) (CLAY ? CL::field(POLICY) : p.lay.field)
#define LCN(field) (CLAY ? CL::field : p.lay.field)               // NT, NQ
#define LCUQ (CLAY ? CL::UQ(POLICY) : p.lay.UQ)
#define ARRC(T, field) (reinterpret_cast<T*>(sb + LC(field)))
#define TABC(T, field) (reinterpret_cast<const T*>(p.tblob + LC(field)))
#define RES(T, f, i) (*reinterpret_cast<T*>(sb + (uint32_t)(i) * kAosResStride + kRes_##f))
#define PRD(T, f, i) (*reinterpret_cast<T*>(sb + LC(p_fg) + (uint32_t)(i) * kAosPrdStride + kPrd_##f))
#define PORD(T, f, i) (*reinterpret_cast<T*>(sb + LC(po_rel) + (uint32_t)(i) * kPoStride + kPo_##f))
#define DORD(T, f, i) (*reinterpret_cast<T*>(sb + kDoBase_##f + (uint32_t)(i) * kDoStr_##f + kDo_##f))
#define PO_NEXT (sb + LC(po_rel) + kPo_next)
#define DO_NEXT (sb + p.lay.do_prod + kDo_next)

// Launch syntax and the dynamic shared-memory declaration go through two macros so that the host-side SIMT emulator
// under tests/emu (test infrastructure: runs THIS source on CPU fibers for sanitizer and parity checks, never part of
// the product) can substitute its own.
#ifndef MSIM_LAUNCH
#define MSIM_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MSIM_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

enum Op : uint32_t {
    OP_NONE = 0,
    OP_TO_WARMUP, OP_TO_DEMAND, OP_PUT_INBOUND, OP_SCHED_GOT, OP_PUT_OUTBOUND,
    OP_INIT_RELEASE, OP_REL_PUT1, OP_REL_PUT2,
    OP_MACH_GOT, OP_MACH_PUTPROC, OP_MACH_REQ, OP_TO_SETUP, OP_TO_PROC, OP_TO_TTR, OP_MACH_GOTPROC, OP_MACH_PUTOUT,
    OP_TRANS_GOT, OP_TRANS_PUT, OP_TRANS_GOT2_FINAL, OP_TRANS_GOT2_NEXT, OP_TRANS_FGPUT, OP_TRANS_WIPGET,
    OP_TRANS_PUTIN,
    OP_DELIV_GOT, OP_INIT_SHIP, OP_TO_SHIP, OP_SHIP_GOT,
    // DBR (examples/toc_dbr/simulation.py)
    OP_TO_TICK, OP_INIT_DORDER, OP_TO_DORDER, OP_SEED_PUT, OP_FIN_PUT, OP_DRAIN_GOT,
    OP_HOP,   // not a SimPy event: re-queues itself b-1 more times, then queues TO_SETUP of resource a (see OP_MACH_REQ)
};

enum Loc : uint8_t { LOC_NONE = 0, LOC_IN = 1, LOC_PROC = 2, LOC_OUT = 3, LOC_TRANS = 4 };
enum Req : int { REQ_ADVANCE = 0, REQ_FLUSH = 1, REQ_DONE = 2, REQ_SELECT = 3 };
// lane 0 "attention" bits: the fast path of the event loop (pop NORMAL fifo, dispatch) runs while this word is zero
enum { A_PENDING = 1, A_URGENT = 2, A_SAME = 4, A_FLUSH = 8, A_STATUS = 16, A_SELECT = 32 };
// DBR per-resource flags: python-typed (int / float64) values of the dynamically typed clock (Appendix B.2)
enum { F_TIMER_PY = 1, F_START_PY = 2, F_UTF_PY = 4 };

// product / resource / general metric indices (enum order of metrics.py:43-67)
enum { PM_ONTIME = 0, PM_LATE, PM_LOST, PM_FLOW, PM_LEAD, PM_TARDY, PM_EARLY, PM_WIP, PM_FG, PM_INV, PM_RELEASED };
enum { RM_UTIL = 0, RM_BREAKDOWN, RM_SETUP, RM_QUEUE };
enum { GM_WIP = 0, GM_FG, GM_INV };

__device__ __forceinline__ uint32_t EV(uint32_t op, uint32_t a, uint32_t b) { return op | (a << 8) | (b << 16); }
// DBR only (python-typed clock): Environment.step() sets _now to the popped entry's own time
// OBJECT, so the clock's type is a property of every queued event -- an event scheduled with the default delay 0
// inherits the type the clock had when it was scheduled, a Timeout with a numpy float32 delay is float32-typed even
// when now + delay rounds back to `now`.  Bit 24 of the event word: the clock is python-typed when this event pops.
constexpr uint32_t EV_PY = 1u << 24;

// round(np.float32 x, 6)  (factory_sim.py:284,443)
__device__ __forceinline__ float round6(float x) { return __fdiv_rn(rintf(__fmul_rn(x, 1e6f)), 1e6f); }
// round(np.float32 x, 3)  (toc_dbr/simulation.py:229-236)
__device__ __forceinline__ float round3(float x) { return __fdiv_rn(rintf(__fmul_rn(x, 1e3f)), 1e3f); }
// round(python float x, 6): CPython rounds the exact decimal value.  p + err is x*1e6 without rounding error (one FMA);
// only when p sits exactly on a half-way point can the rounding of the product have hidden on which side the true
// value lies -- then err decides (err == 0: ties to even).  Matches round(x, 6) on 1.6e6 adversarial near-tie inputs
// where rint(x*1e6)/1e6 alone differs on 13 % of them.
__device__ __forceinline__ double pyround6(double x) {
    const double p = __dmul_rn(x, 1e6);
    const double err = fma(x, 1e6, -p);
    double q = rint(p);
    const double fl = floor(p);
    if (p - fl == 0.5 && err != 0.0) q = err > 0.0 ? fl + 1.0 : fl;
    return __ddiv_rn(q, 1e6);
}

// Slow path of a Philox-mode draw, one out-of-line copy per kernel: used by the cold call sites (boot, breakdowns,
// demand triplets) and as the fallback of the inlined fast paths (buffer miss, Marsaglia-Tsang squeeze failure,
// shape < 1, batches of more than one sample, tape recording).  Bumps the stream's draw index itself.
// count < 0: scalar draw (DistributionGenerator.random_number, clamped at 0); count >= 0: batch sum.
template <int OPTS>
__device__ __noinline__ double philox_draw_slow(const KParams& p, Scal* sc, const float2* rb, long long rep, int stream,
                                                const DDist* d, int count) {
    const uint32_t idx = sc->draw[stream]++;
    double out;
    if (d->kind == MSIM_DIST_CONSTANT || count == 0) {
        out = count < 0 ? (double)(d->d0 > 0.0f ? d->d0 : 0.0f) : (double)count * d->raw0;
    } else {
        XU r;
        const bool buffered = stream == 1 || stream == 2;      // only the per-operation streams are precomputed
        uint32_t off = buffered ? idx - sc->rb_base[buffered ? stream : 1] : 0xFFFFFFFFu;
        if (off < (uint32_t)kXuBuf) {
            float2 v = rb[(stream - 1) * kXuBuf + off];
            r.x = v.x; r.u = v.y;
        } else {
            r = raw_xu(sc->k0, sc->k1, (uint32_t)stream, idx, 0u);
        }
        PhiloxStream s{sc->k0, sc->k1, (uint32_t)stream, idx};
        float v;
        if (count > 1) {
            v = batch_many(s, d->kind, d->d0, d->d1, count, r);
        } else {
            v = sample_from_xu(s, d->kind, d->d0, d->d1, d->mt_d, d->mt_c, r);
            if ((count < 0 || d->kind == MSIM_DIST_NORMAL) && !(v > 0.0f)) v = 0.0f;   // utils.py:67 / :97-98
        }
        out = (double)v;
    }
    if (RECORDING && (long long)idx < p.a.rec_cap) {
        if (stream == 2) p.a.rec_C[rep * p.a.rec_cap + idx] = out;
        else (stream == 0 ? p.a.rec_A : (stream == 1 ? p.a.rec_B : p.a.rec_D))[rep * p.a.rec_cap + idx] = (float)out;
    }
    return out;
}

// Compile-time tuning per kernel family.  The kernel is instruction-fetch bound and its speed responds to code LAYOUT more
// than to instruction count: semantically equal variants differ by up to +-14 % on a B200, not monotonically in code size
// (profiles/r02_ab_variants_*.jsonl; every variant bit-identical on the CPU-fiber emulation first).  What each family
// gets is what measured best for it:
//   loop2  calendar pop handed over through the attention word + one flush site   FIFO -4 %, max-priority +9 %, DBR +6 %
//   emit2  record sequence numbers filled in by the flush, one staging check per dispatched event   } together, with the
//   mode3  `instantly` delivery compiled as its own kernel instead of a run-time branch in asReady  } rolled Philox rounds:
//                                                                         max-priority +14 %, FIFO -10 %, DBR 0
template <int POLICY>
struct Tune {
    static constexpr bool prio = POLICY == MSIM_POLICY_MAX_PRIORITY || POLICY == MSIM_POLICY_EDD;
    static constexpr bool loop2 = POLICY != MSIM_POLICY_FIFO;
    static constexpr bool emit2 = prio;
    static constexpr bool mode3 = prio;
};

// MODE: 0 = the handle delivers asReady or instantly (told apart at run time, a few instructions in INIT_SHIP), 1 = onDue
// (per-order delivery timers in the calendar), 2 = decided at run time from KParams (the general OPTS = 3 kernels).  The
// kernel is instruction-fetch bound: compiling the other mode's code out is worth more than the branch it removes.
// replay mode: next value of tape `stream` (A, B, D); an exhausted tape ends the replication with a status code
__device__ __noinline__ float replay_scalar(Scal* sc, int stream) {
    const uint32_t idx = sc->draw[stream]++;
    if (idx >= sc->tlen[stream]) { sc->x_status = MSIM_ST_TAPE_EXHAUSTED; return 0.0f; }
    return __ldg(reinterpret_cast<const float*>(sc->tbase[stream]) + idx);
}

template <int POLICY, int RNG, int OPTS, int MODE>
struct Sim {
    const KParams& p;
    unsigned char* sb;        // this warp's slab
    Scal* sc;
    long long rep;
    // ---- lane-0 state kept in registers (touched by every micro-event)
    float now;
    uint32_t warm;
    uint32_t nqh, nqt;
    uint32_t nstage;
    uint32_t nev;
    uint32_t attn;
    uint32_t force;           // 1 while the machine chain MACH_GOT -> PUTPROC -> REQ runs ahead of the transport chain
    // ---- lane-0 state that lives in the shared-memory slab (touched rarely; keeps the register budget at the
    //      48-warps-per-SM level): references into Scal
    uint32_t &uqh, &uqt, &logc, &tseq, &same_left, &pending, &rb_need, &now_py, &sel_req;
    int& status;
    double& now_d;            // DBR only: exact python-typed clock (now is its float32 view)
    static constexpr bool DBR = POLICY == MSIM_POLICY_DBR;
    static constexpr bool PRIO = POLICY == MSIM_POLICY_MAX_PRIORITY || POLICY == MSIM_POLICY_EDD;
    static constexpr bool EMIT2 = Tune<POLICY>::emit2;
    static constexpr bool CLAY = OPTS == 0;                  // lean kernels: constant layout (msim_device.cuh `CL`)
    __device__ __forceinline__ bool on_due() const { return MODE == 2 ? p.delivery_mode == MSIM_DELIVERY_ON_DUE : MODE == 1; }
    __device__ __forceinline__ bool instantly() const {
        if constexpr (Tune<POLICY>::mode3) return MODE == 2 ? p.delivery_mode == MSIM_DELIVERY_INSTANTLY : MODE == 3;
        else return !on_due() && p.delivery_mode == MSIM_DELIVERY_INSTANTLY;
    }

    __device__ __forceinline__ Sim(const KParams& p_, unsigned char* sb_, Scal* sc_)
        : p(p_), sb(sb_), sc(sc_), uqh(sc_->x_uqh), uqt(sc_->x_uqt), logc(sc_->x_logc), tseq(sc_->x_tseq),
          same_left(sc_->x_same_left), pending(sc_->x_pending), rb_need(sc_->x_rb_need), now_py(sc_->x_now_py),
          sel_req(sc_->x_sel_req), status(sc_->x_status), now_d(sc_->x_now_d) {}

    static constexpr int kPoStride = aos_po_stride(POLICY), kDoStride = kAosDoLinkStride;

    __device__ __forceinline__ int ringP(int m, int prod) const { return m * p.lay.P + prod; }
    __device__ __forceinline__ int ringR(int m, int r) const { return 11 * p.lay.P + m * p.lay.R + r; }
    __device__ __forceinline__ int ringG(int g) const { return 11 * p.lay.P + 4 * p.lay.R + g; }

    __device__ __forceinline__ void fail(int code) { status = code; attn |= A_STATUS; }

    // ------------------------------------------------------------------ event queues
    __device__ __forceinline__ void pushN_raw(uint32_t ev) {
        ARRC(uint32_t, nq)[nqt & (uint32_t)(LCN(NQ) - 1)] = ev;
        nqt++;       // overflow is checked once per dispatched event (drain); the ring index is masked, so a
    }                // transient overrun stays inside the slab and aborts the replication with a status code
    __device__ __forceinline__ void pushN(uint32_t ev) { pushN_raw(DBR ? ev | (now_py << 24) : ev); }   // inherits the clock type
    __device__ __forceinline__ void pushU(uint32_t ev) {
        if (DBR) ev |= now_py << 24;
        ARRC(uint32_t, uq)[uqt & (uint32_t)(LCUQ - 1)] = ev;
        uqt++;
        attn |= A_URGENT;
        if (uqt - uqh > (uint32_t)LCUQ) fail(MSIM_ST_FIFO_OVERFLOW);
    }
    // env.timeout(delay) issued by a static process that owns calendar slot `slot`
    // `f32`: the delay is float32-typed (false: python int, e.g. the warm-up instant or a zero setup)
    __device__ __forceinline__ void timer(int slot, float t, uint32_t ev, bool f32 = true) {
        if (t == now) {
            if (DBR && f32) pushN_raw(ev); else pushN(ev);   // lands on the current timestamp: larger eid than everything pending
        } else {
            ARRC(float, tm_time)[slot] = t;
            ARRC(uint32_t, tm_eid)[slot] = tseq++;
            ARRC(uint32_t, tm_ev)[slot] = ev;     // what the calendar hands back when this slot fires
        }
    }

    // ------------------------------------------------------------------ variates
    // Inlined fast path shared by the two per-operation draws (setup, processing time): the (x,u) pair is already
    // in the warp-precomputed buffer and the distribution is normal, or gamma/erlang with shape >= 1 accepted by
    // the Marsaglia-Tsang squeeze.  Anything else goes to the out-of-line slow path, which yields the same value.
    __device__ __forceinline__ bool fast_sample(int stream, const DDist& d, float& v) {
        const uint32_t idx = sc->draw[stream];
        const uint32_t off = idx - sc->rb_base[stream];
        if (off >= (uint32_t)kXuBuf - kXuRefillAt) rb_need |= 1u << stream;
        if (off >= (uint32_t)kXuBuf || RECORDING) return false;
        const float2 xu = ARRC(float2, rb)[(stream - 1) * kXuBuf + off];
        if (d.kind == MSIM_DIST_NORMAL) {
            v = d.d0 + d.d1 * xu.x;
        } else if ((d.kind == MSIM_DIST_GAMMA || d.kind == MSIM_DIST_ERLANG) && d.d0 >= 1.0f) {
            const float w = 1.0f + d.mt_c * xu.x;
            const float x2 = xu.x * xu.x;
            if (!(w > 0.0f) || !(xu.y < 1.0f - 0.0331f * x2 * x2)) return false;
            v = d.mt_d * (w * w * w) * d.d1;
        } else {
            return false;
        }
        sc->draw[stream] = idx + 1u;
        return true;
    }
    __device__ __forceinline__ float draw_scalar_hot(int stream, const DDist& d) {
        if (RNG == MSIM_RNG_REPLAY) return draw_scalar(stream, d);
        float v;
        if (fast_sample(stream, d, v)) return v > 0.0f ? v : 0.0f;
        return (float)philox_draw_slow<OPTS>(p, sc, ARRC(float2, rb), rep, stream, &d, -1);
    }
    // scalar draws of the COLD call sites (demand triplets, breakdowns, boot): one call, no inlined fast path -- the kernel
    // is instruction-fetch bound and these sites run once per order or less (constants are handled inside the callee)
    __device__ __forceinline__ float draw_scalar(int stream, const DDist& d) {
        if (RNG == MSIM_RNG_REPLAY) {
            uint32_t idx = sc->draw[stream]++;
            if (idx >= sc->tlen[stream]) { fail(MSIM_ST_TAPE_EXHAUSTED); return 0.0f; }
            return __ldg(reinterpret_cast<const float*>(sc->tbase[stream]) + idx);
        }
        if (d.kind == MSIM_DIST_CONSTANT && NOT_RECORDING) {   // max(0, np.float32(params[0])): no generator involved
            sc->draw[stream]++;
            return d.d0 > 0.0f ? d.d0 : 0.0f;
        }
        return (float)philox_draw_slow<OPTS>(p, sc, ARRC(float2, rb), rep, stream, &d, -1);
    }
    __device__ __forceinline__ double draw_batch(const DDist& d, int count) {
        if (RNG == MSIM_RNG_REPLAY) {
            uint32_t idx = sc->draw[2]++;
            if (idx >= sc->tlen[2]) { fail(MSIM_ST_TAPE_EXHAUSTED); return 0.0; }
            return __ldg(reinterpret_cast<const double*>(sc->tbase[2]) + idx);
        }
        float v;
        if (count == 1 && fast_sample(2, d, v)) return (double)((d.kind == MSIM_DIST_NORMAL && !(v > 0.0f)) ? 0.0f : v);
        return philox_draw_slow<OPTS>(p, sc, ARRC(float2, rb), rep, 2, &d, count < 0 ? 0 : count);
    }

    // ------------------------------------------------------------------ records (Logger.log, logs.py:44-66)
    __device__ __forceinline__ void emit(int ring, float t, float v) {
        if constexpr (EMIT2) {
            // (the record's sequence number is its position: the flush fills it in; whether the staging area needs flushing
            //  is looked at once per dispatched event in drain(), not per record -- a dispatch emits at most 7 records)
            ARRC(uint4, stage)[nstage] = make_uint4(__float_as_uint(t), __float_as_uint(v), (uint32_t)ring, 0u);
            nstage++;
        } else {
            ARRC(uint4, stage)[nstage] = make_uint4(__float_as_uint(t), __float_as_uint(v), (uint32_t)ring, logc + nstage);
            nstage++;
            if (nstage >= (uint32_t)kStageFlushAt) attn |= A_FLUSH;
        }
    }
    // _log_inventory_metrics, factory_sim.py:551-601 (caller checked warm)
    __device__ __forceinline__ void inventory(int prod, bool with_wip, bool with_fg) {
        float w = PRD(float, wip, prod);
        float f = PRD(float, fg, prod);
        if (with_wip) {
            emit(ringP(PM_WIP, prod), now, w);
            emit(ringG(GM_WIP), now, sc->tot_wip);
        }
        if (with_fg) {
            float old = PRD(float, fgc, prod);
            PRD(float, fgc, prod) = f;
            sc->fgc_sum = __fadd_rn(sc->fgc_sum, __fsub_rn(f, old));
            emit(ringP(PM_FG, prod), now, f);
            emit(ringG(GM_FG), now, sc->fgc_sum);
        }
        emit(ringP(PM_INV, prod), now, __fadd_rn(w, f));
        emit(ringG(GM_INV), now, __fadd_rn(sc->tot_wip, sc->fgc_sum));
    }
    // _log_delivery_performance + _log_lead_time, factory_sim.py:606-654
    __device__ __forceinline__ void delivery_records(int d, int prod, bool delivered) {
        float q = DORD(float, qty, d);
        float due = DORD(float, due, d);
        if (!delivered || now == 0.0f) {          // `if not delivered` (also true for delivered == 0, quirk E6)
            emit(ringP(PM_LOST, prod), now, q);
        } else if (now <= due) {
            emit(ringP(PM_ONTIME, prod), now, q);
            emit(ringP(PM_EARLY, prod), now, __fsub_rn(due, now));
        } else {
            emit(ringP(PM_LATE, prod), now, q);
            emit(ringP(PM_TARDY, prod), now, __fsub_rn(now, due));
        }
        emit(ringP(PM_LEAD, prod), now, __fsub_rn(now, DORD(float, arr, d)));
    }
    // resource `queue` records, factory_sim.py:227-242 / 343-358
    __device__ __forceinline__ void queue_log_add(int r, float q) {
        float last = ARR(float, r_lastq)[r];
        float val;
        if (isnan(last)) {
            val = ARR(float, r_qq)[r];
        } else {
            val = __fadd_rn(last, q);
            ARR(float, r_qq)[r] = val;
        }
        emit(ringR(RM_QUEUE, r), now, val);
        ARR(float, r_lastq)[r] = val;
    }

    // ------------------------------------------------------------------ pools and stores
    __device__ __forceinline__ uint32_t alloc_po() {
        uint32_t i = sc->po_free;
        if (i == kNone) { fail(MSIM_ST_PO_POOL_OVERFLOW); return 0; }
        sc->po_free = PORD(uint8_t, next, i);
        if (DBR) ARR(float, po_tt)[i] = CUDART_INF_F;
        return i;
    }
    __device__ __forceinline__ void free_po(uint32_t i) {
        PORD(uint8_t, next, i) = sc->po_free;
        PORD(uint8_t, loc, i) = LOC_NONE;
        sc->po_free = (uint8_t)i;
    }
    __device__ __forceinline__ uint32_t alloc_do() {
        uint32_t i = sc->do_free;
        if (i == kNone) { fail(MSIM_ST_DO_POOL_OVERFLOW); return 0; }
        sc->do_free = DORD(uint8_t, next, i);
        return i;
    }
    __device__ __forceinline__ void free_do(uint32_t i) {
        DORD(uint8_t, next, i) = sc->do_free;
        sc->do_free = (uint8_t)i;
    }
    // singly linked FIFOs through the `next` byte of the order pools; `next0` addresses entry 0's byte and STRIDE is
    // the distance to the next entry's (1 in the structure-of-arrays slab, the struct size in the AoS slab)
    template <int STRIDE>
    __device__ __forceinline__ void list_append(uint8_t* head, uint8_t* tail, uint8_t* next0, uint32_t i) {
        next0[i * STRIDE] = kNone;
        uint32_t t = *tail;
        if (t == kNone) *head = (uint8_t)i; else next0[t * STRIDE] = (uint8_t)i;
        *tail = (uint8_t)i;
    }
    template <int STRIDE>
    __device__ __forceinline__ uint32_t list_pop(uint8_t* head, uint8_t* tail, const uint8_t* next0) {
        uint32_t h = *head;
        uint32_t n = next0[h * STRIDE];
        *head = (uint8_t)n;
        if (n == kNone) *tail = kNone;
        return h;
    }
    // Container._trigger_get on finished_goods[prod]: serve blocked getters from the head, stop at the first
    // one that cannot be satisfied (SimPy Container._do_get returns None => head-of-line blocking)
    __device__ __forceinline__ void serve_fg(int prod) {
        uint32_t h = PRD(uint8_t, fgqh, prod);
        if (h == kNone) return;
        float lvl = PRD(float, fg, prod);
        while (h != kNone) {
            float q = DORD(float, qty, h);
            if (!(lvl >= q)) break;
            lvl = __fsub_rn(lvl, q);
            pushN(EV(OP_SHIP_GOT, prod, h));
            h = DORD(uint8_t, next, h);
        }
        PRD(float, fg, prod) = lvl;
        PRD(uint8_t, fgqh, prod) = (uint8_t)h;
        if (h == kNone) PRD(uint8_t, fgqt, prod) = kNone;
    }
    // finished_goods[prod].get(quantity) issued by a delivery process (ContainerGet.__init__)
    __device__ __forceinline__ void fg_get(int d, int prod) {
        if (!(DORD(float, qty, d) > 0.0f)) { fail(MSIM_ST_AMOUNT_NONPOSITIVE); return; }
        list_append<kDoStride>(&PRD(uint8_t, fgqh, prod), &PRD(uint8_t, fgqt, prod), DO_NEXT, d);
        serve_fg(prod);
    }
    // FilterStore._trigger_get on resource_input[r] when a put event is processed
    __device__ __forceinline__ void serve_mach(int r) {
        if ((sc->mach_wait >> r) & 1u) {
            if (RES(uint8_t, inh, r) != kNone) {
                uint32_t po = list_pop<kPoStride>(&RES(uint8_t, inh, r), &RES(uint8_t, int, r), PO_NEXT);
                PORD(uint8_t, loc, po) = LOC_NONE;
                sc->mach_wait &= ~(1u << r);
                pushN(EV(OP_MACH_GOT, r, po));
            }
        }
    }
    // _transportation loop top: resource_output[r].get()   factory_sim.py:298-302
    __device__ __forceinline__ void trans_get(int r) {
        if (RES(uint8_t, outh, r) != kNone) {
            uint32_t po = list_pop<kPoStride>(&RES(uint8_t, outh, r), &RES(uint8_t, outt, r), PO_NEXT);
            PORD(uint8_t, loc, po) = LOC_NONE;
            pushN(EV(OP_TRANS_GOT, r, po));
        } else {
            sc->trans_wait |= 1u << r;
        }
    }
    // scheduler loop top: inbound_demand_orders.get()      factory_sim.py:182
    __device__ __forceinline__ void sched_get() {
        if (sc->inb_head != kNone) {
            uint32_t d = list_pop<kDoStride>(&sc->inb_head, &sc->inb_tail, DO_NEXT);
            pushN(EV(OP_SCHED_GOT, 0, d));
        } else {
            sc->sched_wait = 1;
        }
    }
    // _delivery_orders loop top: outbound[prod].get()       factory_sim.py:469-471
    __device__ __forceinline__ void deliv_get(int prod) {
        if (PRD(uint8_t, obh, prod) != kNone) {
            uint32_t d = list_pop<kDoStride>(&PRD(uint8_t, obh, prod), &PRD(uint8_t, obt, prod), DO_NEXT);
            pushN(EV(OP_DELIV_GOT, prod, d));
        } else {
            sc->deliv_wait |= 1u << prod;
        }
    }
    // _generate_demand_orders loop top: three draws then timeout(freq)   factory_sim.py:154-164
    __device__ __forceinline__ void demand_draw(int prod) {
        float gap = draw_scalar(0, TAB(DDist, t_freq)[prod]);
        float q = rintf(draw_scalar(0, TAB(DDist, t_qty)[prod]));     // round(x, 0): ties to even
        float due = draw_scalar(0, TAB(DDist, t_due)[prod]);
        PRD(float, dqty, prod) = q;
        PRD(float, ddue, prod) = due;
        timer(p.lay.R + prod, __fadd_rn(now, gap), EV(OP_TO_DEMAND, prod, 0));
    }
    // order_selection: policy is a compile-time device function
    // returns kNone unless `chain` is set and an order was picked: then the caller runs MACH_GOT itself
    __device__ __forceinline__ uint32_t mach_select(int r, bool chain = false) {
        uint32_t h = RES(uint8_t, inh, r);
        if (h == kNone) { sc->mach_wait |= 1u << r; return kNone; }   // plain get() on an empty queue
        uint32_t pick = h;
        if (DBR) {
            // buffer-penetration dispatch (toc_dbr/simulation.py:291-328) needs a scan of every in-flight order:
            // a single queued order is trivially the minimum, otherwise the whole warp computes the priorities
            if (PORD(uint8_t, next, h) != kNone) { sel_req = (uint32_t)r + 1u; attn |= A_SELECT; return kNone; }
            list_pop<kPoStride>(&RES(uint8_t, inh, r), &RES(uint8_t, int, r), PO_NEXT);
        } else if (PRIO) {
            // max(orders, key=priority): first maximum in queue order (simple_scheduler/simulation.py:36)
            float best = PORD(float, prio, h);
            uint32_t prev = kNone, pick_prev = kNone;
            for (uint32_t i = h; i != kNone; prev = i, i = PORD(uint8_t, next, i)) {
                if (PORD(float, prio, i) > best) { best = PORD(float, prio, i); pick = i; pick_prev = prev; }
            }
            if (pick != h) {      // unlink from the middle
                PORD(uint8_t, next, pick_prev) = PORD(uint8_t, next, pick);
                if (PORD(uint8_t, next, pick) == kNone) RES(uint8_t, int, r) = (uint8_t)pick_prev;
            } else {
                list_pop<kPoStride>(&RES(uint8_t, inh, r), &RES(uint8_t, int, r), PO_NEXT);
            }
        } else {
            list_pop<kPoStride>(&RES(uint8_t, inh, r), &RES(uint8_t, int, r), PO_NEXT);
        }
        PORD(uint8_t, loc, pick) = LOC_NONE;
        if (chain) return pick;
        pushN(EV(OP_MACH_GOT, r, pick));
        return kNone;
    }
    // _breakdowns entry: draw the repair time and sleep   factory_sim.py:267-276
    __device__ __forceinline__ void breakdown_start(int r) {
        RES(float, start, r) = now;                                      // breakdown_start, :272
        float ttr = draw_scalar(3, TAB(DDist, t_ttr)[r]);                  // :275
        if (DBR) ARR(uint8_t, r_flags)[r] &= ~F_TIMER_PY;
        timer(r, __fadd_rn(now, ttr), EV(OP_TO_TTR, r, 0), ttr > 0.0f);      // max(0, np.float32(v)) is the int 0 for v <= 0
    }
    // _production_system loop top   factory_sim.py:375-380
    __device__ __forceinline__ void mach_loop_top(int r) {
        if (RES(float, utf, r) >= RES(float, tbf, r)) { breakdown_start(r); return; }
        mach_select(r);
    }

    // True when the NORMAL event about to be pushed would be the very next one popped (nothing else is pending at
    // this timestamp: no urgent event, no calendar entry tied on `now`, no flush/dispatch request, empty fifo).
    // The handler may then run its successor directly (case fall-through) instead of pushing and re-dispatching it.
    __device__ __forceinline__ bool alone() const { return (attn | (nqh ^ nqt)) == 0u; }

    // ------------------------------------------------------------------ one micro-event
    __device__ __forceinline__ void dispatch(uint32_t ev) {
        if (DBR) now_py = (ev >> 24) & 1u;                       // _now takes the popped entry's type
        const uint32_t op = ev & 0xFFu;
        int a = (ev >> 8) & 0xFF;
        int b = (ev >> 16) & 0xFF;
        nev++;
        uint32_t tail_ev;
        switch (op) {
        case OP_TO_WARMUP:                                       // factory_sim.py:97-99
            warm = 1;
            nev++;                                               // process-end event
            break;
        case OP_TO_DEMAND: {                                     // :166-175
            uint32_t d = alloc_do();
            if (status) break;
            DORD(uint8_t, prod, d) = (uint8_t)a;
            DORD(float, qty, d) = PRD(float, dqty, a);
            DORD(float, due, d) = __fadd_rn(PRD(float, ddue, a), now);
            DORD(float, arr, d) = now;
            list_append<kDoStride>(&sc->inb_head, &sc->inb_tail, DO_NEXT, d);
            if (!alone()) PUSH_AND_LEAVE(EV(OP_PUT_INBOUND, a, 0))
            nev++;                                               // StorePut[inbound] popped right away
        }
        // fall through
        case OP_PUT_INBOUND:                                     // StorePut[inbound] processed
            if (sc->sched_wait && sc->inb_head != kNone) {
                uint32_t d = list_pop<kDoStride>(&sc->inb_head, &sc->inb_tail, DO_NEXT);
                sc->sched_wait = 0;
                pushN(EV(OP_SCHED_GOT, 0, d));
            }
            demand_draw(a);
            break;
        case OP_SCHED_GOT: {                                     // scheduler :183-193
            int prod = DORD(uint8_t, prod, b);
            if (!DBR) {           // DBR: _process_demandOrders only forwards (toc_dbr/simulation.py:408-412)
                uint32_t po = alloc_po();
                if (status) break;
                PORD(uint8_t, prod, po) = (uint8_t)prod;
                PORD(float, qty, po) = DORD(float, qty, b);
                if (POLICY == MSIM_POLICY_MAX_PRIORITY) PORD(float, prio, po) = 0.0f;
                if (POLICY == MSIM_POLICY_EDD) PORD(float, prio, po) = -DORD(float, due, b);
                pushU(EV(OP_INIT_RELEASE, 0, po));
            }
            list_append<kDoStride>(&PRD(uint8_t, obh, prod), &PRD(uint8_t, obt, prod), DO_NEXT, b);
            pushN(EV(OP_PUT_OUTBOUND, prod, 0));
            break;
        }
        case OP_PUT_OUTBOUND:                                    // StorePut[outbound[p]] processed
            if (((sc->deliv_wait >> a) & 1u) && PRD(uint8_t, obh, a) != kNone) {
                uint32_t d = list_pop<kDoStride>(&PRD(uint8_t, obh, a), &PRD(uint8_t, obt, a), DO_NEXT);
                sc->deliv_wait &= ~(1u << a);
                pushN(EV(OP_DELIV_GOT, a, d));
            }
            sched_get();
            break;
        case OP_INIT_RELEASE: {                                  // _release_order :195-211
            int prod = PORD(uint8_t, prod, b);
            int first = TABC(uint8_t, t_step_res)[TABC(uint16_t, t_route_start)[prod]];
            PORD(uint8_t, done, b) = 0;
            PORD(float, rel, b) = now;
            if (DBR) ARR(uint8_t, po_relpy)[b] = (uint8_t)now_py;
            list_append<kPoStride>(&RES(uint8_t, inh, first), &RES(uint8_t, int, first), PO_NEXT, b);
            PORD(uint8_t, loc, b) = LOC_IN;
            if (!alone()) PUSH_AND_LEAVE(EV(OP_REL_PUT1, first, b))
            a = first;
            nev++;                                               // StorePut[input of the first resource] popped right away
        }
        // fall through
        case OP_REL_PUT1: {                                      // :211-212
            serve_mach(a);
            float q = PORD(float, qty, b);
            if (!(q > 0.0f)) { fail(MSIM_ST_AMOUNT_NONPOSITIVE); break; }
            int prod = PORD(uint8_t, prod, b);
            PRD(float, wip, prod) = __fadd_rn(PRD(float, wip, prod), q);
            if (!alone()) PUSH_AND_LEAVE(EV(OP_REL_PUT2, a, b))
            nev++;                                               // ContainerPut[wip] popped right away
        }
        // fall through
        case OP_REL_PUT2: {                                      // :214-242
            float q = PORD(float, qty, b);
            int prod = PORD(uint8_t, prod, b);
            sc->tot_wip = __fadd_rn(sc->tot_wip, q);
            if (LOGQ) ARR(float, r_qq)[a] = __fadd_rn(ARR(float, r_qq)[a], q);
            if (warm) {
                emit(ringP(PM_RELEASED, prod), now, q);
                inventory(prod, true, false);
                if (LOGQ) queue_log_add(a, q);
            }
            nev++;                                               // process-end event
            break;
        }
        case OP_MACH_GOT:                                        // _production_system :380-395
        L_MACH_GOT: {
            if (LOGQ) {
                float q = PORD(float, qty, b);
                ARR(float, r_qq)[a] = __fsub_rn(ARR(float, r_qq)[a], q);
                if (warm) {
                    emit(ringR(RM_QUEUE, a), now, ARR(float, r_qq)[a]);
                    ARR(float, r_lastq)[a] = ARR(float, r_qq)[a];
                }
            }
            RES(uint8_t, cur, a) = (uint8_t)b;
            PORD(uint8_t, loc, b) = LOC_PROC;
            if (!(force || alone())) PUSH_AND_LEAVE(EV(OP_MACH_PUTPROC, a, 0))
            nev++;                                               // StorePut[processing] popped right away
        }
        // fall through
        case OP_MACH_PUTPROC: {                                  // :397-421 (setup drawn on every operation, quirk E1)
            float setup = draw_scalar_hot(1, TABC(DDist, t_setup)[a]);
            RES(float, setup, a) = setup;
            if (warm) emit(ringR(RM_SETUP, a), now, setup);
            if (!(force || alone())) PUSH_AND_LEAVE(EV(OP_MACH_REQ, a, 0))
            nev++;                                               // Request popped right away
        }
        // fall through
        case OP_MACH_REQ: {                                      // :423
            const uint32_t ahead = force;      // 1: MACH_GOT, PUTPROC and this event ran ahead of the transport chain
            force = 0;
            if (DBR) ARR(uint8_t, r_flags)[a] &= ~F_TIMER_PY;
            {
                const float t = __fadd_rn(now, RES(float, setup, a));
                const bool f32 = (RES(float, setup, a) > 0.0f);                  // a positive setup is a numpy float32, a zero one the int 0
                if (ahead && t == now) {
                    // The setup Timeout lands on the current timestamp, so the machine goes on within it -- and SimPy
                    // schedules that Timeout only after the first three transport events (TRANS_GOT, TRANS_PUT,
                    // GOT2_*) that this chain overtook.  OP_HOP re-queues itself behind each of them and then queues
                    // TO_SETUP exactly where SimPy has it (found by long-horizon replays: a zero setup followed by a
                    // zero-time operation emitted its utilisation record ahead of the transport chain's records).
                    pushN(EV(OP_HOP, a, 3));
                    break;
                }
                if (t != now || !alone()) { timer(a, t, EV(OP_TO_SETUP, a, 0), f32); break; }
                nev++;                                           // Timeout(0): lands on this timestamp and is popped right away
                if (DBR && f32) now_py = 0;                      // (a tiny positive setup can round back to `now`)
            }
        }
            // fall through
        case OP_TO_SETUP:
        {                                      // :425-440
            uint32_t po = RES(uint8_t, cur, a);
            int prod = PORD(uint8_t, prod, po);
            int step = TABC(uint16_t, t_route_start)[prod] + PORD(uint8_t, done, po);
            RES(float, start, a) = now;
            double total = draw_batch(TABC(DDist, t_step_dist)[step], (int)PORD(float, qty, po));
            if (total < 0.0) { fail(MSIM_ST_NEGATIVE_DELAY); break; }
            if (DBR) {
                uint8_t fl = ARR(uint8_t, r_flags)[a] & ~(F_TIMER_PY | F_START_PY);
                if (now_py) {
                    // python-typed clock + python float delay: the sum stays a python float (float64)
                    fl |= F_START_PY;
                    ARR(double, r_start_d)[a] = now_d;
                    double td = __dadd_rn(now_d, total);
                    if (td == now_d) {
                        ARR(uint8_t, r_flags)[a] = fl;
                        pushN(EV(OP_TO_PROC, a, 0));
                    } else {
                        ARR(double, r_tpy)[a] = td;
                        ARRC(float, tm_time)[a] = __double2float_rn(td);
                        ARRC(uint32_t, tm_eid)[a] = tseq++;
                        ARRC(uint32_t, tm_ev)[a] = EV(OP_TO_PROC, a, 0);
                        ARR(uint8_t, r_flags)[a] = fl | F_TIMER_PY;
                    }
                    break;
                }
                ARR(uint8_t, r_flags)[a] = fl;
            }
            timer(a, __fadd_rn(now, __double2float_rn(total)), EV(OP_TO_PROC, a, 0));
            break;
        }
        case OP_TO_PROC:
        {                                       // :442-451
            uint32_t po = RES(uint8_t, cur, a);
            float busy;
            if (DBR && now_py && (ARR(uint8_t, r_flags)[a] & F_START_PY)) {
                // both ends python-typed: float64 subtraction and python round (factory_sim.py:443)
                double bd = pyround6(__dsub_rn(now_d, ARR(double, r_start_d)[a]));
                busy = __double2float_rn(bd);
                if (ARR(uint8_t, r_flags)[a] & F_UTF_PY) {
                    double u = __dadd_rn(ARR(double, r_utf_d)[a], bd);
                    ARR(double, r_utf_d)[a] = u;
                    RES(float, utf, a) = __double2float_rn(u);
                } else {
                    RES(float, utf, a) = __fadd_rn(RES(float, utf, a), busy);
                }
            } else {
                busy = round6(__fsub_rn(now, RES(float, start, a)));
                RES(float, utf, a) = __fadd_rn(RES(float, utf, a), busy);
                if (DBR) ARR(uint8_t, r_flags)[a] &= ~F_UTF_PY;
            }
            RES(float, busy, a) = busy;
            PORD(uint8_t, done, po)++;
            PORD(uint8_t, loc, po) = LOC_NONE;
            if (!alone()) PUSH_AND_LEAVE(EV(OP_MACH_GOTPROC, a, 0))
            nev++;                                               // StoreGet[processing] popped right away
        }
        // fall through
        case OP_MACH_GOTPROC: {                                  // :452
            uint32_t po = RES(uint8_t, cur, a);
            list_append<kPoStride>(&RES(uint8_t, outh, a), &RES(uint8_t, outt, a), PO_NEXT, po);
            PORD(uint8_t, loc, po) = LOC_OUT;
            if (!alone()) PUSH_AND_LEAVE(EV(OP_MACH_PUTOUT, a, 0))
            nev++;                                               // StorePut[output] popped right away
        }
        // fall through
        case OP_MACH_PUTOUT: {                                   // :452-461, then loop top :375-380
            // Nothing else pending at this timestamp: the transport chain (TRANS_GOT ...) and the machine's next
            // operation chain (MACH_GOT -> PUTPROC -> REQ) that start here touch disjoint state and emit their records
            // / draws / calendar entries in the same relative order whether interleaved event by event (SimPy) or run
            // one after the other, so the machine chain may run right here.  Not for DBR (its dispatch scans observe
            // order locations) nor with queue logging (both chains write `queue` records).
            const bool solo = !DBR && !LOGQ && alone();
            if (((sc->trans_wait >> a) & 1u) && RES(uint8_t, outh, a) != kNone) {
                uint32_t po = list_pop<kPoStride>(&RES(uint8_t, outh, a), &RES(uint8_t, outt, a), PO_NEXT);
                PORD(uint8_t, loc, po) = LOC_NONE;
                sc->trans_wait &= ~(1u << a);
                pushN(EV(OP_TRANS_GOT, a, po));
            }
            if (DBR && p.dbr_repair && a == p.dbr_ccr) {        // patch P3: stores.resource_finished[CCR].put(po)
                if (sc->fin_n < 5) sc->fin[sc->fin_n++] = RES(uint8_t, cur, a); else fail(MSIM_ST_FIFO_OVERFLOW);
                pushN(EV(OP_FIN_PUT, 0, 0));
            }
            if (warm) emit(ringR(RM_UTIL, a), now, RES(float, busy, a));
            nev++;                                               // Release event of the `with` block
            if (RES(float, utf, a) >= RES(float, tbf, a)) { breakdown_start(a); break; }   // breakdown first (:376-377)
            {
                const bool chain = solo && (attn & ~A_FLUSH) == 0u;
                uint32_t pick = mach_select(a, chain);
                if (pick == kNone) break;                        // pushed, queue empty (the machine waits) or DBR scan requested
                b = (int)pick;
                force = 1;
                nev++;                                           // FilterStoreGet[input] popped ahead of the transport chain
                goto L_MACH_GOT;
            }
        }
        case OP_TO_TTR:                                          // _breakdowns :277-289
            if (warm) emit(ringR(RM_BREAKDOWN, a), RES(float, start, a), round6(__fsub_rn(now, RES(float, start, a))));
            RES(float, tbf, a) = draw_scalar(3, TAB(DDist, t_tbf)[a]);
            RES(float, utf, a) = 0.0f;
            if (DBR) { ARR(double, r_utf_d)[a] = 0.0; ARR(uint8_t, r_flags)[a] |= F_UTF_PY; }   // the int 0
            mach_select(a);
            break;
        case OP_TRANS_GOT:                                       // _transportation :300-304
            RES(uint8_t, tcur, a) = (uint8_t)b;
            PORD(uint8_t, loc, b) = LOC_TRANS;
            if (!alone()) PUSH_AND_LEAVE(EV(OP_TRANS_PUT, a, 0))
            nev++;                                               // StorePut[transport] popped right away
            // fall through
        case OP_TRANS_PUT: {                                     // :306-310 / :329-335
            uint32_t po = RES(uint8_t, tcur, a);
            int prod = PORD(uint8_t, prod, po);
            int total = TABC(uint16_t, t_route_start)[prod + 1] - TABC(uint16_t, t_route_start)[prod];
            PORD(uint8_t, loc, po) = LOC_NONE;
            const bool last = PORD(uint8_t, done, po) == total;
            if (last || !alone()) PUSH_AND_LEAVE(EV(last ? OP_TRANS_GOT2_FINAL : OP_TRANS_GOT2_NEXT, a, 0))
            nev++;                                               // StoreGet[transport] popped right away
        }
        // fall through
        case OP_TRANS_GOT2_NEXT: {                               // :330-336
            uint32_t po = RES(uint8_t, tcur, a);
            int prod = PORD(uint8_t, prod, po);
            int nxt = TABC(uint8_t, t_step_res)[TABC(uint16_t, t_route_start)[prod] + PORD(uint8_t, done, po)];
            list_append<kPoStride>(&RES(uint8_t, inh, nxt), &RES(uint8_t, int, nxt), PO_NEXT, po);
            PORD(uint8_t, loc, po) = LOC_IN;
            if (!alone()) PUSH_AND_LEAVE(EV(OP_TRANS_PUTIN, a, nxt))
            b = nxt;
            nev++;                                               // StorePut[input of the next resource] popped right away
        }
        // fall through
        case OP_TRANS_PUTIN: {                                   // :336-358
            serve_mach(b);
            if (LOGQ) {
                float q = PORD(float, qty, RES(uint8_t, tcur, a));
                ARR(float, r_qq)[b] = __fadd_rn(ARR(float, r_qq)[b], q);
                if (warm) queue_log_add(b, q);
            }
            trans_get(a);
            break;
        }
        case OP_TRANS_GOT2_FINAL: {                              // :311
            uint32_t po = RES(uint8_t, tcur, a);
            int prod = PORD(uint8_t, prod, po);
            PRD(float, fg, prod) = __fadd_rn(PRD(float, fg, prod), PORD(float, qty, po));
            if (!alone()) PUSH_AND_LEAVE(EV(OP_TRANS_FGPUT, a, 0))
            nev++;                                               // ContainerPut[finished goods] popped right away
        }
        // fall through
        case OP_TRANS_FGPUT: {                                   // ContainerPut[FG] processed, then :312
            uint32_t po = RES(uint8_t, tcur, a);
            int prod = PORD(uint8_t, prod, po);
            serve_fg(prod);
            float q = PORD(float, qty, po);
            float w = PRD(float, wip, prod);
            if (!(w >= q)) { fail(MSIM_ST_WIP_GET_BLOCKED); break; }
            PRD(float, wip, prod) = __fsub_rn(w, q);
            if (!alone()) PUSH_AND_LEAVE(EV(OP_TRANS_WIPGET, a, 0))
            nev++;                                               // ContainerGet[wip] popped right away
        }
        // fall through
        case OP_TRANS_WIPGET: {                                  // :314-326
            uint32_t po = RES(uint8_t, tcur, a);
            int prod = PORD(uint8_t, prod, po);
            sc->tot_wip = __fsub_rn(sc->tot_wip, PORD(float, qty, po));
            float f = PRD(float, fg, prod);
            float old = PRD(float, fgc, prod);
            PRD(float, fgc, prod) = f;
            sc->fgc_sum = __fadd_rn(sc->fgc_sum, __fsub_rn(f, old));
            if (warm) {
                float flow = __fsub_rn(now, PORD(float, rel, po));
                if (DBR && now_py && ARR(uint8_t, po_relpy)[po])
                    flow = __double2float_rn(__dsub_rn(now_d, (double)PORD(float, rel, po)));
                emit(ringP(PM_FLOW, prod), now, flow);
                inventory(prod, true, true);
            }
            free_po(po);
            trans_get(a);
            break;
        }
        case OP_DELIV_GOT:                                       // _delivery_orders :469-480
            pushU(EV(OP_INIT_SHIP, a, b));
            deliv_get(a);
            break;
        case OP_INIT_SHIP:                                       // :482-549 (onDue with patch P1)
            // (one finished_goods.get() site for the three delivery modes and the onDue timer: the handler switch is
            //  instruction-cache bound, inlined copies cost more than the branches)
            if (instantly()) {
                if (!(PRD(float, fg, a) >= DORD(float, qty, b))) {      // lost sale: nothing is taken
                    if (warm) {
                        inventory(a, false, true);
                        delivery_records(b, a, false);
                    }
                    nev++;                                       // process-end event
                    free_do(b);
                    break;
                }
            } else if (on_due()) {
                float wait = __fsub_rn(DORD(float, due, b), now);
                if (wait > 0.0f) {
                    float t = __fadd_rn(now, wait);
                    if (t == now) {
                        pushN_raw(EV(OP_TO_SHIP, a, b));          // float32-typed Timeout (due - now)
                    } else {
                        ARR(float, do_tt)[b] = t;
                        ARR(uint32_t, do_eid)[b] = tseq++;
                    }
                    break;
                }
            }
            // fall through
        case OP_TO_SHIP:
            fg_get(b, a);
            break;
        case OP_SHIP_GOT:                                        // ContainerGet[FG] processed
            if (warm) {
                inventory(a, false, true);
                delivery_records(b, a, true);
            }
            nev++;                                               // process-end event
            free_do(b);
            break;
        case OP_TO_TICK:                                         // DBRSimulation.scheduler, toc_dbr/simulation.py:200-275
            if (DBR) rope_tick();
            break;
        case OP_INIT_DORDER:                                     // process_order :277-283
            if (DBR) {
                float sched = ARR(float, po_tt)[b];
                ARR(float, po_tt)[b] = CUDART_INF_F;
                if (sched > now) {
                    float t = __fadd_rn(now, __fsub_rn(sched, now));
                    if (t == now) {
                        pushN_raw(EV(OP_TO_DORDER, 0, b));        // float32-typed Timeout (schedule - now)
                    } else {
                        ARR(float, po_tt)[b] = t;
                        ARR(uint32_t, po_eid)[b] = tseq++;
                    }
                } else {
                    dorder_go(b);
                }
            }
            break;
        case OP_TO_DORDER:                                       // process_order :284-289
            if (DBR) dorder_go(b);
            break;
        case OP_SEED_PUT:                                        // _update_finished_goods :71-73, ContainerPut processed
            serve_fg(a);
            nev++;                                               // process-end event
            break;
        case OP_FIN_PUT:                                         // StorePut[resource_finished[CCR]] processed
            if (DBR && sc->drain_wait && sc->fin_n) {
                uint32_t po = sc->fin[0];
                for (int i = 1; i < sc->fin_n; ++i) sc->fin[i - 1] = sc->fin[i];
                sc->fin_n--;
                sc->drain_wait = 0;
                pushN(EV(OP_DRAIN_GOT, 0, po));
            }
            break;
        case OP_DRAIN_GOT:                                       // _update_constraint_buffer :173-190
            if (DBR) {
                int prod = PORD(uint8_t, prod, b);
                int step = TABC(uint16_t, t_route_start)[prod] + PORD(uint8_t, done, b) - 1;
                float mean_pt = __double2float_rn(TABC(DDist, t_step_dist)[step].raw0);
                float ccr_time = __fadd_rn(__fmul_rn(mean_pt, PORD(float, qty, b)), __double2float_rn(p.dbr_ccr_setup));
                sc->cb_level = __fsub_rn(sc->cb_level, ccr_time);
                if (sc->fin_n) {
                    uint32_t po = sc->fin[0];
                    for (int i = 1; i < sc->fin_n; ++i) sc->fin[i - 1] = sc->fin[i];
                    sc->fin_n--;
                    pushN(EV(OP_DRAIN_GOT, 0, po));
                } else {
                    sc->drain_wait = 1;
                }
            }
            break;
        case OP_HOP:                                             // bookkeeping only, not an Environment.step()
            nev--;
            pushN(b > 1 ? EV(OP_HOP, a, b - 1) : EV(OP_TO_SETUP, a, 0));
            break;
        default:
            break;
        }
        return;
    L_TAIL_PUSH:
        pushN(tail_ev);
    }

    // process_order continuation: constraint_buffer_level += ccr_add; env.process(_release_order(po))
    __device__ __forceinline__ void dorder_go(int po) {
        sc->cb_level = __fadd_rn(sc->cb_level, PORD(float, ccr, po));
        pushU(EV(OP_INIT_RELEASE, 0, po));
        nev++;                                                   // process-end event
    }

    // one rope tick (toc_dbr/simulation.py:206-275); runs at python-int instants 0, interval, 2*interval, ...
    __device__ __forceinline__ void rope_tick() {
        const int P = p.lay.P;
        float rpr[32], qty[32];
        uint8_t ord[32];
        for (int i = 0; i < P; ++i) {
            float target = __double2float_rn(TAB(double, t_shipbuf)[i]);
            float have = __fadd_rn(PRD(float, wip, i), PRD(float, fg, i));
            float repl = __fsub_rn(target, have);
            if (!(repl > 0.0f)) repl = 0.0f;
            float lot = __double2float_rn(p.dbr_min_lot);
            qty[i] = repl >= lot ? repl : lot;                  // max(replenishment, min_lotsize)
            float key = round3(__fdiv_rn(repl, target));
            int j = i;                                           // stable insertion, descending release priority
            while (j > 0 && rpr[j - 1] < key) { rpr[j] = rpr[j - 1]; ord[j] = ord[j - 1]; --j; }
            rpr[j] = key; ord[j] = (uint8_t)i;
        }
        float room = p.dbr_ccr_limit != 0.0 ? __double2float_rn(p.dbr_interval * p.dbr_ccr_limit)
                                            : __fsub_rn(__double2float_rn(p.dbr_cb_target), sc->cb_level);
        if (room > 0.0f) {
            double n_rel = 0.0;
            for (int k = 0; k < P; ++k) {
                int i = ord[k];
                double cpt = TAB(double, t_ccrpt)[i];
                float ccr_t = 0.0f, sched = now;
                bool go = false;
                if (cpt > 0.0) {
                    if (room > 0.0f && n_rel < p.dbr_release_limit) {
                        ccr_t = __fadd_rn(__fmul_rn(qty[i], __double2float_rn(cpt)), __double2float_rn(p.dbr_ccr_setup));
                        sched = __fadd_rn(now, ccr_t);
                        go = true;
                        n_rel += 1.0;
                    }
                } else {
                    go = true;
                }
                if (qty[i] > 0.0f && go) {
                    uint32_t po = alloc_po();
                    if (status) return;
                    PORD(uint8_t, prod, po) = (uint8_t)i;
                    PORD(float, qty, po) = qty[i];
                    ARR(float, po_tt)[po] = sched;               // holds `schedule` until Initialize[process_order]
                    PORD(float, ccr, po) = ccr_t;
                    pushU(EV(OP_INIT_DORDER, 0, po));
                    room = __fsub_rn(room, ccr_t);
                }
            }
        }
        double nt = __dadd_rn(now_d, p.dbr_interval);            // yield env.timeout(scheduler_interval)
        sc->tick_d = nt;
        ARRC(float, tm_time)[p.lay.R + p.lay.P + 1] = __double2float_rn(nt);
        ARRC(uint32_t, tm_eid)[p.lay.R + p.lay.P + 1] = tseq++;
        ARRC(uint32_t, tm_ev)[p.lay.R + p.lay.P + 1] = EV(OP_TO_TICK, 0, 0);
    }

    // lane 0: run zero-delay events until the warp has to cooperate again.  Order of service at one timestamp
    // (SURVEY Appendix A.8): URGENT fifo -> calendar entries landing on `now` -> NORMAL fifo.
    // `first` = the calendar event the warp just selected (0 = none): it runs before anything queued, because the
    // URGENT fifo is empty whenever the calendar is consulted.
    // (two shapes of the same loop, chosen per kernel family at compile time: LOOP2 in sim_kernel)
    __device__ __forceinline__ int drain1(uint32_t first) {
        for (;;) {
            uint32_t ev;
            if (first) {                 // (single dispatch site: the handler switch must not be inlined twice)
                ev = first;
                first = 0;
            } else if (attn) {
                if (attn & A_STATUS) return REQ_DONE;
                if (attn & A_SELECT) { attn &= ~A_SELECT; return REQ_SELECT; }
                if (attn & A_FLUSH) { attn &= ~A_FLUSH; return REQ_FLUSH; }
                if (attn & A_PENDING) {
                    ev = pending;
                    attn &= ~A_PENDING;
                } else if (attn & A_URGENT) {
                    ev = ARRC(uint32_t, uq)[uqh & (uint32_t)(LCUQ - 1)];
                    uqh++;
                    if (uqh == uqt) attn &= ~A_URGENT;
                } else {
                    return REQ_ADVANCE;      // A_SAME: calendar entries at `now` precede NORMAL events created at `now`
                }
            } else {
                if (nqh == nqt) return REQ_ADVANCE;
                ev = ARRC(uint32_t, nq)[nqh & (uint32_t)(LCN(NQ) - 1)];
                nqh++;
            }
            dispatch(ev);
            if (nqt - nqh > (uint32_t)LCN(NQ)) fail(MSIM_ST_FIFO_OVERFLOW);
            if constexpr (EMIT2) { if (nstage >= (uint32_t)kStageFlushAt) return REQ_FLUSH; }
        }
    }

    __device__ __forceinline__ int drain2() {
        for (;;) {
            uint32_t ev;
            if (attn) {
                if (attn & A_STATUS) return REQ_DONE;
                if (attn & A_SELECT) { attn &= ~A_SELECT; return REQ_SELECT; }
                if (attn & A_FLUSH) { attn &= ~A_FLUSH; return REQ_FLUSH; }
                if (attn & A_PENDING) {      // the calendar entry the warp just selected (or Initialize[scheduler] at boot)
                    ev = pending;
                    attn &= ~A_PENDING;
                } else if (attn & A_URGENT) {
                    ev = ARRC(uint32_t, uq)[uqh & (uint32_t)(LCUQ - 1)];
                    uqh++;
                    if (uqh == uqt) attn &= ~A_URGENT;
                } else {
                    return REQ_ADVANCE;      // A_SAME: calendar entries at `now` precede NORMAL events created at `now`
                }
            } else {
                // one unsigned compare covers both ends: empty fifo (d - 1 wraps) and overrun (d > NQ)
                const uint32_t d = nqt - nqh;
                if (d - 1u >= (uint32_t)LCN(NQ)) {
                    if (d == 0u) return REQ_ADVANCE;
                    fail(MSIM_ST_FIFO_OVERFLOW);
                    return REQ_DONE;
                }
                ev = ARRC(uint32_t, nq)[nqh & (uint32_t)(LCN(NQ) - 1)];
                nqh++;
            }
            dispatch(ev);       // (single dispatch site: the handler switch must not be inlined twice)
            if constexpr (EMIT2) { if (nstage >= (uint32_t)kStageFlushAt) return REQ_FLUSH; }
        }
    }

    // t = 0: construction order of FactorySimulation._initiate_environment (factory_sim.py:61-89) -- the Initialize
    // events are URGENT and already ordered by eid, so they are executed directly in that order.
    __device__ __forceinline__ void boot() {
        const int R = p.lay.R, P = p.lay.P;
        now = 0.0f; warm = 0; nqh = nqt = uqh = uqt = 0; nstage = 0; logc = 0; nev = 0; tseq = 0; same_left = 0;
        status = 0; pending = 0; attn = 0; force = 0; rb_need = (uint32_t)p.rng_streams & 6u;
        now_d = 0.0; now_py = 1; sel_req = 0;
        for (int r = 0; r < R; ++r) {                            // _start_resources :248-261 (draws at construction)
            const DDist& tb_ = TAB(DDist, t_tbf)[r];
            RES(float, tbf, r) = tb_.kind == MSIM_DIST_NONE ? CUDART_INF_F : draw_scalar(3, tb_);
        }
        nev++;                                                   // Initialize[warmup]
        timer(R + P, p.warmup, EV(OP_TO_WARMUP, 0, 0), false);
        for (int r = 0; r < R; ++r) {
            nev += 2;                                            // Initialize[_transportation], Initialize[_production_system]
            sc->trans_wait |= 1u << r;
            mach_loop_top(r);
        }
        for (int prod = 0; prod < P; ++prod) {
            nev += 2;                                            // Initialize[_generate_demand_orders], Initialize[_delivery_orders]
            demand_draw(prod);
            sc->deliv_wait |= 1u << prod;
        }
        if (DBR) {                                               // _start_custom_process, toc_dbr/simulation.py:37-47
            for (int r = 0; r < R; ++r) ARR(uint8_t, r_flags)[r] = F_UTF_PY;
            nev += 2;                                            // Initialize[_update_constraint_buffer], [_process_demandOrders]
            sc->drain_wait = 1;
            sc->sched_wait = 1;                                  // the forwarder is the getter of inbound_demand_orders
            for (int prod = 0; prod < P; ++prod) {               // Initialize[_update_finished_goods(p)]: FG.put(shipping_buffer)
                nev++;
                float sbuf = __double2float_rn(TAB(double, t_shipbuf)[prod]);
                if (!(sbuf > 0.0f)) { fail(MSIM_ST_AMOUNT_NONPOSITIVE); return; }
                PRD(float, fg, prod) = sbuf;
                pushN(EV(OP_SEED_PUT, prod, 0));
            }
            pending = EV(OP_TO_TICK, 0, 0) | EV_PY; attn |= A_PENDING;   // Initialize[scheduler] (the rope): first tick at t = 0,
            return;                                              // counted and executed by dispatch before any NORMAL event
        }
        nev++;                                                   // Initialize[scheduler]
        sc->sched_wait = 1;
    }
};

// DBR only (python-typed clock): the order of calendar entries that share one float32 timestamp.
// SimPy's heap compares the time OBJECTS pairwise: two python floats compare exactly (float64); as soon as one side is a
// numpy float32 the python float is rounded to float32 first (NEP 50), the times tie and (priority, eid) decides
// (examples/toc_dbr/simulation.py:200-289 on simpy core's heap of (time, priority, eid, event) tuples).  `slot` is the
// entry with the smallest eid among those tied on the float32 bits `tbits`.  If that entry is python-typed, a python-
// typed entry of the same float32 cell with a smaller float64 time precedes it; the entry chosen that way still has to
// wait for the python-typed zero-delay events of an EARLIER float64 instant of the cell (`head_py`: the NORMAL fifo
// is not empty and its head is python-typed): then -1 is returned and the caller runs that event first.
// Runs on lane 0 only, a few times per 10^5 events (several calendar entries inside one float32 spacing of the clock).
__device__ __noinline__ int dbr_cell_order(const KParams& p, const unsigned char* sb, const Scal* sc, uint32_t tbits,
                                           int slot, double now_d, bool head_py) {
    const Lay& L = p.lay;
    const int R = L.R, P = L.P;
    if (slot >= L.NT) return slot;                               // order timers are float32-typed Timeouts
    const uint32_t* tt = reinterpret_cast<const uint32_t*>(sb + L.tm_time);
    const uint32_t* te = reinterpret_cast<const uint32_t*>(sb + L.tm_eid);
    const uint8_t* fl = sb + L.r_flags;
    const double* tpy = reinterpret_cast<const double*>(sb + L.r_tpy);
    int best = -1;
    double bd = 0.0;
    uint32_t be = 0u;
    for (int i = 0; i < R + P + 2; ++i) {
        if (tt[i] != tbits) continue;
        double d;
        if (i < R) { if (!(fl[i] & F_TIMER_PY)) continue; d = tpy[i]; }
        else if (i == R + P) d = p.warmup_d;
        else if (i == R + P + 1) d = sc->tick_d;
        else continue;                                           // demand generators: float32-typed gaps
        if (best < 0 || d < bd || (d == bd && te[i] < be)) { best = i; bd = d; be = te[i]; }
    }
    if (best < 0) return slot;                                   // no python-typed entry in this cell
    {   // is the smallest-eid entry itself python-typed?  (otherwise float32 ties decide by eid: keep it)
        bool slot_py = slot < R ? (fl[slot] & F_TIMER_PY) != 0 : (slot == R + P || slot == R + P + 1);
        if (!slot_py) return slot;
    }
    if (head_py && bd > now_d) return -1;                        // a later float64 instant: the pending events come first
    return best;
}

// cooperative log flush: one 128-bit record per lane (st.global.v4, 512 contiguous bytes per warp) into the
// replication's HBM ring + fused per-ring reductions.  Lanes holding records of the same ring are grouped with
// match.any; the group leader adds the group's values in lane (= emission) order, so the float64 sums are
// deterministic and independent of how replications are sharded.
template <bool FILLSEQ>
__device__ __forceinline__ void warp_flush(const KParams& p, const unsigned char* sb, int stage_off, long long rep, int lane,
                                           uint32_t n, uint32_t base) {
    const Lay& L = p.lay;
    uint4 rec = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
    if ((uint32_t)lane < n) rec = reinterpret_cast<const uint4*>(sb + stage_off)[lane];
    if (FILLSEQ) rec.w = base + (uint32_t)lane;
    if ((uint32_t)lane < n && p.a.log != nullptr && rep < p.a.n_log_reps) {
        unsigned long long at = ((unsigned long long)base + (unsigned)lane) % (unsigned long long)p.a.log_cap;
        reinterpret_cast<uint4*>(p.a.log)[rep * p.a.log_cap + (long long)at] = rec;
    }
    unsigned grp = __match_any_sync(FULLMASK, rec.z);
    bool leader = ((uint32_t)lane < n) && (lane == __ffs(grp) - 1);
    double tot = 0.0;
    float v = __uint_as_float(rec.y);
    for (uint32_t j = 0; j < n; ++j) {
        float vj = __shfl_sync(FULLMASK, v, (int)j);
        if ((grp >> j) & 1u) tot += (double)vj;
    }
    if (leader) {
        p.a.acc_sum[rep * L.K + rec.z] += tot;
        p.a.acc_cnt[rep * L.K + rec.z] += (uint32_t)__popc(grp);
    }
}

// ---------------------------------------------------------------------------------------------------------
constexpr int shape_threads(int shape) { return shape == 1 ? 192 : 256; }
constexpr int shape_ctas(int shape) { return shape == 1 ? 6 : (shape == 2 ? 5 : 4); }

template <int POLICY, int RNG, int OPTS, int SHAPE, int MODE>
__global__ void __launch_bounds__(shape_threads(SHAPE), shape_ctas(SHAPE)) sim_kernel(const __grid_constant__ KParams p) {
    MSIM_DYN_SMEM(smem);
    constexpr bool CLAY = OPTS == 0;                         // lean kernels: constant layout (msim_device.cuh `CL`)
    const Lay& L = p.lay;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char* sb = smem + (size_t)warp * L.slab_bytes;
    Scal* sc = reinterpret_cast<Scal*>(sb + LC(scal));
    Sim<POLICY, RNG, OPTS, MODE> s(p, sb, sc);
    const bool on_due = s.on_due();
    constexpr bool DBR = POLICY == MSIM_POLICY_DBR;
    constexpr bool LOOP2 = Tune<POLICY>::loop2;             // loop shape per kernel family (see the event loop below)
    constexpr int kPoStride = aos_po_stride(POLICY), kDoStride = kAosDoLinkStride;
    const uint32_t until_bits = __float_as_uint(p.run_until);

    for (;;) {
        long long rep = 0;
        if (lane == 0) rep = (long long)atomicAdd(p.work_counter, 1ull);
        rep = __shfl_sync(FULLMASK, rep, 0);
        if (rep >= p.a.n_reps) break;
        s.rep = rep;

        // ---- cooperative slab initialisation
        #pragma unroll 1
        for (int i = lane; i < L.slab_bytes / 4; i += 32) reinterpret_cast<uint32_t*>(sb)[i] = 0u;
        __syncwarp();
        #pragma unroll 1
        for (int i = lane; i < LCN(NT); i += 32) reinterpret_cast<float*>(sb + LC(tm_time))[i] = CUDART_INF_F;
        #pragma unroll 1
        for (int i = lane; i < L.MD; i += 32) {
            if (on_due) reinterpret_cast<float*>(sb + L.do_tt)[i] = CUDART_INF_F;
            DORD(uint8_t, next, i) = (uint8_t)(i + 1 < L.MD ? i + 1 : kNone);
        }
        #pragma unroll 1
        for (int i = lane; i < L.MP; i += 32) {
            PORD(uint8_t, next, i) = (uint8_t)(i + 1 < L.MP ? i + 1 : kNone);
            if (DBR) reinterpret_cast<float*>(sb + L.po_tt)[i] = CUDART_INF_F;
        }
        #pragma unroll 1
        for (int i = lane; i < L.R; i += 32) {
            RES(uint8_t, inh, i) = kNone; RES(uint8_t, int, i) = kNone; RES(uint8_t, outh, i) = kNone; RES(uint8_t, outt, i) = kNone;
            if (LOGQ) reinterpret_cast<float*>(sb + L.r_lastq)[i] = CUDART_NAN_F;
        }
        #pragma unroll 1
        for (int i = lane; i < L.P; i += 32) {
            PRD(uint8_t, obh, i) = kNone; PRD(uint8_t, obt, i) = kNone; PRD(uint8_t, fgqh, i) = kNone; PRD(uint8_t, fgqt, i) = kNone;
        }
        #pragma unroll 1
        for (int i = lane; i < L.K; i += 32) {
            p.a.acc_sum[rep * L.K + i] = 0.0;
            p.a.acc_cnt[rep * L.K + i] = 0u;
        }
        if (lane == 0) {
            sc->inb_head = kNone; sc->inb_tail = kNone; sc->po_free = 0; sc->do_free = 0;
            sc->rb_base[0] = sc->rb_base[1] = sc->rb_base[2] = 0u - (uint32_t)kXuBuf;   // empty: index 0 is 16 past the base
            if (RNG == MSIM_RNG_REPLAY) {
                long long oa = p.a.tape_A_off[rep], ob = p.a.tape_B_off[rep], oc = p.a.tape_C_off[rep], od = p.a.tape_D_off[rep];
                sc->tbase[0] = p.a.tape_A + oa; sc->tlen[0] = (uint32_t)(p.a.tape_A_off[rep + 1] - oa);
                sc->tbase[1] = p.a.tape_B + ob; sc->tlen[1] = (uint32_t)(p.a.tape_B_off[rep + 1] - ob);
                sc->tbase[2] = p.a.tape_C + oc; sc->tlen[2] = (uint32_t)(p.a.tape_C_off[rep + 1] - oc);
                sc->tbase[3] = p.a.tape_D + od; sc->tlen[3] = (uint32_t)(p.a.tape_D_off[rep + 1] - od);
            } else {
                uint64_t key = p.a.rep_seeds ? p.a.rep_seeds[rep] : rep_key(p.a.philox_seed, (uint64_t)(p.a.rep_offset + rep));
                sc->k0 = (uint32_t)key; sc->k1 = (uint32_t)(key >> 32);
            }
        }
        __syncwarp();
        if (lane == 0) s.boot();
        __syncwarp();

        // ---- event loop.  Two shapes of the loop head, same semantics; which one a kernel family gets was decided on the
        // B200 (profiles/r02_ab_variants_*.jsonl -- the kernel is instruction-fetch bound and responds to code layout more
        // than to instruction count): LOOP2 hands the selected calendar entry to lane 0 through the attention word and has
        // ONE flush site (`finishing`, warp-uniform: the replication is over, one more pass writes out what is staged);
        // the other shape passes it as an argument and flushes again after the loop.
        uint32_t first_ev = 0;
        bool finishing = false;
        for (;;) {
            int req;
            if constexpr (LOOP2) {
                req = REQ_FLUSH;
                if (!finishing) {
                    req = REQ_ADVANCE;
                    if (lane == 0) req = s.drain2();
                    req = __shfl_sync(FULLMASK, req, 0);
                    if (req == REQ_DONE) { finishing = true; continue; }
                }
                if (req == REQ_FLUSH) {
                    const uint32_t n = __shfl_sync(FULLMASK, s.nstage, 0);
                    if (n > 0) {
                        warp_flush<Tune<POLICY>::emit2>(p, sb, LC(stage), rep, lane, n, __shfl_sync(FULLMASK, s.logc, 0));
                        if (lane == 0) { s.logc += n; s.nstage = 0; }
                        __syncwarp();
                    }
                    if (finishing) break;
                    continue;
                }
            } else {
                req = REQ_ADVANCE;
                if (lane == 0) { req = s.drain1(first_ev); first_ev = 0; }
                req = __shfl_sync(FULLMASK, req, 0);
                uint32_t n = __shfl_sync(FULLMASK, s.nstage, 0);
                if (req == REQ_FLUSH) {
                    warp_flush<Tune<POLICY>::emit2>(p, sb, LC(stage), rep, lane, n, __shfl_sync(FULLMASK, s.logc, 0));
                    if (lane == 0) { s.logc += n; s.nstage = 0; }
                    __syncwarp();
                    continue;
                }
                if (req == REQ_DONE) break;
            }
            if (DBR && req == REQ_SELECT) {
                // ---- cooperative DBR dispatch (toc_dbr/simulation.py:291-328): for every queued order, the
                // quantity of same-product orders released strictly earlier that sit in ANY input/output/transport/
                // processing store, plus the FG level, over the shipping buffer; first minimum wins.
                const int r = __shfl_sync(FULLMASK, (int)s.sel_req, 0) - 1;
                const double* sbuf = reinterpret_cast<const double*>(p.tblob + L.t_shipbuf);
                uint32_t best = kNone, best_prev = kNone, prev = kNone;
                float best_pr = 0.0f;
#pragma unroll 1
                for (uint32_t o = RES(uint8_t, inh, r); o != kNone; prev = o, o = PORD(uint8_t, next, o)) {
                    const int op = PORD(uint8_t, prod, o);
                    const float ro = PORD(float, rel, o);
                    float part = 0.0f;
#pragma unroll 1
                    for (int i = lane; i < L.MP; i += 32)
                        if (PORD(uint8_t, loc, i) != LOC_NONE && PORD(uint8_t, prod, i) == op && PORD(float, rel, i) < ro)
                            part += PORD(float, qty, i);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(FULLMASK, part, off);
                    const float pr = __fdiv_rn(__fadd_rn(part, PRD(float, fg, op)), __double2float_rn(sbuf[op]));
                    if (best == kNone || pr < best_pr) { best = o; best_pr = pr; best_prev = prev; }
                }
                if (lane == 0) {
                    const uint8_t after = PORD(uint8_t, next, best);
                    if (best_prev == kNone) RES(uint8_t, inh, r) = after; else PORD(uint8_t, next, best_prev) = after;
                    if (after == kNone) RES(uint8_t, int, r) = (uint8_t)best_prev;
                    PORD(uint8_t, loc, best) = LOC_NONE;
                    s.pushN(EV(OP_MACH_GOT, (uint32_t)r, best));
                    s.sel_req = 0;
                }
                __syncwarp();
                continue;
            }

            // ---- cooperative next-event selection: SimPy's heap order (time, eid) over every calendar slot.
            // Each lane finds the minimum of its own slots, then three redux.sync reductions give the earliest
            // time, the smallest eid at that time and the number of entries tied on it.
            uint32_t bt = 0xFFFFFFFFu, be = 0xFFFFFFFFu;
            int bslot = -1, leq = 0;
            {
                const uint32_t* tt = reinterpret_cast<const uint32_t*>(sb + LC(tm_time));
                const uint32_t* te = reinterpret_cast<const uint32_t*>(sb + LC(tm_eid));
                if (LCN(NT) == 32) {          // the common shape (R + P + 2 <= 32): one calendar slot per lane
                    bt = tt[lane]; be = te[lane]; bslot = lane; leq = 1;
                } else {
#pragma unroll 1
                    for (int i = lane; i < LCN(NT); i += 32) {
                        const uint32_t t = tt[i], e = te[i];
                        if (t < bt) { bt = t; be = e; bslot = i; leq = 1; }
                        else if (t == bt) { leq++; if (e < be) { be = e; bslot = i; } }
                    }
                }
                if (on_due) {
                    const uint32_t* dt = reinterpret_cast<const uint32_t*>(sb + L.do_tt);
                    const uint32_t* de = reinterpret_cast<const uint32_t*>(sb + L.do_eid);
#pragma unroll 1
                    for (int i = lane; i < L.MD; i += 32) {
                        const uint32_t t = dt[i], e = de[i];
                        if (t < bt) { bt = t; be = e; bslot = LCN(NT) + i; leq = 1; }
                        else if (t == bt) { leq++; if (e < be) { be = e; bslot = LCN(NT) + i; } }
                    }
                }
                if (DBR) {
                    const uint32_t* pt = reinterpret_cast<const uint32_t*>(sb + L.po_tt);
                    const uint32_t* pe = reinterpret_cast<const uint32_t*>(sb + L.po_eid);
#pragma unroll 1
                    for (int i = lane; i < L.MP; i += 32) {
                        const uint32_t t = pt[i], e = pe[i];
                        if (t < bt) { bt = t; be = e; bslot = LCN(NT) + L.MD + i; leq = 1; }
                        else if (t == bt) { leq++; if (e < be) { be = e; bslot = LCN(NT) + L.MD + i; } }
                    }
                }
            }
            if (RNG == MSIM_RNG_PHILOX) {
                // ---- opportunistic bulk refill of the precomputed (x,u) pairs: the warp is converged here anyway
                uint32_t need = __shfl_sync(FULLMASK, s.rb_need, 0);
                if (need) {
#pragma unroll 1
                    for (int st = 1; st < 3; ++st) {
                        if (!((need >> st) & 1u)) continue;
                        const uint32_t next = sc->draw[st];
                        __syncwarp();
                        if (lane < kXuBuf) {
                            XU r = raw_xu(sc->k0, sc->k1, (uint32_t)st, next + (uint32_t)lane, 0u);
                            reinterpret_cast<float2*>(sb + LC(rb))[(st - 1) * kXuBuf + lane] = make_float2(r.x, r.u);
                        }
                        __syncwarp();
                        if (lane == 0) sc->rb_base[st] = next;
                    }
                    if (lane == 0) s.rb_need = 0;
                    __syncwarp();
                }
            }
            const uint32_t tbits = __reduce_min_sync(FULLMASK, bt);
            if (tbits >= until_bits) {           // the URGENT stop event at run_until precedes everything at that time
                if constexpr (LOOP2) { finishing = true; continue; } else break;
            }
            const bool mine = bt == tbits;
            const uint32_t emin = __reduce_min_sync(FULLMASK, mine ? be : 0xFFFFFFFFu);
            const uint32_t cnt = __reduce_add_sync(FULLMASK, mine ? (uint32_t)leq : 0u);
            const unsigned who = __ballot_sync(FULLMASK, mine && be == emin);
            const int slot0 = __shfl_sync(FULLMASK, bslot, __ffs(who) - 1);
            int slot = slot0;
#if !MSIM_X_F32CELL
            if (DBR && lane == 0 && (cnt > 1u || (s.attn & A_SAME))) {
                // several calendar entries inside one float32 cell of the clock: python-typed instants order exactly
                const bool head_py = s.nqh != s.nqt &&
                    (reinterpret_cast<const uint32_t*>(sb + LC(nq))[s.nqh & (uint32_t)(LCN(NQ) - 1)] & EV_PY) != 0u;
                slot = dbr_cell_order(p, sb, sc, tbits, slot0, s.now_d, head_py);
                if (slot < 0) {                              // the fifo head belongs to an earlier float64 instant
                    const uint32_t head_ev = reinterpret_cast<const uint32_t*>(sb + LC(nq))[s.nqh & (uint32_t)(LCN(NQ) - 1)];
                    if constexpr (LOOP2) { s.pending = head_ev; s.attn |= A_PENDING; } else first_ev = head_ev;
                    s.nqh++;
                }
            }
#endif
            if (lane == 0 && slot >= 0) {
                if (s.attn & A_SAME) {
                    if (--s.same_left == 0) s.attn &= ~A_SAME;
                } else {
                    s.now = __uint_as_float(tbits);
                    s.same_left = cnt - 1u;
                    if (cnt > 1u) s.attn |= A_SAME;
                }
                sc->n_timeouts++;
                uint32_t ev;
                if (DBR) s.now_py = 0;     // Environment._now takes the type of the popped entry (Appendix B.2)
                if (slot < LCN(NT)) {
                    ev = reinterpret_cast<const uint32_t*>(sb + LC(tm_ev))[slot];
                    reinterpret_cast<float*>(sb + LC(tm_time))[slot] = CUDART_INF_F;
                    if (DBR) {
                        if (slot < L.R) {
                            if ((sb + L.r_flags)[slot] & F_TIMER_PY) {
                                s.now_py = 1;
                                s.now_d = reinterpret_cast<const double*>(sb + L.r_tpy)[slot];
                            }
                        } else if (slot == L.R + L.P) {
                            s.now_py = 1; s.now_d = p.warmup_d;
                        } else if (slot == L.R + L.P + 1) {
                            s.now_py = 1; s.now_d = sc->tick_d;
                        }
                    }
                } else if (on_due && slot < LCN(NT) + L.MD) {
                    int d = slot - LCN(NT);
                    ev = EV(OP_TO_SHIP, DORD(uint8_t, prod, d), d);
                    reinterpret_cast<float*>(sb + L.do_tt)[d] = CUDART_INF_F;
                } else {
                    int po = slot - LCN(NT) - L.MD;
                    ev = EV(OP_TO_DORDER, 0, po);
                    reinterpret_cast<float*>(sb + L.po_tt)[po] = CUDART_INF_F;
                }
                // (it runs before anything queued: the URGENT fifo is empty whenever the calendar is consulted)
                if constexpr (LOOP2) { s.pending = DBR ? ev | (s.now_py << 24) : ev; s.attn |= A_PENDING; }
                else first_ev = DBR ? ev | (s.now_py << 24) : ev;
            }
            __syncwarp();
        }

        if constexpr (!LOOP2) {
            uint32_t n = __shfl_sync(FULLMASK, s.nstage, 0);
            if (n > 0) {
                warp_flush<Tune<POLICY>::emit2>(p, sb, LC(stage), rep, lane, n, __shfl_sync(FULLMASK, s.logc, 0));
                if (lane == 0) { s.logc += n; s.nstage = 0; }
                __syncwarp();
            }
        }
        // ---- per-replication outputs
        if (lane == 0) {
            p.a.n_events[rep] = (unsigned long long)s.nev + 1ull;     // + the stop event of run(until)
            p.a.status[rep] = s.status;
            if (p.a.n_timeouts) p.a.n_timeouts[rep] = sc->n_timeouts;
            if (p.a.log_count && rep < p.a.n_log_reps) p.a.log_count[rep] = s.logc;
            if (RNG == MSIM_RNG_PHILOX && p.a.rec_count) {
                for (int k = 0; k < 4; ++k) p.a.rec_count[rep * 4 + k] = sc->draw[k];
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
typedef void (*sim_fn)(const KParams);
template <int OPTS, int SHAPE, int MODE>
static sim_fn pick_rng(int policy, int rng) {
    if (policy == MSIM_POLICY_FIFO)
        return rng == MSIM_RNG_REPLAY ? sim_kernel<MSIM_POLICY_FIFO, MSIM_RNG_REPLAY, OPTS, SHAPE, MODE>
                                      : sim_kernel<MSIM_POLICY_FIFO, MSIM_RNG_PHILOX, OPTS, SHAPE, MODE>;
    if (policy == MSIM_POLICY_MAX_PRIORITY)
        return rng == MSIM_RNG_REPLAY ? sim_kernel<MSIM_POLICY_MAX_PRIORITY, MSIM_RNG_REPLAY, OPTS, SHAPE, MODE>
                                      : sim_kernel<MSIM_POLICY_MAX_PRIORITY, MSIM_RNG_PHILOX, OPTS, SHAPE, MODE>;
    if (policy == MSIM_POLICY_DBR)
        return rng == MSIM_RNG_REPLAY ? sim_kernel<MSIM_POLICY_DBR, MSIM_RNG_REPLAY, OPTS, SHAPE, MODE>
                                      : sim_kernel<MSIM_POLICY_DBR, MSIM_RNG_PHILOX, OPTS, SHAPE, MODE>;
    if (policy == MSIM_POLICY_EDD)
        return rng == MSIM_RNG_REPLAY ? sim_kernel<MSIM_POLICY_EDD, MSIM_RNG_REPLAY, OPTS, SHAPE, MODE>
                                      : sim_kernel<MSIM_POLICY_EDD, MSIM_RNG_PHILOX, OPTS, SHAPE, MODE>;
    return nullptr;
}
template <int SHAPE>
static sim_fn pick_lean(int policy, int rng, int delivery_mode) {
    if (delivery_mode == MSIM_DELIVERY_INSTANTLY && (policy == MSIM_POLICY_MAX_PRIORITY || policy == MSIM_POLICY_EDD)) {
        // Tune<POLICY>::mode3 families: `instantly` is its own kernel (MODE 3)
        if (policy == MSIM_POLICY_MAX_PRIORITY)
            return rng == MSIM_RNG_REPLAY ? sim_kernel<MSIM_POLICY_MAX_PRIORITY, MSIM_RNG_REPLAY, 0, SHAPE, 3>
                                          : sim_kernel<MSIM_POLICY_MAX_PRIORITY, MSIM_RNG_PHILOX, 0, SHAPE, 3>;
        return rng == MSIM_RNG_REPLAY ? sim_kernel<MSIM_POLICY_EDD, MSIM_RNG_REPLAY, 0, SHAPE, 3>
                                      : sim_kernel<MSIM_POLICY_EDD, MSIM_RNG_PHILOX, 0, SHAPE, 3>;
    }
    return delivery_mode == MSIM_DELIVERY_ON_DUE ? pick_rng<0, SHAPE, 1>(policy, rng) : pick_rng<0, SHAPE, 0>(policy, rng);
}
// lean: the run uses neither queue records nor tape recording (OPTS = 0) -> one kernel per delivery family; the general
// kernels (OPTS = 3) decide the delivery mode at run time and exist in the default shape only.  Shapes 0 and 2 are
// compiled with -DMSIM_ALL_SHAPES=1 (launch-shape sweeps); otherwise every request falls back to shape 1 -- except the
// onDue family, which also exists in shape 0: its per-order delivery timers make the slab so large that shared memory holds
// 30-32 warps per SM whatever the register budget, so it takes the 64 registers (+3..7 %, profiles/r02_ab_ondue_64regs_z.jsonl).
static sim_fn pick(int policy, int rng, bool lean, int shape, int delivery_mode) {
    if (!lean) return pick_rng<3, 1, 2>(policy, rng);
#if MSIM_ALL_SHAPES
    if (shape == 0) return pick_lean<0>(policy, rng, delivery_mode);
    if (shape == 2) return pick_lean<2>(policy, rng, delivery_mode);
#else
    if (shape == 0 && delivery_mode == MSIM_DELIVERY_ON_DUE) return pick_rng<0, 0, 1>(policy, rng);
#endif
    return pick_lean<1>(policy, rng, delivery_mode);
}

int sim_shape_compiled(int shape, int delivery_mode) {
    if (MSIM_ALL_SHAPES) return shape;
    return shape == 0 && delivery_mode == MSIM_DELIVERY_ON_DUE ? 0 : 1;
}
int sim_shape_max_warps(int shape) { return shape_threads(shape) / 32; }

int sim_kernel_attrs(int policy, int rng_mode, int shape, int delivery_mode, size_t smem, int* regs, int block, int* blocks_per_sm) {
    // residency of the LEAN kernel of this shape; a run that records tapes or logs queue lengths launches the general
    // kernel with the same grid (CTAs that do not fit simply queue)
    sim_fn f = pick(policy, rng_mode, true, shape, delivery_mode), g = pick(policy, rng_mode, false, shape, delivery_mode);
    if (!f || !g) return MSIM_E_UNSUPPORTED;
    cudaFuncAttributes at;
    if (cudaFuncGetAttributes(&at, f) != cudaSuccess) return MSIM_E_CUDA;
    if (regs) *regs = at.numRegs;
    if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MSIM_E_CUDA;
    if (blocks_per_sm) {
        int nb = 0, ng = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f, block, smem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ng, g, block, smem) != cudaSuccess) return MSIM_E_CUDA;
        *blocks_per_sm = nb;          // (a general-kernel launch with more CTAs than fit simply queues the rest)
        (void)ng;
    }
    return 0;
}

int launch_sim(const KParams& p, int policy, int rng_mode, int shape, bool const_layout, int grid, int block, size_t smem,
               void* stream) {
    // lean kernels: the run uses neither queue records nor tape recording AND the scenario has the standard shape whose
    // layout offsets they carry as immediates (msim_device.cuh `CL`).  MSIM_NO_LEAN=1 (diagnostic, tests): general kernel.
    const char* no_lean = getenv("MSIM_NO_LEAN");
    const bool lean = const_layout && !p.log_queues && p.a.rec_cap == 0 && !(no_lean && no_lean[0] == '1');
    sim_fn f = pick(policy, rng_mode, lean, shape, p.delivery_mode);
    if (!f) return MSIM_E_UNSUPPORTED;
    if (!lean && block > shape_threads(1)) {
        // the general kernels exist in the 192-thread shape only: same slabs, narrower CTAs, proportionally more of them
        // (a persistent grid pulling replications from a counter: CTAs that do not fit simply queue)
        const int wpb = block / 32, w1 = shape_threads(1) / 32;
        grid = (int)(((long long)grid * wpb + w1 - 1) / w1);
        smem = smem / (size_t)wpb * (size_t)w1;
        block = w1 * 32;
    }
    // The opt-in shared-memory limit is per FUNCTION, not per handle: handles with different slab sizes share a kernel
    // and msim_create lowers the attribute again while it probes CTA widths, so it is set before every launch (a host-
    // side attribute write, negligible next to the launch).
    if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MSIM_E_CUDA;
    MSIM_LAUNCH(f, grid, block, smem, (cudaStream_t)stream, p);
    return cudaGetLastError() == cudaSuccess ? 0 : MSIM_E_CUDA;
}

#undef ARR
#undef TAB
#undef RES
#undef PRD
#undef PORD
#undef DORD
#undef PO_NEXT
#undef DO_NEXT
#undef LOGQ
#undef RECORDING
#undef NOT_RECORDING
#undef PUSH_AND_LEAVE

} // namespace msim
