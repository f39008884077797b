// emu_runtime.cpp -- TEST INFRASTRUCTURE, not product code (see include/cuda_runtime.h).
//
// Fiber scheduler behind the SIMT shim: one fiber per CUDA thread of the block a host thread is running.
// A fiber runs until it reaches a warp collective or a block barrier; the last participant to arrive publishes the
// 32 contributions and releases the others.  That is one legal schedule of the CUDA memory/execution model
// (independent thread scheduling: only the *_sync primitives order lanes), chosen to be deterministic.
// A pass of the scheduler that finds unfinished fibers but none runnable is a divergent collective (a hang on the
// GPU): the launch is aborted and reported through cudaGetLastError().
#include <cuda_runtime.h>
#include <sys/mman.h>

#if !defined(__x86_64__)
#error "the fiber switch below is written for x86-64 (the container and the GPU boxes); port emu_switch to build elsewhere"
#endif

// Minimal cooperative context switch (callee-saved registers + stack pointer); ucontext's swapcontext costs a
// sigprocmask system call per switch, and a warp collective is 32 switches.
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
    .text
    .globl emu_switch
    .type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size emu_switch,.-emu_switch
    .section .note.GNU-stack,"",@progbits
    .text
)");

#if defined(__SANITIZE_ADDRESS__)
extern "C" void __sanitizer_start_switch_fiber(void** fake_stack_save, const void* bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void* fake_stack_save, const void** bottom_old, size_t* size_old);
#define ASAN_START(save, bottom, size) __sanitizer_start_switch_fiber((save), (bottom), (size))
#define ASAN_FINISH(fake, bottom_old, size_old) __sanitizer_finish_switch_fiber((fake), (bottom_old), (size_old))
#else
#define ASAN_START(save, bottom, size) ((void)0)
#define ASAN_FINISH(fake, bottom_old, size_old) ((void)0)
#endif

#include <atomic>
#include <thread>

namespace emu {

static constexpr size_t kStack = 1u << 20;

struct Warp {
    uint64_t vals[32];
    uint64_t result[32];
    unsigned arrived = 0;
    unsigned alive = 0;
    int kind = 0;           // primitive the lanes of the pending collective arrived at
    unsigned gen = 0;
};

struct Block;
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    void* fake = nullptr;   // AddressSanitizer fake-stack handle
    Block* blk = nullptr;
    unsigned tid = 0;
    bool done = false;
    int wait_kind = 0;      // 0 runnable, 1 warp collective, 2 block barrier
    unsigned wait_gen = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    void* sched_sp = nullptr;
    const void* sched_bottom = nullptr;
    size_t sched_size = 0;
    long long bid = 0;
    int bdim = 0;
    unsigned char* smem = nullptr;
    int bar_arrived = 0, bar_alive = 0;
    bool mismatch = false;
    unsigned bar_gen = 0;
    const std::function<void()>* body = nullptr;
};

thread_local Fiber* g_cur = nullptr;
static thread_local Block* g_blk = nullptr;
static thread_local std::vector<char*>* g_stacks = nullptr;
static std::atomic<int> g_error{0};

emu_dim3 tid() { return emu_dim3{g_cur->tid, 0, 0}; }
emu_dim3 bid() { return emu_dim3{(unsigned)g_blk->bid, 0, 0}; }
emu_dim3 bdim() { return emu_dim3{(unsigned)g_blk->bdim, 1, 1}; }
unsigned char* dyn_smem() { return g_blk->smem; }
int last_error() { return g_error.exchange(0); }

static inline void yield_to_scheduler() {
    Fiber* f = g_cur;
    ASAN_START(f->done ? nullptr : &f->fake, f->blk->sched_bottom, f->blk->sched_size);
    emu_switch(&f->sp, f->blk->sched_sp);
    ASAN_FINISH(f->fake, nullptr, nullptr);
}

const uint64_t* warp_gather(unsigned mask, uint64_t v, int kind) {
    Fiber* f = g_cur;
    Warp& w = f->blk->warps[f->tid >> 5];
    const unsigned lane = f->tid & 31u;
    if (w.arrived == 0) {
        w.kind = kind;
    } else if (w.kind != kind && !f->blk->mismatch) {
        fprintf(stderr, "[msim emu] block %lld warp %u: lanes meet at different warp primitives (%d vs %d)\n",
                f->blk->bid, f->tid >> 5, w.kind, kind);
        f->blk->mismatch = true;
    }
    w.vals[lane] = v;
    w.arrived |= 1u << lane;
    const unsigned expected = mask & w.alive;
    if ((w.arrived & expected) == expected) {
        memcpy(w.result, w.vals, sizeof(w.vals));
        w.arrived = 0;
        w.gen++;
    } else {
        f->wait_kind = 1;
        f->wait_gen = w.gen;
        yield_to_scheduler();
    }
    return w.result;
}

void block_barrier() {
    Fiber* f = g_cur;
    Block* b = f->blk;
    if (++b->bar_arrived >= b->bar_alive) {
        b->bar_arrived = 0;
        b->bar_gen++;
    } else {
        f->wait_kind = 2;
        f->wait_gen = b->bar_gen;
        yield_to_scheduler();
    }
}

static void fiber_main() {
    Fiber* f = g_cur;
    ASAN_FINISH(nullptr, &f->blk->sched_bottom, &f->blk->sched_size);
    (*f->blk->body)();
    f->done = true;
    Block* b = f->blk;
    Warp& w = b->warps[f->tid >> 5];
    w.alive &= ~(1u << (f->tid & 31u));
    b->bar_alive--;
    // a lane that exits can complete a collective / barrier the others are waiting on
    const unsigned expected = 0xFFFFFFFFu & w.alive;
    if (w.arrived && expected && (w.arrived & expected) == expected) {
        memcpy(w.result, w.vals, sizeof(w.vals));
        w.arrived = 0;
        w.gen++;
    }
    if (b->bar_alive > 0 && b->bar_arrived >= b->bar_alive) {
        b->bar_arrived = 0;
        b->bar_gen++;
    }
    yield_to_scheduler();
    abort();   // a finished fiber is never resumed
}

static char* stack_for(int i) {
    if (!g_stacks) g_stacks = new std::vector<char*>();
    while ((int)g_stacks->size() <= i) {
        void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("mmap"); abort(); }
        g_stacks->push_back((char*)p);
    }
    return (*g_stacks)[i];
}

static bool run_block(long long bid, int block, size_t smem, const std::function<void()>& body) {
    Block b;
    b.bid = bid;
    b.bdim = block;
    b.body = &body;
    b.fibers.resize(block);
    b.warps.resize((block + 31) / 32);
    b.bar_alive = block;
    void* sm = nullptr;
    if (posix_memalign(&sm, 128, smem ? smem : 16)) abort();
    memset(sm, 0xA5, smem ? smem : 16);          // shared memory starts uninitialised on the GPU
    b.smem = (unsigned char*)sm;
    g_blk = &b;
    for (int t = 0; t < block; ++t) {
        Fiber& f = b.fibers[t];
        f.blk = &b;
        f.tid = (unsigned)t;
        b.warps[t >> 5].alive |= 1u << (t & 31);
        f.stack = stack_for(t);
        // initial frame: six callee-saved registers, then fiber_main as the return address of emu_switch; after
        // that `ret` the stack pointer is 8 (mod 16), as at any function entry
        void** top = reinterpret_cast<void**>(f.stack + kStack);
        top[-1] = nullptr;
        top[-2] = reinterpret_cast<void*>(&fiber_main);
        for (int k = 3; k <= 8; ++k) top[-k] = nullptr;
        f.sp = top - 8;
    }
    bool ok = true;
    for (;;) {
        int ran = 0, left = 0;
        for (int t = 0; t < block; ++t) {
            Fiber& f = b.fibers[t];
            if (f.done) continue;
            left++;
            if (f.wait_kind == 1 && b.warps[t >> 5].gen == f.wait_gen) continue;
            if (f.wait_kind == 2 && b.bar_gen == f.wait_gen) continue;
            f.wait_kind = 0;
            g_cur = &f;
            {
                void* fake = nullptr;
                ASAN_START(&fake, f.stack, kStack);
                emu_switch(&b.sched_sp, f.sp);
                ASAN_FINISH(fake, nullptr, nullptr);
            }
            ran++;
        }
        if (left == 0) break;
        if (ran == 0) {
            fprintf(stderr, "[msim emu] block %lld: %d threads wait on a collective that can never complete "
                            "(divergent *_sync / barrier)\n", bid, left);
            ok = false;
            break;
        }
    }
    g_cur = nullptr;
    g_blk = nullptr;
    free(sm);
    return ok && !b.mismatch;
}

void launch(long long grid, int block, size_t smem, const std::function<void()>& body) {
    if (grid <= 0 || block <= 0 || block > 1024 || smem > 227u * 1024u) { g_error = 1; return; }
    int nthr = 0;
    if (const char* s = getenv("MSIM_EMU_THREADS")) nthr = atoi(s);
    if (nthr <= 0) nthr = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
    if ((long long)nthr > grid) nthr = (int)grid;
    std::atomic<long long> next{0};
    std::atomic<int> bad{0};
    auto worker = [&]() {
        for (;;) {
            long long b = next.fetch_add(1);
            if (b >= grid || bad.load()) break;
            if (!run_block(b, block, smem, body)) bad = 1;
        }
        if (g_stacks) {                       // fiber stacks live as long as the host thread that used them
            for (char* p : *g_stacks) munmap(p, kStack);
            delete g_stacks;
            g_stacks = nullptr;
        }
    };
    if (nthr == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int i = 0; i < nthr; ++i) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    if (bad.load()) g_error = 1;
}

}  // namespace emu
