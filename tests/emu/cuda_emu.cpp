// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY: the runtime of the SIMT interpreter declared in cuda_emu.h (fibers, warp
// collectives, CTA barrier, guarded device allocations, mbarrier / bulk-copy model).  See the header for what it is for
// and why it is not a product path.
#include "cuda_emu.h"

#include <sys/mman.h>

#include <cstdarg>
#include <map>
#include <mutex>
#include <vector>

// ---- fibers: a minimal x86-64 context switch (callee-saved registers + stack pointer) ------------------------------------
extern "C" void emu_switch(void **save_sp, void *load_sp);
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
)");

namespace emu
{
namespace
{
constexpr size_t STACK_BYTES = 256 * 1024;

struct Warp
{
    unsigned alive = 0, arrived = 0, mask = 0, gen = 0;
    int kind = 0;
    uint64_t vals[32], res[32];
};

struct Fiber
{
    ThreadCtx ctx;
    void *sp = nullptr;
    void *stack = nullptr;
    bool done = false;
    unsigned lane = 0;
    Warp *warp = nullptr;
    // what the fiber waits for: *wait_gen still equal to wait_val means "not yet"
    const unsigned *wait_gen = nullptr;
    unsigned wait_val = 0;
};

struct Block
{
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    unsigned alive = 0, bar_arrived = 0, bar_gen = 0;
    void (*body)(void *) = nullptr;
    void *arg = nullptr;
    unsigned char *dyn_smem = nullptr;
};

Block *g_block = nullptr;
Fiber *g_cur = nullptr;
void *g_sched_sp = nullptr;
std::vector<void *> g_stack_pool;
ThreadCtx g_host_ctx{};

struct PendingCopy
{
    uint32_t dst, bytes, bar;
    const void *src;
};
std::vector<PendingCopy> g_pending;
bool g_tma_lazy = true;
// VRT_EMU_ORDER=reverse runs the threads of a CTA (and the lanes of a warp) from the highest index down.  After a collective
// the last lane to arrive runs on first (lane 31 forward, lane 0 reversed), so code that only works for one order -- shared
// memory handed between lanes without a __syncwarp -- reads stale data in the other; the tests run both
bool g_reverse = false;
uint32_t g_random = 0; // xorshift state, 0 = off

// guarded allocations: [guard | payload | guard]
#ifdef __SANITIZE_ADDRESS__
constexpr size_t GUARD = 0; // an AddressSanitizer build: its own red zones sit right at the block edges and also catch reads
#else
constexpr size_t GUARD = 256;
#endif
constexpr unsigned char GUARD_BYTE = 0xE7, FRESH_BYTE = 0xA5;
std::map<void *, size_t> g_allocs;

void yield_to_scheduler() { emu_switch(&g_cur->sp, g_sched_sp); }

void warp_try_complete(Warp &w)
{
    if (w.arrived != 0 && w.arrived == (w.mask & w.alive))
    {
        for (int i = 0; i < 32; ++i) w.res[i] = ((w.arrived >> i) & 1u) ? w.vals[i] : 0ull;
        w.arrived = 0;
        w.gen++;
    }
}
void block_try_complete(Block &b)
{
    if (b.bar_arrived != 0 && b.bar_arrived == b.alive)
    {
        b.bar_arrived = 0;
        b.bar_gen++;
    }
}

void fiber_main()
{
    Fiber *f = g_cur;
    g_block->body(g_block->arg);
    f->done = true;
    // leaving: collectives of the remaining lanes / threads no longer wait for this one
    f->warp->alive &= ~(1u << f->lane);
    g_block->alive--;
    warp_try_complete(*f->warp);
    block_try_complete(*g_block);
    yield_to_scheduler();
    fatal("a finished fiber was resumed");
}

void *stack_get()
{
    if (!g_stack_pool.empty())
    {
        void *s = g_stack_pool.back();
        g_stack_pool.pop_back();
        return s;
    }
    void *s = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (s == MAP_FAILED) fatal("mmap of a fiber stack failed");
    return s;
}

void fiber_init(Fiber &f)
{
    f.stack = stack_get();
    uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
    void **sp = (void **)top;
    *--sp = nullptr;             // fake return address of fiber_main (never used)
    *--sp = (void *)&fiber_main; // `ret` of emu_switch jumps here with rsp = 8 (mod 16), as after a call
    for (int i = 0; i < 6; ++i) *--sp = nullptr; // rbp rbx r12 r13 r14 r15
    f.sp = sp;
}

void check_guards(const char *when)
{
    for (const auto &a : g_allocs)
    {
        const unsigned char *base = (const unsigned char *)a.first - GUARD;
        for (size_t i = 0; i < GUARD; ++i)
            if (base[i] != GUARD_BYTE || base[GUARD + a.second + i] != GUARD_BYTE)
                fatal("write outside a device allocation of %zu bytes (%s the block, offset %zu), detected %s", a.second, base[i] != GUARD_BYTE ? "before" : "after",
                      i, when);
    }
}

bool runnable(const Fiber &f) { return !f.done && !(f.wait_gen && *f.wait_gen == f.wait_val); }

void run_block(Block &b)
{
    g_block = &b;
    for (;;)
    {
        bool progress = false;
        unsigned remaining = 0;
        for (size_t i = 0; i < b.fibers.size(); ++i)
        {
            Fiber &f = b.fibers[g_reverse ? b.fibers.size() - 1 - i : i];
            if (f.done) continue;
            remaining++;
            if (!runnable(f)) continue;
            // the fiber (and the lanes of its warp it hands over to, see wait_in_warp) runs until something yields back here
            g_cur = &f;
            emu_switch(&g_sched_sp, f.sp);
            g_cur = nullptr;
            progress = true;
        }
        for (Fiber &f : b.fibers)
            if (f.done && f.stack)
            {
                g_stack_pool.push_back(f.stack);
                f.stack = nullptr;
            }
        if (remaining == 0) break;
        if (!progress)
        {
            unsigned waiting_warp = 0, waiting_block = 0;
            for (Fiber &f : b.fibers)
                if (!f.done) (f.wait_gen == &b.bar_gen ? waiting_block : waiting_warp)++;
            fatal("deadlock in block (%u,%u): %u threads wait in a warp collective that the other lanes never reach, %u in __syncthreads", b.fibers[0].ctx.bid.x,
                  b.fibers[0].ctx.bid.y, waiting_warp, waiting_block);
        }
    }
    g_block = nullptr;
}

// A lane that has to wait hands the processor straight to the next runnable lane of its own warp (one switch per lane and
// collective); only when the whole warp is blocked does control go back to the scheduler.
void wait_in_warp(Fiber &f)
{
    Fiber *first = &g_block->fibers[(size_t)(&f - g_block->fibers.data()) & ~(size_t)31];
    // VRT_EMU_ORDER=random:<seed>: the search for the next lane starts at a pseudo-random lane (xorshift), so lanes interleave
    // differently at every collective
    unsigned jump = 0;
    if (g_random)
    {
        g_random ^= g_random << 13;
        g_random ^= g_random >> 17;
        g_random ^= g_random << 5;
        jump = g_random & 31u;
    }
    for (unsigned k = 1; k < 32; ++k)
    {
        Fiber &n = first[(g_reverse ? f.lane + 32 - k : f.lane + k + jump) & 31];
        if (runnable(n))
        {
            g_cur = &n;
            emu_switch(&f.sp, n.sp);
            return;
        }
    }
    yield_to_scheduler();
}

void flush_copies(uint32_t bar)
{
    for (size_t i = 0; i < g_pending.size();)
    {
        if (g_pending[i].bar != bar) { ++i; continue; }
        const PendingCopy c = g_pending[i];
        g_pending.erase(g_pending.begin() + (long)i);
        std::memcpy(from_shared(c.dst), c.src, c.bytes);
        // complete_tx
        int32_t *st = (int32_t *)from_shared(c.bar);
        st[1] -= (int32_t)c.bytes;
    }
}
} // namespace

static std::recursive_mutex g_api_mutex;
ApiLock::ApiLock() { g_api_mutex.lock(); }
ApiLock::~ApiLock() { g_api_mutex.unlock(); }

ThreadCtx &cur() { return g_cur ? g_cur->ctx : g_host_ctx; }
unsigned lane_id() { return g_cur ? g_cur->lane : 0; }
unsigned char *dynamic_smem() { return g_block ? g_block->dyn_smem : nullptr; }

void fatal(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    std::fprintf(stderr, "[cuda_emu] FATAL: ");
    std::vfprintf(stderr, fmt, ap);
    std::fprintf(stderr, "\n");
    va_end(ap);
    std::fflush(stderr);
    std::abort();
}

void warp_gather(unsigned mask, uint64_t v, uint64_t out[32], int kind)
{
    if (!g_cur) fatal("warp collective outside a kernel");
    Fiber &f = *g_cur;
    Warp &w = *f.warp;
    const unsigned bit = 1u << f.lane;
    if (!(mask & bit)) fatal("lane %u calls a collective whose mask %08x does not name it", f.lane, mask);
    if (w.arrived == 0)
    {
        w.kind = kind;
        w.mask = mask;
    }
    else if (w.kind != kind || w.mask != mask)
        fatal("divergent collectives in one warp (block %u, warp of thread %u): kind %d mask %08x meets kind %d mask %08x", f.ctx.bid.x, f.ctx.tid.x, kind, mask,
              w.kind, w.mask);
    if (w.arrived & bit) fatal("lane %u arrived twice at one collective", f.lane);
    w.vals[f.lane] = v;
    w.arrived |= bit;
    const unsigned gen = w.gen;
    warp_try_complete(w);
    while (w.gen == gen)
    {
        f.wait_gen = &w.gen;
        f.wait_val = gen;
        wait_in_warp(f);
    }
    f.wait_gen = nullptr;
    std::memcpy(out, w.res, sizeof(w.res));
}

void block_barrier()
{
    if (!g_cur) fatal("__syncthreads outside a kernel");
    Block &b = *g_block;
    Fiber &f = *g_cur;
    const unsigned gen = b.bar_gen;
    b.bar_arrived++;
    block_try_complete(b);
    while (b.bar_gen == gen)
    {
        f.wait_gen = &b.bar_gen;
        f.wait_val = gen;
        yield_to_scheduler();
    }
    f.wait_gen = nullptr;
}

void run_grid(dim3 grid, dim3 block, size_t smem, void (*body)(void *), void *arg)
{
    if (g_block) fatal("nested kernel launch");
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024 || nthreads % 32) fatal("unsupported block size %u", nthreads);
    if ((uint64_t)grid.x * grid.y * grid.z == 0) fatal("empty grid"); // cudaErrorInvalidConfiguration on the GPU
    static const char *lazy = std::getenv("VRT_EMU_TMA");
    g_tma_lazy = !(lazy && std::strcmp(lazy, "eager") == 0);
    static const char *order = std::getenv("VRT_EMU_ORDER");
    g_reverse = order && std::strcmp(order, "reverse") == 0;
    if (order && std::strncmp(order, "random:", 7) == 0 && g_random == 0) g_random = (uint32_t)std::atoi(order + 7) * 2654435761u + 1u;
    std::vector<unsigned char> dyn(smem + 128);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx)
            {
                Block b;
                b.body = body;
                b.arg = arg;
                b.dyn_smem = (unsigned char *)(((uintptr_t)dyn.data() + 127) & ~(uintptr_t)127);
                std::memset(dyn.data(), 0xC3, dyn.size()); // shared memory is not initialised on the GPU either
                b.fibers.resize(nthreads);
                b.warps.resize(nthreads / 32);
                b.alive = nthreads;
                for (unsigned t = 0; t < nthreads; ++t)
                {
                    Fiber &f = b.fibers[t];
                    f.ctx.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    f.ctx.bid = uint3{bx, by, bz};
                    f.ctx.bdim = block;
                    f.ctx.gdim = grid;
                    f.lane = t & 31;
                    f.warp = &b.warps[t / 32];
                    f.warp->alive |= 1u << f.lane;
                    fiber_init(f);
                }
                run_block(b);
                if (!g_pending.empty()) fatal("%zu bulk copies were issued but never waited for", g_pending.size());
            }
    check_guards("after a kernel");
}

// ---- shared-memory handles, mbarrier and bulk-copy model -------------------------------------------------------------------
// Static __shared__ arrays are statics of this shared object, within +-2 GB of this anchor.
static char g_anchor;
size_t to_shared(const void *p)
{
    const intptr_t d = (const char *)p - &g_anchor;
    if (d != (intptr_t)(int32_t)d) fatal("shared-memory address out of the 32-bit handle range");
    return (size_t)(uint32_t)(int32_t)d;
}
void *from_shared(uint32_t h) { return &g_anchor + (intptr_t)(int32_t)h; }

// barrier state in the product's 8 bytes: { phase | pending arrivals << 8 | count << 16 , outstanding tx bytes }
void mbar_init(uint32_t bar, uint32_t count)
{
    int32_t *st = (int32_t *)from_shared(bar);
    st[0] = (int32_t)((count << 8) | (count << 16));
    st[1] = 0;
}
void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    int32_t *st = (int32_t *)from_shared(bar);
    uint32_t s = (uint32_t)st[0];
    const uint32_t pending = (s >> 8) & 0xff;
    if (pending == 0) fatal("mbarrier: more arrivals than its count");
    s = (s & ~0xff00u) | ((pending - 1) << 8);
    st[0] = (int32_t)s;
    st[1] += (int32_t)bytes;
}
uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    flush_copies(bar);
    int32_t *st = (int32_t *)from_shared(bar);
    uint32_t s = (uint32_t)st[0];
    if (((s >> 8) & 0xff) == 0 && st[1] == 0)
    {
        // phase complete: flip the parity, re-arm the arrival count
        const uint32_t count = (s >> 16) & 0xff;
        s = ((s & 1u) ^ 1u) | (count << 8) | (count << 16);
        st[0] = (int32_t)s;
    }
    if ((s & 1u) == (parity & 1u))
    {
        // not complete yet (the issuing lane has not run that far): let the other fibers run; the product's bounded spin
        // loop traps if it never completes
        if (g_cur)
        {
            g_cur->wait_gen = nullptr;
            yield_to_scheduler();
        }
        return 0u;
    }
    return 1u;
}
void bulk_copy(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    if (bytes == 0 || bytes % 16 || ((uintptr_t)src & 15) || ((uintptr_t)from_shared(dst) & 15))
        fatal("cp.async.bulk needs 16-byte aligned addresses and a size that is a multiple of 16 (dst %p src %p bytes %u)", from_shared(dst), src, bytes);
    g_pending.push_back(PendingCopy{dst, bytes, bar, src});
    if (!g_tma_lazy) flush_copies(bar);
}
} // namespace emu

// ---- runtime API ---------------------------------------------------------------------------------------------------------
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "invalid value"); }
// VRT_EMU_DEVICES pretends to have several GPUs (contexts are independent; the API lock serialises them)
static int device_count()
{
    const char *n = std::getenv("VRT_EMU_DEVICES");
    return n ? std::max(1, std::min(16, std::atoi(n))) : 1;
}
cudaError_t cudaGetDeviceCount(int *n)
{
    *n = device_count();
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int d) { return d >= 0 && d < device_count() ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    std::memset(p, 0, sizeof(*p));
    p->major = 10;
    p->minor = 0;
    const char *sms = std::getenv("VRT_EMU_SMS");
    p->multiProcessorCount = sms ? std::max(1, std::atoi(sms)) : 2;
    std::snprintf(p->name, sizeof(p->name), "cuda_emu (CPU SIMT interpreter, tests only)");
    return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n)
{
    unsigned char *base = (unsigned char *)std::malloc(n + 2 * emu::GUARD);
    if (!base) return cudaErrorMemoryAllocation;
    std::memset(base, emu::GUARD_BYTE, emu::GUARD);
    std::memset(base + emu::GUARD, emu::FRESH_BYTE, n); // device memory comes back uninitialised
    std::memset(base + emu::GUARD + n, emu::GUARD_BYTE, emu::GUARD);
    *p = base + emu::GUARD;
    emu::g_allocs[*p] = n;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    auto it = emu::g_allocs.find(p);
    if (it == emu::g_allocs.end()) return cudaErrorInvalidValue;
    emu::check_guards("at cudaFree");
    emu::g_allocs.erase(it);
    std::free((unsigned char *)p - emu::GUARD);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t)
{
    if (n && (!dst || !src)) return cudaErrorInvalidValue;
    if (n) std::memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *dst, int v, size_t n, cudaStream_t)
{
    if (n && !dst) return cudaErrorInvalidValue;
    if (n) std::memset(dst, v, n);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned)
{
    *s = (cudaStream_t) new char;
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s)
{
    delete (char *)s;
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e)
{
    *e = new emuEvent{std::chrono::steady_clock::now()};
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t)
{
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e)
{
    delete e;
    return cudaSuccess;
}
