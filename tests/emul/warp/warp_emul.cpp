// TEST INFRASTRUCTURE -- the scheduler behind tests/emul/warp/cuda_runtime.h: one fiber per CUDA thread, blocks one after
// another, lanes of a warp in lock-step at every *_sync collective.  Single OS thread; not re-entrant (a global lock
// serialises kernel launches of different host threads).
#include <cuda_runtime.h>
#include <stdio.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <mutex>
#include <vector>

// Thread-sanitizer build (kernels.cu / engine.cu with -fsanitize=thread, this file with -DEMU_TSAN but NOT instrumented: the
// scheduler's own bookkeeping is shared between fibers by design): every fiber is a TSan fiber, switches carry NO synchronisation, and the only
// happens-before edges are the ones the CUDA model gives: a completed warp collective orders its participants, __syncthreads
// the block, block / kernel boundaries everything (blocks run one after another here, so races BETWEEN blocks are not looked
// for).  Two lanes touching the same shared or global word without a collective in between are reported as a data race --
// what compute-sanitizer racecheck looks for, on the CPU.
#if defined(EMU_TSAN)
#include <sanitizer/tsan_interface.h>
#define TSAN_RELEASE(p) __tsan_release(p)
#define TSAN_ACQUIRE(p) __tsan_acquire(p)
#else
#define TSAN_RELEASE(p) ((void)0)
#define TSAN_ACQUIRE(p) ((void)0)
#endif

EmuThread *emu_self;
uint3 emu_block_idx;
dim3 emu_block_dim, emu_grid_dim;

void emu_recheck(int warp);

namespace {

struct Collective {
    unsigned mask = 0, arrived = 0;
    int op = 0, width = 32;
    uint32_t val[32];
    int arg[32];
};
struct Fiber : EmuThread {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false, waiting = false;      // waiting: parked in a collective or the block barrier until released
    uint32_t result = 0;
    int warp = 0, lane = 0;
    void *tsan = nullptr;                    // TSan fiber context
    void *acquire = nullptr;                 // sync object to acquire when the fiber resumes after a collective / barrier
};
struct Warp {
    unsigned exited = 0;
    char sync[32];                           // TSan sync objects of the collectives, one per lowest lane of the mask
    Collective open[8];                      // collectives in progress, keyed by mask (disjoint lane groups sync independently);
                                             // mask == 0 = free slot (fixed slots, no copies: a sanitizer build intercepts memmove)
};

constexpr size_t STACK = 256 << 10;
std::mutex g_launch;
std::vector<Fiber> g_fib;
std::vector<Warp> g_warp;
ucontext_t g_sched;
Fiber *g_cur;
const std::function<void()> *g_body;
const char *g_name;
int g_barrier_arrived, g_alive;
char g_sync_block, g_sync_grid;              // TSan sync objects: __syncthreads, block / kernel boundaries
void *g_tsan_sched;
unsigned long g_switches;
std::vector<unsigned> g_order;
int g_order_kind = -1;

void to_scheduler()
{
#if defined(EMU_TSAN)
    __tsan_switch_to_fiber(g_tsan_sched, __tsan_switch_to_fiber_no_sync);
#endif
    Fiber *f = g_cur;
    swapcontext(&f->ctx, &g_sched);
    if (f->acquire) { TSAN_ACQUIRE(f->acquire); f->acquire = nullptr; }
}

void fiber_main()
{
    TSAN_ACQUIRE(&g_sync_grid);              // everything before this block (host copies, earlier kernels and blocks)
    (*g_body)();
    Fiber *f = g_cur;
    TSAN_RELEASE(&g_sync_grid);
    f->done = true;
    g_alive--;
    Warp &w = g_warp[f->warp];
    w.exited |= 1u << f->lane;
    // a lane that leaves may complete a collective the others are parked in (they only wait for lanes still alive)
    emu_recheck(f->warp);
    if (g_barrier_arrived && g_barrier_arrived == g_alive) {          // ... or the block barrier
        for (Fiber &o : g_fib) o.waiting = false;
        g_barrier_arrived = 0;
    }
    to_scheduler();
}

uint32_t resolve(const Collective &c, int lane)
{
    const int seg = lane & ~(c.width - 1), end = seg + c.width - 1;
    int src = lane;
    switch (c.op) {
    case EMU_SYNCWARP: return 0;
    case EMU_SHFL_IDX: src = seg | (c.arg[lane] & (c.width - 1)); break;
    case EMU_SHFL_XOR: src = lane ^ c.arg[lane]; if (src > end || src < seg) src = lane; break;
    case EMU_SHFL_UP: src = lane - c.arg[lane]; if (src < seg) src = lane; break;
    case EMU_SHFL_DOWN: src = lane + c.arg[lane]; if (src > end) src = lane; break;
    case EMU_ANY: case EMU_ALL: case EMU_BALLOT: {
        unsigned b = 0;
        for (int l = 0; l < 32; l++) if ((c.arrived >> l & 1) && c.val[l]) b |= 1u << l;
        return c.op == EMU_ANY ? b != 0 : c.op == EMU_ALL ? b == c.arrived : b;
    }
    }
    return (c.arrived >> src & 1) ? c.val[src] : c.val[lane];          // reading a lane that does not take part is undefined on the device
}

bool try_complete(int warp, size_t k)
{
    Warp &w = g_warp[warp];
    Collective &c = w.open[k];
    if (!c.mask || c.arrived != (c.mask & ~w.exited)) return false;
    for (int l = 0; l < 32; l++)
        if (c.arrived >> l & 1) {
            Fiber &f = g_fib[warp * 32 + l];
            f.result = resolve(c, l);
            f.waiting = false;
        }
    c.mask = c.arrived = 0;
    return true;
}

}  // namespace

void emu_recheck(int warp)
{
    for (size_t k = 0; k < 8; k++) try_complete(warp, k);
}

uint32_t emu_collective(unsigned mask, int op, uint32_t value, int arg, int width)
{
    Fiber *f = g_cur;
    Warp &w = g_warp[f->warp];
    if (!(mask >> f->lane & 1)) { fprintf(stderr, "emu: %s: lane %d calls a collective whose mask %08x excludes it\n", g_name, f->lane, mask); abort(); }
    size_t k = 0;
    while (k < 8 && !(w.open[k].mask == mask && !(w.open[k].arrived >> f->lane & 1))) k++;
    if (k == 8) {
        for (k = 0; k < 8 && w.open[k].mask; k++) {}
        if (k == 8) { fprintf(stderr, "emu: %s: more than 8 collectives in progress in one warp\n", g_name); abort(); }
        w.open[k].mask = mask; w.open[k].arrived = 0; w.open[k].op = op; w.open[k].width = width;
    }
    Collective &c = w.open[k];
    if (c.op != op || c.width != width) { fprintf(stderr, "emu: %s: lanes of one warp meet in different collectives (mask %08x: op %d vs %d)\n", g_name, mask, c.op, op); abort(); }
    c.arrived |= 1u << f->lane;
    c.val[f->lane] = value;
    c.arg[f->lane] = arg;
    f->waiting = true;
    void *sync = &w.sync[__builtin_ctz(mask)];
    TSAN_RELEASE(sync);                      // every participant releases on arrival and acquires when it goes on: all-to-all order
    if (!try_complete(f->warp, k)) { f->acquire = sync; to_scheduler(); }
    else TSAN_ACQUIRE(sync);
    return f->result;
}

void emu_syncthreads()
{
    Fiber *f = g_cur;
    g_barrier_arrived++;
    TSAN_RELEASE(&g_sync_block);
    if (g_barrier_arrived == g_alive) {
        for (Fiber &o : g_fib) o.waiting = false;
        g_barrier_arrived = 0;
        TSAN_ACQUIRE(&g_sync_block);
        return;
    }
    f->waiting = true;
    f->acquire = &g_sync_block;
    to_scheduler();
}

void emu_yield() { if (g_cur) to_scheduler(); }

void emu_run_grid(dim3 grid, dim3 block, const std::function<void()> &body, const char *name)
{
    std::lock_guard<std::mutex> lock(g_launch);
    const unsigned n = block.x * block.y * block.z;
    if (n == 0 || grid.x == 0 || grid.y == 0 || grid.z == 0) return;
    if (g_fib.size() < n) {
        const size_t old = g_fib.size();
        g_fib.resize(n);
        for (size_t i = old; i < n; i++) {
            g_fib[i].stack = (char *)mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (g_fib[i].stack == MAP_FAILED) { perror("emu: fiber stack"); abort(); }
        }
    }
    // The order in which runnable fibers get the CPU between two collectives.  Correct kernels do not care; one that reads what
    // a neighbouring lane wrote without a barrier in between passes in one order and fails in another.  B200_EMUL_ORDER =
    // "reverse" or "random[:seed]" (default: ascending thread index).
    if (g_order.size() != n || g_order_kind < 0) {
        const char *e = getenv("B200_EMUL_ORDER");
        g_order_kind = !e ? 0 : !strncmp(e, "reverse", 7) ? 1 : !strncmp(e, "random", 6) ? 2 : 0;
        g_order.resize(n);
        for (unsigned t = 0; t < n; t++) g_order[t] = g_order_kind == 1 ? n - 1 - t : t;
        if (g_order_kind == 2) {
            uint64_t x = 0x9e3779b97f4a7c15ull ^ (e[6] == ':' ? strtoull(e + 7, nullptr, 0) : 1);
            for (unsigned t = n - 1; t > 0; t--) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                std::swap(g_order[t], g_order[x % (t + 1)]);
            }
        }
    }
#if defined(EMU_TSAN)
    g_tsan_sched = __tsan_get_current_fiber();
#endif
    g_body = &body;
    g_name = name;
    emu_block_dim = block;
    emu_grid_dim = grid;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                emu_block_idx = uint3{ bx, by, bz };
                g_warp.assign((n + 31) / 32, Warp());
                if (n & 31) g_warp.back().exited = ~0u << (n & 31);       // lanes that do not exist
                g_alive = (int)n;
                g_barrier_arrived = 0;
                TSAN_RELEASE(&g_sync_grid);
                for (unsigned t = 0; t < n; t++) {
                    Fiber &f = g_fib[t];
                    f.tid = uint3{ t % block.x, t / block.x % block.y, t / (block.x * block.y) };
                    f.done = f.waiting = false;
                    f.acquire = nullptr;
#if defined(EMU_TSAN)
                    if (!f.tsan) f.tsan = __tsan_create_fiber(0);     // reused by later blocks: creating one costs milliseconds
#endif
                    f.warp = (int)t / 32;
                    f.lane = (int)t & 31;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = STACK;
                    f.ctx.uc_link = nullptr;
                    makecontext(&f.ctx, fiber_main, 0);
                }
                for (int left = (int)n; left;) {
                    bool ran = false;
                    for (unsigned k = 0; k < n; k++) {
                        Fiber &f = g_fib[g_order[k]];
                        if (f.done || f.waiting) continue;
                        g_cur = &f;
                        emu_self = &f;
                        g_switches++;
#if defined(EMU_TSAN)
                        __tsan_switch_to_fiber(f.tsan, __tsan_switch_to_fiber_no_sync);
#endif
                        swapcontext(&g_sched, &f.ctx);
                        ran = true;
                        if (f.done) left--;
                    }
                    if (!ran && left) {
                        fprintf(stderr, "emu: %s: block (%u,%u,%u) is stuck: %d threads wait in collectives nobody completes\n", name, bx, by, bz, left);
                        for (size_t wi = 0; wi < g_warp.size(); wi++)
                            for (const Collective &c : g_warp[wi].open)
                                if (c.mask) fprintf(stderr, "  warp %zu: op %d mask %08x arrived %08x exited %08x\n", wi, c.op, c.mask, c.arrived, g_warp[wi].exited);
                        abort();
                    }
                }
            }
    TSAN_ACQUIRE(&g_sync_grid);
    g_cur = nullptr;
    emu_self = nullptr;
}

// ---- runtime API stubs that own memory -------------------------------------------------------------------------------------
cudaError_t cudaMalloc(void **p, size_t n)
{
    if (posix_memalign(p, 256, n ? n : 1)) return cudaErrorInvalidValue;     // exactly n bytes: an address sanitizer build sees every overrun
    memset(*p, 0xA5, n);                       // device memory comes uninitialised: make reads of it show
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return posix_memalign(p, 4096, n ? n : 1) ? cudaErrorInvalidValue : cudaSuccess; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { memset(a, 0, sizeof(*a)); a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
struct EmuStream { int unused; };
struct EmuEvent { int unused; };
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new EmuStream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new EmuEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
