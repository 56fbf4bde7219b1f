// tests/emu/cuda_emu.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See cuda_emu.h.
#include <string>      // before cuda_emu.h: its __noinline__ macro would break libstdc++ headers
#include "cuda_emu.h"

namespace emu {

BlockCtx* g_blk = nullptr;
Fiber* g_cur = nullptr;
uint3_ g_blockIdx{0, 0, 0};
dim3 g_blockDim, g_gridDim;

static const size_t kStack = 256 * 1024;
static unsigned char* g_stacks = nullptr;
static size_t g_stacks_n = 0;
static unsigned long long g_spin = 0;
static int g_order = 0;                         // 0 forward, 1 reverse, 2 random
int g_preempt = 0;                              // 1: a thread yields after every atomic
static unsigned long long g_rng = 1;
static inline unsigned next_rand() {
    g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull;
    return (unsigned)(g_rng >> 33);
}
static void read_order() {
    const char* env = std::getenv("IPB_EMU_ORDER");
    g_order = g_preempt = 0;
    if (!env || !*env) return;
    std::string v(env);
    const size_t pp = v.find(":preempt");
    if (pp != std::string::npos) { g_preempt = 1; v.erase(pp, 8); }
    const char* e = v.c_str();
    if (!*e || !std::strcmp(e, "forward")) return;
    if (!std::strcmp(e, "reverse")) { g_order = 1; return; }
    if (!std::strncmp(e, "random", 6) && (e[6] == 0 || e[6] == ':')) {
        g_order = 2;
        g_rng = e[6] == ':' ? std::strtoull(e + 7, nullptr, 10) * 2 + 1 : 12345;
        return;
    }
    std::fprintf(stderr, "cuda_emu: IPB_EMU_ORDER=%s not understood (forward | reverse | random[:seed])\n", e);
    std::abort();
}

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

void yield() {
    if (++g_spin > 2000000000ull) {
        std::fprintf(stderr, "cuda_emu: deadlock suspected (block %u,%u thread %u)\n",
                     g_blockIdx.x, g_blockIdx.y, g_cur ? g_cur->linear : 0);
        std::abort();
    }
    Fiber* me = g_cur;
    emu_switch(&me->sp, g_blk->sched_sp);
}

static void fiber_entry() {
    g_blk->body();
    g_cur->done = true;
    g_blk->alive--;
    // a finished thread may complete a barrier the others are waiting on
    Fiber* me = g_cur;
    emu_switch(&me->sp, g_blk->sched_sp);
    std::abort();  // never resumed
}

static void run_block(BlockCtx& b) {
    g_blk = &b;
    unsigned n = b.nthreads;
    if (g_stacks_n < n) {
        if (g_stacks) munmap(g_stacks, g_stacks_n * kStack);
        g_stacks = (unsigned char*)mmap(nullptr, (size_t)n * kStack, PROT_READ | PROT_WRITE,
                                        MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_stacks == MAP_FAILED) { std::perror("mmap"); std::abort(); }
        g_stacks_n = n;
    }
    for (unsigned i = 0; i < n; ++i) {
        Fiber& f = b.fibers[i];
        f.done = false;
        uintptr_t top = (uintptr_t)(g_stacks + (size_t)(i + 1) * kStack);
        top &= ~(uintptr_t)15;
        // layout (low->high): r15 r14 r13 r12 rbx rbp ret ; after `ret`, rsp % 16 == 8
        void** sp = (void**)(top - 8);
        *--sp = (void*)fiber_entry;
        for (int k = 0; k < 6; ++k) *--sp = nullptr;
        f.sp = (void*)sp;
    }
    b.alive = n;
    b.bar_arrived = 0;
    b.bar_gen = 0;
    // Scheduling order of the block's threads between two yields (IPB_EMU_ORDER): "forward" (default:
    // thread 0 runs to its next barrier / collective, then thread 1, ...), "reverse", or "random[:seed]"
    // (a fresh permutation of the warps and of the lanes inside each warp on every sweep).  Code that is
    // correctly synchronised gives the same results under every order; a missing barrier between a
    // producer and a consumer phase shows up under at least one of them (tests/test_emu_orders.py).
    unsigned live = n;
    std::vector<unsigned> order(n);
    for (unsigned i = 0; i < n; ++i) order[i] = g_order == 1 ? n - 1 - i : i;
    while (live) {
        live = 0;
        if (g_order == 2) {
            const unsigned nw = (n + 31) / 32;
            std::vector<unsigned> wp(nw);
            for (unsigned w = 0; w < nw; ++w) wp[w] = w;
            for (unsigned w = nw; w > 1; --w) std::swap(wp[w - 1], wp[next_rand() % w]);
            unsigned k = 0;
            for (unsigned w = 0; w < nw; ++w) {
                unsigned lo = wp[w] * 32, cnt = std::min(32u, n - lo), first = k;
                for (unsigned l = 0; l < cnt; ++l) order[k++] = lo + l;
                for (unsigned l = cnt; l > 1; --l) std::swap(order[first + l - 1], order[first + next_rand() % l]);
            }
        }
        for (unsigned k = 0; k < n; ++k) {
            Fiber& f = b.fibers[order[k]];
            if (f.done) continue;
            g_cur = &f;
            emu_switch(&b.sched_sp, f.sp);
            if (!f.done) ++live;
        }
    }
    g_cur = nullptr;
}

void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body) {
    BlockCtx b;
    b.nthreads = block.x * block.y * block.z;
    b.fibers.resize(b.nthreads);
    b.warps.resize((b.nthreads + 31) / 32);
    b.body = body;
    std::vector<unsigned char> smem(dyn_smem + 64);
    b.dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
    g_blockDim = block;
    g_gridDim = grid;
    unsigned lin = 0;
    for (unsigned z = 0; z < block.z; ++z)
        for (unsigned y = 0; y < block.y; ++y)
            for (unsigned x = 0; x < block.x; ++x) {
                b.fibers[lin].tid = uint3_{x, y, z};
                b.fibers[lin].linear = lin;
                ++lin;
            }
    read_order();
    // blocks run one after the other; under "reverse" / "random" the grid is walked backwards, so a kernel
    // whose blocks depend on each other's order (they must not) is seen as well
    const unsigned long long nb = (unsigned long long)grid.x * grid.y * grid.z;
    for (unsigned long long q = 0; q < nb; ++q) {
        const unsigned long long lin_b = g_order ? nb - 1 - q : q;
        g_blockIdx = uint3_{(unsigned)(lin_b % grid.x), (unsigned)((lin_b / grid.x) % grid.y),
                            (unsigned)(lin_b / ((unsigned long long)grid.x * grid.y))};
        for (auto& w : b.warps) { w.arrived = w.departed = 0; }
        g_spin = 0;
        std::memset(b.dyn_smem, 0xA5, dyn_smem);      // a block never finds its dynamic shared memory zeroed (or as the last block left it)
        run_block(b);
    }
    g_blk = nullptr;
}

}  // namespace emu
