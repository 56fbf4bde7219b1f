// tests/emu/cuda_emu.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See cuda_emu.h.
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
    unsigned live = n;
    while (live) {
        live = 0;
        for (unsigned i = 0; i < n; ++i) {
            Fiber& f = b.fibers[i];
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
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = uint3_{bx, by, bz};
                for (auto& w : b.warps) { w.arrived = w.departed = 0; }
                g_spin = 0;
                run_block(b);
            }
    g_blk = nullptr;
}

}  // namespace emu
