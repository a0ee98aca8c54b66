// tests/emu/cuda_emu.cpp — fibers + the collective scheduler of the warp emulator (TEST INFRASTRUCTURE;
// see cuda_runtime.h in this directory).  x86-64 only: the context switch is six pushes and a stack swap.
#include "cuda_runtime.h"
#include <sys/mman.h>

#if !defined(__x86_64__)
#error "the warp emulator's context switch is written for x86-64"
#endif

extern "C" void emu_switch(void **save_sp, void *new_sp);
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

namespace emu {

dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
Fiber *g_cur = nullptr;
void *g_dyn_smem = nullptr;

static void *g_sched_sp = nullptr;
static void (*g_entry)(void *) = nullptr;
static void *g_args = nullptr;
static const size_t STACK = 192 << 10;

static void fiber_main() {
    g_entry(g_args);
    g_cur->state = 3;
    emu_switch(&g_cur->sp, g_sched_sp);
    abort();
}

static void resume(Fiber *f) {
    g_cur = f;
    g_threadIdx = f->tid;
    emu_switch(&g_sched_sp, f->sp);
    g_cur = nullptr;
}

unsigned long long collective(int op, unsigned long long v, int arg) {
    Fiber *f = g_cur;
    f->op = op; f->in = v; f->arg = arg;
    f->state = (op == OP_SYNCTHREADS) ? 2 : 1;
    emu_switch(&f->sp, g_sched_sp);
    g_threadIdx = f->tid;
    return f->out;
}

static void die(const char *what, unsigned b, unsigned w) {
    fprintf(stderr, "emu: %s (block %u, warp %u)\n", what, b, w);
    abort();
}

static void resolve_warp(Fiber *L, int n, unsigned b, unsigned w) {
    const int op = L[0].op;
    for (int i = 1; i < n; ++i)
        if (L[i].op != op) { fprintf(stderr, "emu: lane 0 at collective %d, lane %d at %d\n", op, i, L[i].op); die("divergent warp collective", b, w); }
    switch (op) {
    case OP_BALLOT: case OP_ANY: {
        unsigned m = 0; for (int i = 0; i < n; ++i) if (L[i].in) m |= 1u << i;
        for (int i = 0; i < n; ++i) L[i].out = (op == OP_ANY) ? (m != 0) : m;
        break; }
    case OP_SHFL: { unsigned long long t[32]; for (int i = 0; i < n; ++i) t[i] = L[i].in;
        for (int i = 0; i < n; ++i) L[i].out = t[L[i].arg & 31]; break; }
    case OP_SHFL_UP: { unsigned long long t[32]; for (int i = 0; i < n; ++i) t[i] = L[i].in;
        for (int i = 0; i < n; ++i) L[i].out = (i - L[i].arg >= 0) ? t[i - L[i].arg] : t[i]; break; }
    case OP_SHFL_DOWN: { unsigned long long t[32]; for (int i = 0; i < n; ++i) t[i] = L[i].in;
        for (int i = 0; i < n; ++i) L[i].out = (i + L[i].arg < n) ? t[i + L[i].arg] : t[i]; break; }
    case OP_REDUCE_ADD: { unsigned s = 0; for (int i = 0; i < n; ++i) s += (unsigned)L[i].in; for (int i = 0; i < n; ++i) L[i].out = s; break; }
    case OP_REDUCE_OR: { unsigned s = 0; for (int i = 0; i < n; ++i) s |= (unsigned)L[i].in; for (int i = 0; i < n; ++i) L[i].out = s; break; }
    case OP_SYNCWARP: break;
    default: die("unknown collective", b, w);
    }
    for (int i = 0; i < n; ++i) { L[i].state = 0; L[i].op = OP_NONE; }
}

void launch(void (*entry)(void *), void *args, size_t, dim3 grid, dim3 block, size_t smem) {
    const unsigned nthr = block.x * block.y * block.z;
    if (nthr == 0 || nthr % 32 != 0) { fprintf(stderr, "emu: block size %u must be a multiple of 32\n", nthr); abort(); }
    const unsigned nwarps = nthr / 32;
    static std::vector<char *> stacks;
    while (stacks.size() < nthr) {
        void *p = mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("emu: mmap"); abort(); }
        stacks.push_back((char *)p);
    }
    std::vector<Fiber> F(nthr);
    std::vector<char> dyn(smem + 64);
    g_entry = entry; g_args = args; g_blockDim = block; g_gridDim = grid;
    g_dyn_smem = (void *)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
    for (unsigned b = 0; b < grid.x; ++b) {
        g_blockIdx = dim3(b, 0, 0);
        for (unsigned t = 0; t < nthr; ++t) {
            Fiber &f = F[t];
            f = Fiber();
            f.stack = stacks[t];
            f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            // initial frame: six callee-saved registers (zero) + return address = fiber_main; after the `ret`
            // the stack pointer must be 8 modulo 16, as after a call
            uintptr_t top = ((uintptr_t)f.stack + STACK) & ~(uintptr_t)15;
            void **sp = (void **)(top - 8);           // where fiber_main's "return address slot" would be
            *--sp = (void *)fiber_main;
            for (int i = 0; i < 6; ++i) *--sp = nullptr;
            f.sp = sp;
        }
        unsigned done = 0;
        while (done < nthr) {
            bool progress = false;
            unsigned at_barrier = 0;
            done = 0;
            for (unsigned w = 0; w < nwarps; ++w) {
                Fiber *L = &F[w * 32];
                int n_wait = 0, n_done = 0, n_cta = 0;
                for (int i = 0; i < 32; ++i) {
                    if (L[i].state == 0) { resume(&L[i]); progress = true; }
                    n_wait += L[i].state == 1; n_cta += L[i].state == 2; n_done += L[i].state == 3;
                }
                if (n_wait == 32) { resolve_warp(L, 32, b, w); progress = true; }
                else if (n_wait && (n_done || n_cta)) die("some lanes wait at a warp collective while others exited or sit at __syncthreads", b, w);
                at_barrier += n_cta; done += n_done;
            }
            if (at_barrier && at_barrier + done == nthr) {
                if (done) die("__syncthreads reached by only part of the CTA", b, 0);
                for (unsigned t = 0; t < nthr; ++t) { F[t].state = 0; F[t].op = OP_NONE; }
                progress = true;
            }
            if (!progress && done < nthr) die("deadlock", b, 0);
        }
    }
    g_dyn_smem = nullptr;
}

}  // namespace emu
