// Fiber scheduler of the CUDA-on-CPU test shim (see cuda_emu.h).
#include "cuda_emu.h"

#include <omp.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <vector>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace emu {
namespace {
enum State { READY = 0, WAIT_BLOCK = 1, WAIT_WARP = 2, DONE = 3 };
constexpr size_t kStack = 256 * 1024;

struct Worker {
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<char*> stacks;
    std::vector<int> state;
    std::vector<uint3> tids;
    int cur = 0;
    unsigned n = 0;
    const std::function<void()>* body = nullptr;
    std::vector<unsigned char> dyn;
    std::vector<unsigned long long> scratch;   // per warp 32 x 8 bytes
};
thread_local Worker W;

void fiber_entry() {
    (*W.body)();
    W.state[W.cur] = DONE;
    swapcontext(&W.ctx[W.cur], &W.sched);
}

void yield_as(int st) {
    W.state[W.cur] = st;
    swapcontext(&W.ctx[W.cur], &W.sched);
}

void run_block(dim3 block, size_t smem, const std::function<void()>& body) {
    const unsigned n = block.x * block.y * block.z;
    W.n = n;
    W.body = &body;
    if (W.ctx.size() < n) {
        size_t old = W.ctx.size();
        W.ctx.resize(n); W.state.resize(n); W.tids.resize(n); W.stacks.resize(n, nullptr);
        for (size_t i = old; i < n; ++i) {
            W.stacks[i] = (char*)mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (W.stacks[i] == MAP_FAILED) { perror("emu: mmap"); abort(); }
        }
    }
    if (W.dyn.size() < smem + 64) W.dyn.resize(smem + 64);
    W.scratch.assign(((n + 31) / 32) * 32, 0ull);
    for (unsigned t = 0; t < n; ++t) {
        getcontext(&W.ctx[t]);
        W.ctx[t].uc_stack.ss_sp = W.stacks[t];
        W.ctx[t].uc_stack.ss_size = kStack;
        W.ctx[t].uc_link = &W.sched;
        makecontext(&W.ctx[t], fiber_entry, 0);
        W.state[t] = READY;
        W.tids[t] = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
    }
    unsigned done = 0;
    while (done < n) {
        bool ran = false;
        for (unsigned t = 0; t < n; ++t) {
            if (W.state[t] != READY) continue;
            W.cur = (int)t;
            threadIdx = W.tids[t];
            swapcontext(&W.sched, &W.ctx[t]);
            ran = true;
            if (W.state[t] == DONE) ++done;
        }
        if (done == n) break;
        // release barriers
        unsigned nblock = 0, alive = 0;
        for (unsigned t = 0; t < n; ++t) { alive += W.state[t] != DONE; nblock += W.state[t] == WAIT_BLOCK; }
        bool released = false;
        if (nblock == alive && alive > 0) {
            for (unsigned t = 0; t < n; ++t) if (W.state[t] == WAIT_BLOCK) W.state[t] = READY;
            released = true;
        } else {
            for (unsigned w0 = 0; w0 < n; w0 += 32) {
                unsigned a = 0, ww = 0;
                for (unsigned t = w0; t < w0 + 32 && t < n; ++t) { a += W.state[t] != DONE; ww += W.state[t] == WAIT_WARP; }
                if (a > 0 && ww == a) {
                    for (unsigned t = w0; t < w0 + 32 && t < n; ++t) if (W.state[t] == WAIT_WARP) W.state[t] = READY;
                    released = true;
                }
            }
        }
        if (!released && !ran) { fprintf(stderr, "emu: barrier deadlock (divergent __syncthreads?)\n"); abort(); }
        if (!released) {
            bool any_ready = false;
            for (unsigned t = 0; t < n; ++t) any_ready |= W.state[t] == READY;
            if (!any_ready) { fprintf(stderr, "emu: barrier deadlock: %u alive, %u at block barrier\n", alive, nblock); abort(); }
        }
    }
}
}  // namespace

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const long nb = (long)grid.x * grid.y * grid.z;
    if (nb == 0) return;
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < nb; ++b) {
        gridDim = grid;
        blockDim = block;
        blockIdx = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long)grid.x * grid.y))};
        run_block(block, smem, body);
    }
}

void* dyn_smem() {
    uintptr_t p = (uintptr_t)W.dyn.data();
    return (void*)((p + 63) & ~(uintptr_t)63);
}
void sync_block() { yield_as(WAIT_BLOCK); }
void yield() { yield_as(READY); }
void sync_warp() { yield_as(WAIT_WARP); }
void* warp_scratch() { return &W.scratch[(W.cur / 32) * 32]; }
int lane_id() { return W.cur % 32; }
int warp_alive_mask_has(int lane) {
    unsigned t = (W.cur / 32) * 32 + lane;
    return t < W.n && W.state[t] != DONE;
}
}  // namespace emu
