// tests/warp_emu/emu.cpp -- TEST-ONLY (see emu.h): coroutine scheduler of the single-warp emulator.
#include "emu.h"
#include <ucontext.h>
#include <stdlib.h>
#include <stdio.h>

EmuDim emu_thread_idx[32];
EmuDim emu_block_idx, emu_block_dim = {32, 1, 1}, emu_grid_dim;
int emu_cur_lane = 0;
float smem[65536];   // `extern __shared__ float smem[]` of the kernels

static ucontext_t g_sched, g_lane_ctx[32];
static bool g_done[32];
static uint32_t g_slot[2][32];
static unsigned g_count[32];
static EmuKernelBody g_body;
static void *g_args;
static const size_t kStack = 1 << 20;
static char *g_stacks = nullptr;

static void yield_to_sched() { swapcontext(&g_lane_ctx[emu_cur_lane], &g_sched); }

uint32_t emu_exchange(uint32_t v, int src) {
    const int me = emu_cur_lane, par = g_count[me] & 1;
    g_slot[par][me] = v; g_count[me]++;
    yield_to_sched();
    return g_slot[par][src];
}
uint32_t emu_ballot(bool p) {
    const int me = emu_cur_lane, par = g_count[me] & 1;
    g_slot[par][me] = p ? 1u : 0u; g_count[me]++;
    yield_to_sched();
    uint32_t m = 0;
    for (int l = 0; l < 32; l++) m |= (g_slot[par][l] & 1u) << l;
    return m;
}
void emu_sync() { (void)emu_exchange(0, emu_cur_lane); }

static void lane_main() {
    g_body(g_args);
    g_done[emu_cur_lane] = true;
    yield_to_sched();
}

void emu_launch(EmuKernelBody body, void *args, int blocks, int blocks_y) {
    if (!g_stacks) g_stacks = (char *)malloc(kStack * 32);
    g_body = body; g_args = args;
    emu_grid_dim = {blocks, blocks_y, 1};
    for (int bb = 0; bb < blocks * blocks_y; bb++) {
        const int b = bb % blocks;
        emu_block_idx = {b, bb / blocks, 0};
        for (int l = 0; l < 32; l++) {
            emu_thread_idx[l] = {l, 0, 0};
            g_done[l] = false; g_count[l] = 0;
            getcontext(&g_lane_ctx[l]);
            g_lane_ctx[l].uc_stack.ss_sp = g_stacks + kStack * l;
            g_lane_ctx[l].uc_stack.ss_size = kStack;
            g_lane_ctx[l].uc_link = &g_sched;
            makecontext(&g_lane_ctx[l], lane_main, 0);
        }
        int remaining = 32;
        while (remaining > 0) {
            remaining = 0;
            unsigned cmin = ~0u, cmax = 0;
            for (int l = 0; l < 32; l++) {
                if (g_done[l]) continue;
                emu_cur_lane = l;
                swapcontext(&g_sched, &g_lane_ctx[l]);
                if (!g_done[l]) { remaining++; cmin = std::min(cmin, g_count[l]); cmax = std::max(cmax, g_count[l]); }
            }
            if (remaining > 0 && remaining < 32 && cmin != ~0u) {
                // some lanes returned while others wait at a collective: divergent exit
                bool any_done = false; for (int l = 0; l < 32; l++) any_done |= g_done[l];
                if (any_done) { fprintf(stderr, "warp_emu: divergent kernel exit with pending collectives\n"); abort(); }
            }
            if (remaining > 0 && cmin != cmax) { fprintf(stderr, "warp_emu: lanes disagree on the collective count (%u vs %u)\n", cmin, cmax); abort(); }
        }
    }
}
