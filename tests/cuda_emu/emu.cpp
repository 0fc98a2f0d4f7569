// tests/cuda_emu/emu.cpp -- TEST INFRASTRUCTURE ONLY: fiber scheduler behind cuda_emu/cuda_runtime.h.
// One CTA at a time; one ucontext fiber per CUDA thread; __syncthreads yields to the scheduler,
// which resumes the fibers once every live fiber of the CTA has arrived.
#include "cuda_runtime.h"
#include <ucontext.h>
#include <stdio.h>
#include <vector>

namespace pft_emu {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char *g_dyn_smem = nullptr;

enum { READY = 0, WAITING = 1, DONE = 2 };
struct Fiber { ucontext_t ctx; int state; uint3 tid; };

static const size_t STACK = 96 * 1024;
static std::vector<Fiber> fibers;
static std::vector<unsigned char> stacks;
static std::vector<unsigned char> smem_buf;
static ucontext_t sched_ctx;
static int cur = -1;
static const std::function<void()> *cur_body = nullptr;

static void trampoline()
{
	(*cur_body)();
	fibers[cur].state = DONE;
	swapcontext(&fibers[cur].ctx, &sched_ctx);
}

void syncthreads()
{
	fibers[cur].state = WAITING;
	swapcontext(&fibers[cur].ctx, &sched_ctx);
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body)
{
	const int nt = (int)(block.x * block.y * block.z);
	if ((int)fibers.size() < nt) fibers.resize(nt);
	if (stacks.size() < (size_t)nt * STACK) stacks.resize((size_t)nt * STACK);
	if (smem_buf.size() < smem_bytes + 64) smem_buf.resize(smem_bytes + 64);
	g_dyn_smem = smem_buf.data();
	g_blockDim = block; g_gridDim = grid;
	cur_body = &body;
	for (unsigned bz = 0; bz < grid.z; bz++)
	for (unsigned by = 0; by < grid.y; by++)
	for (unsigned bx = 0; bx < grid.x; bx++) {
		g_blockIdx = uint3{bx, by, bz};
		memset(smem_buf.data(), 0xff, smem_buf.size());      // poison: uninitialised shared memory reads as NaN
		int t = 0;
		for (unsigned z = 0; z < block.z; z++) for (unsigned y = 0; y < block.y; y++) for (unsigned x = 0; x < block.x; x++, t++) {
			Fiber &f = fibers[t];
			getcontext(&f.ctx);
			f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * STACK;
			f.ctx.uc_stack.ss_size = STACK;
			f.ctx.uc_link = &sched_ctx;
			makecontext(&f.ctx, trampoline, 0);
			f.state = READY; f.tid = uint3{x, y, z};
		}
		int alive = nt;
		while (alive > 0) {
			int waiting = 0;
			for (int i = 0; i < nt; i++) {
				if (fibers[i].state != READY) continue;
				cur = i; g_threadIdx = fibers[i].tid;
				swapcontext(&sched_ctx, &fibers[i].ctx);
				if (fibers[i].state == DONE) alive--;
				else waiting++;
			}
			// CUDA requires every non-exited thread of the CTA to reach the barrier
			for (int i = 0; i < nt; i++) if (fibers[i].state == WAITING) fibers[i].state = READY;
			(void)waiting;
		}
	}
	cur = -1;
}

}  // namespace pft_emu
