// tests/cuda_emu/emu.cpp -- TEST INFRASTRUCTURE ONLY: fiber scheduler behind cuda_emu/cuda_runtime.h.
// One CTA at a time; one ucontext fiber per CUDA thread; __syncthreads yields to the scheduler,
// which resumes the fibers once every live fiber of the CTA has arrived.
#include "cuda_runtime.h"
#include "cuda.h"
#include <ucontext.h>
#include <stdio.h>
#include <algorithm>
#include <vector>

namespace pft_emu {

thread_local uint3 g_threadIdx, g_blockIdx;
thread_local dim3 g_blockDim, g_gridDim;
thread_local unsigned char *g_dyn_smem = nullptr;

enum { READY = 0, WAITING = 1, DONE = 2 };
struct Fiber { ucontext_t ctx; int state; uint3 tid; };

static const size_t STACK = 96 * 1024;
static thread_local std::vector<Fiber> fibers;
static thread_local std::vector<unsigned char> stacks;
static thread_local std::vector<unsigned char> smem_buf;
static thread_local ucontext_t sched_ctx;
static thread_local int cur = -1;
static thread_local const std::function<void()> *cur_body = nullptr;

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

static thread_local int w_count[64], w_gen[64], w_size[64];

void syncwarp()
{
	const int w = cur >> 5;
	const int gen = w_gen[w];
	if (++w_count[w] == w_size[w]) { w_count[w] = 0; w_gen[w]++; }
	else while (w_gen[w] == gen) yield();
}

void yield()
{
	swapcontext(&fibers[cur].ctx, &sched_ctx);      // stays READY
}

uint32_t tensor_load_3d(void *dst, const void *map_, int x, int y, int z)
{
	const CUtensorMap *m = (const CUtensorMap *)map_;
	double *d = (double *)dst;
	if ((x * 8) & 15) { fprintf(stderr, "cuda_emu: TMA inner coordinate %d is not 16-byte aligned (B200 raises an illegal-instruction fault)\n", x); abort(); }
	if ((uintptr_t)dst & 127) { fprintf(stderr, "cuda_emu: TMA shared-memory destination is not 128-byte aligned\n"); abort(); }
	const char *base = (const char *)m->base;
	for (uint32_t bz = 0; bz < m->box[2]; bz++)
		for (uint32_t by = 0; by < m->box[1]; by++)
			for (uint32_t bx = 0; bx < m->box[0]; bx++) {
				long long gx = (long long)x + bx, gy = (long long)y + by, gz = (long long)z + bz;
				double v = 0.0;
				if (gx >= 0 && gy >= 0 && gz >= 0 && gx < (long long)m->dim[0] && gy < (long long)m->dim[1] && gz < (long long)m->dim[2])
					v = *(const double *)(base + gz * (long long)m->stride[1] + gy * (long long)m->stride[0] + gx * 8);
				d[((size_t)bz * m->box[1] + by) * m->box[0] + bx] = v;
			}
	return m->box[0] * m->box[1] * m->box[2] * 8u;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body)
{
	const int nt = (int)(block.x * block.y * block.z);
	if ((int)fibers.size() < nt) fibers.resize(nt);
	if (stacks.size() < (size_t)nt * STACK) stacks.resize((size_t)nt * STACK);
	if (smem_buf.size() < smem_bytes + 256) smem_buf.resize(smem_bytes + 256);
	g_dyn_smem = (unsigned char *)(((uintptr_t)smem_buf.data() + 127) & ~(uintptr_t)127);
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
		for (int w = 0; w < 64; w++) { w_count[w] = 0; w_gen[w] = 0; w_size[w] = std::max(0, std::min(32, nt - 32 * w)); }
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
			// __syncthreads: the fibers at the barrier are released once EVERY live fiber of the CTA has arrived; fibers that
			// merely yielded (a spin on an mbarrier, a warp barrier, a flag of another device) keep running until they do
			(void)waiting;
			int at_barrier = 0;
			for (int i = 0; i < nt; i++) if (fibers[i].state == WAITING) at_barrier++;
			if (at_barrier == alive)
				for (int i = 0; i < nt; i++) if (fibers[i].state == WAITING) fibers[i].state = READY;
		}
	}
	cur = -1;
}

}  // namespace pft_emu

CUresult cuTensorMapEncodeTiled(CUtensorMap *m, CUtensorMapDataType, cuuint32_t rank, void *addr, const cuuint64_t *dim,
                                const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)
{
	if (rank != 3) return 1;
	if (((uintptr_t)addr & 15) || (strides[0] & 15) || (strides[1] & 15) || ((box[0] * 8) & 15) || box[0] > 256 || box[1] > 256) return 1;   // the real constraints
	memset(m, 0, sizeof *m);
	m->base = addr;
	for (int i = 0; i < 3; i++) { m->dim[i] = dim[i]; m->box[i] = box[i]; }
	m->stride[0] = strides[0]; m->stride[1] = strides[1];
	return CUDA_SUCCESS;
}

cudaError_t cudaGetDriverEntryPoint(const char *symbol, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *st)
{
	*fn = nullptr;
	if (!strcmp(symbol, "cuTensorMapEncodeTiled")) *fn = (void *)cuTensorMapEncodeTiled;
	if (st) *st = *fn ? cudaDriverEntryPointSuccess : 1;
	return *fn ? cudaSuccess : 1;
}
