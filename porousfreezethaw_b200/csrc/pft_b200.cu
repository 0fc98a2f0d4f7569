// pft_b200.cu -- C ABI (include/pft_b200.h) over the sm_100a kernels of pft_kernels.cuh.
//
// Host side of the Runge-Kutta-Merson integrator.  The control flow restates
// modules/RK_MPI_SAsolver_hybrid2/RK_MPI_SAsolver_hybrid2.c:215-770 of the reference for a
// device-resident state:
//   * the five RHS evaluations and all vector combinations run as five fused stage kernels;
//   * the step-size controller stays on the host because it needs glibc's pow() to reproduce the
//     reference's accepted-step sequence bit for bit (the device publishes the error max into a
//     pinned, mapped slot; the host spins on its sequence number);
//   * while the host decides, the device already evaluates K1 of the NEXT step from the
//     candidate x_new, gated on the accept flag the last CTA of stage 5 computed (the accept
//     test itself is an exact comparison, only the new step length needs pow);
//   * after a rejected attempt K1 is reused (x and t did not change), which the reference
//     recomputes needlessly (:373).
//
// This file is compiled with -fmad=false and the host compiler with -ffp-contract=off: every
// floating-point expression keeps the reference's association.
#include "../../include/pft_b200.h"
#include "pft_kernels.cuh"
#include "pft_kernels_tma.cuh"

#include <dlfcn.h>
#include <emmintrin.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <new>
#include <vector>

#include <nccl.h>

using namespace pft;

// Kernel launch.  tests/cuda_emu/ compiles this same file with g++ against a fiber-based
// emulation of the CUDA runtime (PFT_CUDA_EMU) so that the host control flow and the kernel
// indexing can be checked against the oracle on a machine without a GPU.  The emulation is
// test infrastructure: the product library is only ever built by nvcc for sm_100a.
#ifdef PFT_CUDA_EMU
#define PFT_LAUNCH(kernel, grid, block, smem, stream, ...) pft_emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define PFT_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, arg) PFT_LAUNCH(kernel, grid, block, smem, stream, arg)
#else
#define PFT_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// stage kernels: programmatic stream serialization -- the launch may begin while its predecessor in the
// stream drains; the kernels order themselves with griddepcontrol (pdl_wait / pdl_launch, pft_kernels.cuh)
#define PFT_LAUNCH_PDL(pdl_, kernel_, grid_, block_, smem_, stream_, arg_)                           \
	do {                                                                                             \
		cudaLaunchConfig_t cfg_ = {};                                                                \
		cfg_.gridDim = (grid_); cfg_.blockDim = (block_); cfg_.dynamicSmemBytes = (smem_); cfg_.stream = (stream_); \
		cudaLaunchAttribute at_[1];                                                                  \
		at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                              \
		at_[0].val.programmaticStreamSerializationAllowed = (pdl_) ? 1 : 0;                          \
		cfg_.attrs = at_; cfg_.numAttrs = 1;                                                         \
		cudaLaunchKernelEx(&cfg_, kernel_, arg_);                                                    \
	} while (0)
#endif

// ------------------------------------------------------------------------------------------
// errors

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return code;
}

#define CU(call)                                                                                     \
	do {                                                                                             \
		cudaError_t e_ = (call);                                                                     \
		if (e_ != cudaSuccess) return fail(PFT_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

extern "C" const char *pft_last_error(void) { return g_err; }
#ifdef PFT_FAST_FP
extern "C" const char *pft_version(void) { return "pft-b200 0.1 (sm_100a, float64, FAST variant: FMA contraction allowed, not bit-exact)"; }
#else
extern "C" const char *pft_version(void) { return "pft-b200 0.1 (sm_100a, float64, strict IEEE association)"; }
#endif

// ------------------------------------------------------------------------------------------
// NCCL through dlopen: single-GPU use must not require the library

struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)(void) = nullptr;
	ncclResult_t (*GroupEnd)(void) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(void)
{
	if (g_nccl.h) return 0;
	const char *names[] = {getenv("PFT_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
	void *h = nullptr;
	for (int i = 0; i < 3 && !h; i++)
		if (names[i] && *names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
	if (!h) return fail(PFT_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                               \
	*(void **)(&g_nccl.field) = dlsym(h, name);                                        \
	if (!g_nccl.field) return fail(PFT_ENCCL, "libnccl: missing symbol %s", name);
	SYM(GetUniqueId, "ncclGetUniqueId")
	SYM(CommInitRank, "ncclCommInitRank")
	SYM(CommDestroy, "ncclCommDestroy")
	SYM(Send, "ncclSend")
	SYM(Recv, "ncclRecv")
	SYM(AllReduce, "ncclAllReduce")
	*(void **)(&g_nccl.AllGather) = dlsym(h, "ncclAllGather");      // optional (absent in the CPU test stand-in)
	SYM(GroupStart, "ncclGroupStart")
	SYM(GroupEnd, "ncclGroupEnd")
	SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
	g_nccl.h = h;
	return 0;
}

#define NC(call)                                                                                     \
	do {                                                                                             \
		ncclResult_t r_ = (call);                                                                    \
		if (r_ != ncclSuccess) return fail(PFT_ENCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
	} while (0)

// ------------------------------------------------------------------------------------------

// parameter indices: apps/intertrack-hybrid-S-freezing/model.c:44-59
enum {
	P_u_star, P_L, P_xi, P_a, P_b, P_alpha, P_mu,
	P_beads_scaling, P_beads_offset_x, P_beads_offset_y, P_beads_offset_z,
	P_xi_gl, P_zeta, P_p_eps0, P_p_eps1, P_gamma,
	P_water_cp, P_ice_cp, P_glass_cp, P_water_lambda, P_ice_lambda, P_glass_lambda,
	P_water_rho, P_ice_rho, P_glass_rho, P_top_temp1, P_top_temp2, P_phase_switch_time,
	P_u_noise_amp, P_ball_radius
};

enum { A_X = 0, A_K1, A_A3, A_K3, A_K4, A_XN, A_COUNT };   // array pairs (u,p); order = pft_debug_fetch `which`

struct StageTimer { cudaEvent_t a, b; int stage; };

struct pft_solver {
	pft_case cs;
	unsigned flags;
	int device;
	Consts c;
	Geom g;
	int model;            // 0,1,2
	bool heat;
	bool share;
	bool speculate, speculate_req;
	int TX, TY, nzc;
	dim3 grid, block;
	size_t smem;
	long long field;      // doubles per field array
	double *arr[A_COUNT][2];   // [pair][u|p]
	double *gl;
	double *stage_buf;    // padded-layout staging (3*SIZE doubles)
	long long padded;     // 3*SIZE
	int N1, N2, N3;
	Ctl *ctl;
	double *red;          // device: {eps, nonfinite} of the last stage 5 (multi-GPU reduction input)
	double *red_host;     // pinned copy
	HostSlot *slot;       // pinned + mapped
	HostSlot *slot_dev;   // device alias
	unsigned long long seq;
	cudaStream_t stream;
	cudaEvent_t ev0, ev1, ev_red;
	bool k1_valid;
	double k1_t;
	ncclComm_t comm;
	bool have_comm;
	bool profile;
	bool pdl;             // stage kernels use programmatic dependent launch (off while profiling stages)
	std::vector<StageTimer> timers;
	pft_stats st;
	// one device slab: [SyncBlock][13 field arrays]; peers map it with CUDA IPC
	char *slab;
	size_t slab_bytes;
	size_t off[A_COUNT][2], off_gl;            // byte offsets of the arrays inside the slab (same on every rank)
	// peer path (multi-GPU without NCCL in the time loop)
	bool p2p;
	char *peer_base[PFT_MAX_RANKS];            // every rank's slab in this process' address space
	int peer_n3[PFT_MAX_RANKS];
	unsigned int *halo_tickets;
	unsigned long long hseq;                   // stage-launch sequence number (identical on all ranks)
	// TMA path (heat modes): one tensor map per array, box = tile + halo ring
	bool use_tma;
	bool split;           // split-phase mbarriers instead of a CTA barrier per plane
	bool rows2;           // two vertically adjacent cells per thread (merson_stage_tma2)
	CUtensorMap tmap[A_COUNT][2];
	CUtensorMap tmap_gl;
	int bx, by, as;
	size_t smem_tma[5];
	// development build (-DPFT_TRACE): device timeline records, 16 x u64 per stage launch
	unsigned long long *trace, *trace_cta;
	std::vector<int> trace_stage;
};
#ifdef PFT_TRACE
constexpr int TRACE_CAP = 4096, TRACE_REC = 16, TRACE_CTA_LAUNCHES = 12;
#endif

static double dirichlet(const pft_solver *s, double t)
{
	// equation.c:96-111
	return t < s->cs.param[P_phase_switch_time] ? s->cs.param[P_top_temp1] : s->cs.param[P_top_temp2];
}

static int env_int(const char *name, int dflt)
{
	const char *v = getenv(name);
	return (v && *v) ? atoi(v) : dflt;
}

// tile / chunk heuristics ------------------------------------------------------------------

static void choose_tile(int n1, int n2, int max_threads, int *TX, int *TY)
{
	// Cost model fitted to the B200 sweep in profiles/round1_tuning.md: wasted threads of ragged
	// tiles, the halo ring the TMA box drags along ((TX+4) x (TY+2) for TX x TY cells), and a
	// penalty for short box rows (a 26-wide tile measured 15% slower than a 50-wide one with the
	// same thread count).
	double best = 1e300;
	int btx = std::min(n1, 64), bty = std::max(3, std::min(n2, max_threads / std::max(1, btx)));
	for (int nx = 1; nx <= 64; nx++) {
		int tx = (n1 + nx - 1) / nx;
		tx += tx & 1;                         // tile origins must be even (TMA inner-offset alignment)
		if (tx > 128) continue;
		if (tx < 8 && nx > 1) break;
		for (int ty = std::min(n2, max_threads / tx); ty >= 3; ty--) {
			if (tx * ty < 2 * tx + 2 * ty) continue;
			if (tx * ty < std::min(96, n1 * n2)) break;
			double tiles = (double)((n1 + tx - 1) / tx) * ((n2 + ty - 1) / ty);
			double waste = tiles * tx * ty / ((double)n1 * n2);
			double box = (double)(tx + 4) * (ty + 2) / ((double)tx * ty);
			double rows = 1.0 + 24.0 / tx;
			double small = (tx * ty < 192) ? 1.0 + 0.2 * (192 - tx * ty) / 192.0 : 1.0;
			double flat = (ty < 5) ? 1.0 + 0.06 * (5 - ty) : 1.0;     // 3-4 row tiles measured slower
			double cost = waste * (0.6 + 0.4 * box) * rows * small * flat;
			if (cost < best) { best = cost; btx = tx; bty = ty; }
		}
	}
	// degenerate grids (a single row or column): pad the tile with inactive threads so that the
	// halo-ring duty (2*TX+2*TY cells) still fits into the CTA
	if (btx < 4) btx = 4;
	if (btx * bty < 2 * btx + 2 * bty) bty = 4;
	*TX = btx; *TY = bty;
}

// z-chunks per tile column.  Measured on B200 (profiles/round1_tuning.md, sweeps on 100x100x200,
// 200x200x50, 200x200x400, 400x400x50): throughput is flat once a launch has 2.5-7 waves of CTAs
// (1000-3000 at three resident CTAs per SM) and falls off below 7 planes per chunk (two extra planes
// are fetched per chunk) and below ~2 waves (tail: late CTAs run alone).  So: aim at 4 waves,
// never below 7 planes per chunk.
static int choose_chunks(int tiles_xy, int n3, int sms)
{
	const long target = (long)(4 * 3 * sms);
	long nzc = (target + tiles_xy / 2) / tiles_xy;
	const long most = std::max(1, n3 / 7);
	nzc = std::max(1L, std::min(nzc, most));
	return (int)nzc;
}

// ------------------------------------------------------------------------------------------

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D tensor map over one field array: (px, n2, n3+2) doubles, box (bx, by, 1)
static int make_tensor_map(pft_solver *s, CUtensorMap *m, double *base)
{
	static encode_tiled_fn encode = nullptr;
	if (!encode) {
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
		if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) return fail(PFT_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
		encode = (encode_tiled_fn)fn;
	}
	const Geom &g = s->g;
	cuuint64_t dim[3] = {(cuuint64_t)g.px, (cuuint64_t)g.n2, (cuuint64_t)(g.n3 + 2)};
	cuuint64_t strides[2] = {(cuuint64_t)g.px * sizeof(double), (cuuint64_t)g.plane * sizeof(double)};
	cuuint32_t box[3] = {(cuuint32_t)s->bx, (cuuint32_t)s->by, 1u};
	cuuint32_t estr[3] = {1u, 1u, 1u};
	CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dim, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) return fail(PFT_ECUDA, "cuTensorMapEncodeTiled failed (%d) for box %dx%d", (int)r, s->bx, s->by);
	return 0;
}

static void make_consts(pft_solver *s)
{
	const pft_case &cs = s->cs;
	const double *p = cs.param;
	Consts &c = s->c;
	// equation.c:605-612 (h3 from total_n3)
	double h1 = ((double)cs.n1) / cs.L1;
	double h2 = ((double)cs.n2) / cs.L2;
	double h3 = ((double)cs.total_n3) / cs.L3;
	c.h1_2 = h1 * h1; c.h1_d2 = 0.5 * h1;
	c.h2_2 = h2 * h2; c.h2_d2 = 0.5 * h2;
	c.h3_2 = h3 * h3; c.h3_d2 = 0.5 * h3;
	// equation.c:442-447
	c.xi_2_inv_a = p[P_a] / (p[P_xi] * p[P_xi]);
	double xi_inv_b_sqrt_a2 = p[P_b] * sqrt(0.5 * p[P_a]) / p[P_xi];
	c.eps2_3 = 3.0 / ((p[P_p_eps1] - p[P_p_eps0]) * (p[P_p_eps1] - p[P_p_eps0]));
	c.eps3_2 = 2.0 / ((p[P_p_eps1] - p[P_p_eps0]) * (p[P_p_eps1] - p[P_p_eps0]) * (p[P_p_eps1] - p[P_p_eps0]));
	// leading constant products of equation.c:370 and :387 (left-to-right, as C evaluates them)
	c.bam = p[P_b] * p[P_alpha] * p[P_mu];
	c.sam = xi_inv_b_sqrt_a2 * p[P_alpha] * p[P_mu];
	c.p_eps0 = p[P_p_eps0]; c.p_eps1 = p[P_p_eps1];
	c.alpha = p[P_alpha]; c.zeta = p[P_zeta]; c.L = p[P_L]; c.u_star = p[P_u_star]; c.gamma = p[P_gamma];
	c.g_rho = p[P_glass_rho]; c.i_rho = p[P_ice_rho]; c.w_rho = p[P_water_rho];
	c.g_cp = p[P_glass_cp]; c.i_cp = p[P_ice_cp]; c.w_cp = p[P_water_cp];
	c.g_lam = p[P_glass_lambda]; c.i_lam = p[P_ice_lambda]; c.w_lam = p[P_water_lambda];
}

template <int STAGE, int MODEL, bool HEAT, bool SHARE>
static cudaError_t launch_t(pft_solver *s, const KParams &kp)
{
	auto fn = merson_stage<STAGE, MODEL, HEAT, SHARE>;
	static bool attr_set_dev[64] = {};        // the attribute is per device (and per template instance)
	bool &attr_set = attr_set_dev[s->device & 63];
	if (!attr_set) {
		cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		if (e != cudaSuccess) return e;
		attr_set = true;
	}
	PFT_LAUNCH_PDL(s->pdl, fn, s->grid, s->block, s->smem, s->stream, kp);
	return cudaGetLastError();
}

template <int STAGE>
static cudaError_t launch_stage_t(pft_solver *s, const KParams &kp)
{
	if (s->model == 2) return s->share ? launch_t<STAGE, 2, true, true>(s, kp) : launch_t<STAGE, 2, true, false>(s, kp);
	if (s->model == 0) {
		if (s->heat) return s->share ? launch_t<STAGE, 0, true, true>(s, kp) : launch_t<STAGE, 0, true, false>(s, kp);
		return launch_t<STAGE, 0, false, false>(s, kp);
	}
	if (s->heat) return s->share ? launch_t<STAGE, 1, true, true>(s, kp) : launch_t<STAGE, 1, true, false>(s, kp);
	return launch_t<STAGE, 1, false, false>(s, kp);
}

template <int STAGE, int MODEL, bool SHARE, bool SPLIT>
static cudaError_t launch_tma_t(pft_solver *s, const TmaParams &kp)
{
	auto fn = merson_stage_tma<STAGE, MODEL, SHARE, SPLIT>;
	static bool attr_set_dev[64] = {};        // the attribute is per device (and per template instance)
	bool &attr_set = attr_set_dev[s->device & 63];
	if (!attr_set) {
		cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
		if (e != cudaSuccess) return e;
		attr_set = true;
	}
	PFT_LAUNCH_PDL(s->pdl, fn, s->grid, s->block, s->smem_tma[STAGE - 1], s->stream, kp);
	return cudaGetLastError();
}

template <int STAGE, int MODEL>
static cudaError_t launch_tma2_t(pft_solver *s, const TmaParams &kp)
{
	auto fn = merson_stage_tma2<STAGE, MODEL>;
	static bool attr_set_dev[64] = {};        // the attribute is per device (and per template instance)
	bool &attr_set = attr_set_dev[s->device & 63];
	if (!attr_set) {
		cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
		if (e != cudaSuccess) return e;
		attr_set = true;
	}
	PFT_LAUNCH_PDL(s->pdl, fn, s->grid, s->block, s->smem_tma[STAGE - 1], s->stream, kp);
	return cudaGetLastError();
}

template <int STAGE>
static cudaError_t launch_stage_tma_t(pft_solver *s, const TmaParams &kp)
{
	if (s->rows2) {
		if (s->model == 2) return launch_tma2_t<STAGE, 2>(s, kp);
		if (s->model == 0) return launch_tma2_t<STAGE, 0>(s, kp);
		return launch_tma2_t<STAGE, 1>(s, kp);
	}
	if (s->share) {      // A/B variant, CTA barriers only
		if (s->model == 2) return launch_tma_t<STAGE, 2, true, false>(s, kp);
		if (s->model == 0) return launch_tma_t<STAGE, 0, true, false>(s, kp);
		return launch_tma_t<STAGE, 1, true, false>(s, kp);
	}
	if (s->model == 2) return s->split ? launch_tma_t<STAGE, 2, false, true>(s, kp) : launch_tma_t<STAGE, 2, false, false>(s, kp);
	if (s->model == 0) return s->split ? launch_tma_t<STAGE, 0, false, true>(s, kp) : launch_tma_t<STAGE, 0, false, false>(s, kp);
	return s->split ? launch_tma_t<STAGE, 1, false, true>(s, kp) : launch_tma_t<STAGE, 1, false, false>(s, kp);
}

// peer path: where this launch pushes its boundary planes and whom it signals
static void fill_peer_args(pft_solver *s, PeerArgs *pr, int out)
{
	const int r = s->cs.rank, R = s->cs.nranks;
	const size_t plane_b = (size_t)s->g.plane * sizeof(double);
	pr->rank = r; pr->nranks = R;
	pr->nzc = s->nzc;
	pr->hseq = ++s->hseq;
	pr->my_flags = ((SyncBlock *)s->slab)->halo_flag;
	pr->tickets = s->halo_tickets;
	pr->wait_lo = (r > 0); pr->wait_hi = (r < R - 1);
	for (int q = 0; q < R; q++) pr->ranks[q] = (SyncBlock *)s->peer_base[q];
	if (r > 0) {        // the lower neighbour's UPPER ghost plane (index n3'+1) and its halo_flag[1]
		char *pb = s->peer_base[r - 1];
		const size_t gp = (size_t)(s->peer_n3[r - 1] + 1) * plane_b;
		pr->lo_p = (double *)(pb + s->off[out][1] + gp);
		pr->lo_u = s->heat ? (double *)(pb + s->off[out][0] + gp) : nullptr;
		pr->lo_flag = &((SyncBlock *)pb)->halo_flag[1];
	}
	if (r < R - 1) {    // the upper neighbour's LOWER ghost plane (index 0) and its halo_flag[0]
		char *pb = s->peer_base[r + 1];
		pr->hi_p = (double *)(pb + s->off[out][1]);
		pr->hi_u = s->heat ? (double *)(pb + s->off[out][0]) : nullptr;
		pr->hi_flag = &((SyncBlock *)pb)->halo_flag[0];
	}
#ifdef PFT_TRACE
	// timing experiment only (wrong results): keep the flags, drop the plane pushes
	static const int nopush = env_int("PFT_DEBUG_NOPUSH", 0);
	if (nopush) pr->lo_p = pr->lo_u = pr->hi_p = pr->hi_u = nullptr;
#endif
}

// stage launcher: fills the argument block for one Merson stage and (multi-GPU) exchanges the
// z-boundary planes of what the stage produced
static int exchange_pair(pft_solver *s, double *u, double *p);

static int launch_stage(pft_solver *s, int stage, double t_stage, double ha, double hb,
                        int xin, int out, const int *cond, const pft_system *sys, double h_abs, double eps_scale)
{
	KParams kp;
	kp.c = s->c;
	kp.g = s->g;
	StageArgs &a = kp.a;
	memset(&a, 0, sizeof a);
	a.xu = s->arr[xin][0]; a.xp = s->arr[xin][1]; a.gl = s->gl;
	a.k1u = s->arr[A_K1][0]; a.k1p = s->arr[A_K1][1];
	a.k3u = s->arr[A_K3][0]; a.k3p = s->arr[A_K3][1];
	a.k4u = s->arr[A_K4][0]; a.k4p = s->arr[A_K4][1];
	a.au = s->arr[A_A3][0]; a.ap = s->arr[A_A3][1];
	a.ou = s->arr[out][0]; a.op = s->arr[out][1];
	a.ha = ha; a.hb = hb;
	a.dir_u = dirichlet(s, t_stage);
	a.cond = cond;
	if (s->p2p) fill_peer_args(s, &a.pr, out);
	if (stage == 5) {
		a.delta = sys->delta; a.h_abs = h_abs; a.h_min = sys->h_min;
		a.eps_scale = eps_scale; a.scale_eps = sys->delta_local ? 1 : 0;
		a.handle_nan = sys->handle_nan;
		a.ctl = s->ctl; a.slot = s->slot_dev; a.seq = ++s->seq;
		a.nblocks = s->grid.x * s->grid.y * s->grid.z;
	}
#ifdef PFT_TRACE
	if (!s->trace) {
		if (cudaMalloc(&s->trace, sizeof(unsigned long long) * TRACE_CAP * TRACE_REC) != cudaSuccess) return fail(PFT_ENOMEM, "trace buffer");
		cudaMemsetAsync(s->trace, 0, sizeof(unsigned long long) * TRACE_CAP * TRACE_REC, s->stream);
	}
	if (!s->trace_cta) {
		if (cudaMalloc(&s->trace_cta, sizeof(unsigned long long) * TRACE_CTA_LAUNCHES * 4096) != cudaSuccess) return fail(PFT_ENOMEM, "trace buffer");
		cudaMemsetAsync(s->trace_cta, 0, sizeof(unsigned long long) * TRACE_CTA_LAUNCHES * 4096, s->stream);
	}
	a.trace = ((int)s->trace_stage.size() < TRACE_CAP) ? s->trace + s->trace_stage.size() * TRACE_REC : nullptr;
	a.trace_cta = ((int)s->trace_stage.size() < TRACE_CTA_LAUNCHES) ? s->trace_cta + s->trace_stage.size() * 4096 : nullptr;
	if (a.trace) s->trace_stage.push_back(stage);
#endif
	StageTimer tm{};
	if (s->profile) {
		cudaEventCreate(&tm.a); cudaEventCreate(&tm.b); tm.stage = stage;
		cudaEventRecord(tm.a, s->stream);
	}
	cudaError_t e;
	if (s->use_tma) {
		TmaParams tp;
		tp.c = kp.c; tp.g = kp.g; tp.a = kp.a;
		tp.bx = s->bx; tp.by = s->by; tp.as = s->as;
		// array order of a stage: p-type arrays, u-type arrays, gl (pft_kernels_tma.cuh, stage_np)
		const int src[5][4] = {{xin, -1, -1, -1}, {xin, A_K1, -1, -1}, {A_A3, -1, -1, -1}, {xin, A_K1, A_K3, -1}, {xin, A_K1, A_K3, A_K4}};
		const int np = stage_np(stage);
		for (int m = 0; m < np; m++) {
			tp.maps[m] = s->tmap[src[stage - 1][m]][1];
			tp.maps[np + m] = s->tmap[src[stage - 1][m]][0];
		}
		tp.maps[2 * np] = s->tmap_gl;
		switch (stage) {
		case 1: e = launch_stage_tma_t<1>(s, tp); break;
		case 2: e = launch_stage_tma_t<2>(s, tp); break;
		case 3: e = launch_stage_tma_t<3>(s, tp); break;
		case 4: e = launch_stage_tma_t<4>(s, tp); break;
		default: e = launch_stage_tma_t<5>(s, tp); break;
		}
	} else
	switch (stage) {
	case 1: e = launch_stage_t<1>(s, kp); break;
	case 2: e = launch_stage_t<2>(s, kp); break;
	case 3: e = launch_stage_t<3>(s, kp); break;
	case 4: e = launch_stage_t<4>(s, kp); break;
	default: e = launch_stage_t<5>(s, kp); break;
	}
	if (e != cudaSuccess) return fail(PFT_ECUDA, "stage %d launch: %s", stage, cudaGetErrorString(e));
	if (s->profile) { cudaEventRecord(tm.b, s->stream); s->timers.push_back(tm); }
	s->st.kernel_launches++;
	s->st.stage_launches[stage - 1]++;
	if (s->cs.nranks > 1 && !s->p2p) {
		int rc = exchange_pair(s, s->heat ? s->arr[out][0] : nullptr, s->arr[out][1]);
		if (rc) return rc;
	}
	return 0;
}

static void drain_timers(pft_solver *s)
{
	for (auto &tm : s->timers) {
		float ms = 0.f;
		cudaEventSynchronize(tm.b);
		cudaEventElapsedTime(&ms, tm.a, tm.b);
		s->st.stage_ms[tm.stage - 1] += ms;
		cudaEventDestroy(tm.a); cudaEventDestroy(tm.b);
	}
	s->timers.clear();
}

// ------------------------------------------------------------------------------------------
// multi-GPU: one z-boundary plane per direction and field (the 7-point stencil reads +-1 only;
// the reference ships 2 planes of all 3 variables, equation.c:304-322)

static int exchange_pair(pft_solver *s, double *u, double *p)
{
	if (s->cs.nranks <= 1) return 0;
	if (!s->have_comm) return fail(PFT_ESTATE, "nranks > 1 but pft_comm_init was not called");
	const size_t n = (size_t)s->g.plane;
	const int r = s->cs.rank, R = s->cs.nranks, n3 = s->g.n3;
	double *f[2] = {u, p};
	NC(g_nccl.GroupStart());
	for (int q = 0; q < 2; q++) {
		if (!f[q]) continue;
		if (r > 0) {
			NC(g_nccl.Send(f[q] + n * 1, n, ncclDouble, r - 1, s->comm, s->stream));
			NC(g_nccl.Recv(f[q], n, ncclDouble, r - 1, s->comm, s->stream));
		}
		if (r < R - 1) {
			NC(g_nccl.Send(f[q] + n * (size_t)n3, n, ncclDouble, r + 1, s->comm, s->stream));
			NC(g_nccl.Recv(f[q] + n * (size_t)(n3 + 1), n, ncclDouble, r + 1, s->comm, s->stream));
		}
	}
	NC(g_nccl.GroupEnd());
	return 0;
}

extern "C" int pft_comm_unique_id(void *id128)
{
	if (!id128) return fail(PFT_EINVAL, "null id buffer");
	int rc = nccl_load();
	if (rc) return rc;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
	ncclUniqueId id;
	NC(g_nccl.GetUniqueId(&id));
	memcpy(id128, &id, sizeof id);
	return 0;
}

// slab depth of rank r: apps/intertrack-hybrid-S-freezing/intertrack.c:1780-1787
static int slab_n3(int total_n3, int nranks, int r)
{
	return total_n3 / nranks + (r < total_n3 % nranks ? 1 : 0);
}

// Peer path: every rank exports its slab with CUDA IPC, the handles travel once through the NCCL
// communicator, and from then on the time loop contains no collective call at all: boundary planes
// are stored straight into the neighbours' ghost planes by the stage kernels, flags and the
// per-slab error maxima travel as peer stores (pft_kernels.cuh, "peer-memory halo exchange").
static int setup_p2p(pft_solver *s)
{
	const int R = s->cs.nranks, r = s->cs.rank;
	if (R > PFT_MAX_RANKS) return fail(PFT_EINVAL, "peer path supports at most %d ranks", PFT_MAX_RANKS);
	struct Export { cudaIpcMemHandle_t h; int n3; int ok; unsigned long long bytes; };
	Export mine;
	memset(&mine, 0, sizeof mine);
	mine.ok = (cudaIpcGetMemHandle(&mine.h, s->slab) == cudaSuccess);
	if (!mine.ok) cudaGetLastError();
	mine.n3 = s->cs.n3;
	mine.bytes = s->slab_bytes;
	Export *d = nullptr;
	std::vector<Export> all(R);
	CU(cudaMalloc(&d, sizeof(Export) * R));
	CU(cudaMemcpyAsync(d + r, &mine, sizeof mine, cudaMemcpyHostToDevice, s->stream));
	NC(g_nccl.AllGather(d + r, d, sizeof(Export), ncclInt8, s->comm, s->stream));
	CU(cudaMemcpyAsync(all.data(), d, sizeof(Export) * R, cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	cudaFree(d);
	bool ok = true;
	for (int q = 0; q < R; q++) {
		ok = ok && all[q].ok && all[q].bytes == s->slab_bytes;
		if (all[q].n3 != slab_n3(s->cs.total_n3, R, q)) return fail(PFT_EINVAL, "rank %d holds %d planes, expected %d", q, all[q].n3, slab_n3(s->cs.total_n3, R, q));
		s->peer_n3[q] = all[q].n3;
	}
	if (ok) {
		for (int q = 0; q < R && ok; q++) {
			if (q == r) { s->peer_base[q] = s->slab; continue; }
			void *p = nullptr;
			cudaError_t e = cudaIpcOpenMemHandle(&p, all[q].h, cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) { cudaGetLastError(); ok = false; break; }
			s->peer_base[q] = (char *)p;
		}
	}
	// every rank must take the same path: agree through the communicator
	double flag = ok ? 1.0 : 0.0, *df = nullptr;
	CU(cudaMalloc(&df, sizeof(double)));
	CU(cudaMemcpyAsync(df, &flag, sizeof flag, cudaMemcpyHostToDevice, s->stream));
	NC(g_nccl.AllReduce(df, df, 1, ncclDouble, ncclMin, s->comm, s->stream));
	CU(cudaMemcpyAsync(&flag, df, sizeof flag, cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	cudaFree(df);
	s->p2p = (flag == 1.0);
	if (!s->p2p)
		for (int q = 0; q < R; q++) {
			if (s->peer_base[q] && s->peer_base[q] != s->slab) cudaIpcCloseMemHandle(s->peer_base[q]);
			s->peer_base[q] = nullptr;
		}
	s->speculate = s->speculate_req && (s->cs.nranks == 1 || s->p2p);
	return 0;
}

extern "C" int pft_comm_init(pft_solver *s, const void *id128)
{
	if (!s || !id128) return fail(PFT_EINVAL, "null argument");
	if (s->cs.nranks <= 1) return 0;
	int rc = nccl_load();
	if (rc) return rc;
	CU(cudaSetDevice(s->device));
	ncclUniqueId id;
	memcpy(&id, id128, sizeof id);
	NC(g_nccl.CommInitRank(&s->comm, s->cs.nranks, id, s->cs.rank));
	s->have_comm = true;
	const char *mode = getenv("PFT_COMM");       // "p2p" (default) | "nccl"
	if (!(mode && !strcmp(mode, "nccl")) && g_nccl.AllGather) {
		rc = setup_p2p(s);
		if (rc) return rc;
	}
	return 0;
}

// ------------------------------------------------------------------------------------------

// inside pft_create, once the solver object exists: release it on the way out of a failed CUDA call
#define CUS(call)                                                                                    \
	do {                                                                                             \
		cudaError_t e_ = (call);                                                                     \
		if (e_ != cudaSuccess) {                                                                     \
			pft_destroy(s);                                                                          \
			return fail(PFT_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
		}                                                                                            \
	} while (0)

extern "C" int pft_create(const pft_case *cs, int device, unsigned flags, pft_solver **out)
{
	if (!cs || !out) return fail(PFT_EINVAL, "null argument");
	*out = nullptr;
	if (cs->n1 < 1 || cs->n2 < 1 || cs->n3 < 1 || cs->total_n3 < cs->n3) return fail(PFT_EINVAL, "bad grid %dx%dx%d (total_n3 %d)", cs->n1, cs->n2, cs->n3, cs->total_n3);
	if (cs->nranks < 1 || cs->rank < 0 || cs->rank >= cs->nranks) return fail(PFT_EINVAL, "bad rank %d of %d", cs->rank, cs->nranks);
	switch (cs->calc_mode) { case 0: case 1: case 2: case 10: case 11: break; default: return fail(PFT_EINVAL, "invalid calc_mode %d", cs->calc_mode); }
	if (cs->param[P_u_noise_amp] != 0.0)
		return fail(PFT_EINVAL, "u_noise_amp != 0 is not supported (the reference seeds its noise field from time(), intertrack.c:1278)");
	if ((long long)3 * (cs->n1 + 4) * (cs->n2 + 4) * (long long)(cs->n3 + 4) > 2147483647LL)
		return fail(PFT_EINVAL, "block too large for the reference's int-indexed chunk table");

	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev < 1) return fail(PFT_ENODEVICE, "no CUDA device (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
	if (device < 0 || device >= ndev) return fail(PFT_ENODEVICE, "device %d out of range (%d present)", device, ndev);
	CU(cudaSetDevice(device));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) return fail(PFT_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

	pft_solver *s = new (std::nothrow) pft_solver();
	if (!s) return fail(PFT_ENOMEM, "out of host memory");
	s->cs = *cs;
	s->flags = flags;
	s->device = device;
	s->model = (cs->calc_mode == 2) ? 2 : (cs->calc_mode % 10);
	s->heat = !(cs->calc_mode == 10 || cs->calc_mode == 11);
	// Sharing face fluxes through shared memory saves ~25% of the fp64 work but costs a second CTA
	// barrier per plane; measured on B200 it loses (profiles/round1_tuning.md), so it is opt-in.
	s->share = s->heat && !(flags & PFT_FLAG_NO_FLUX_SHARING) && ((flags & PFT_FLAG_FLUX_SHARING) || env_int("PFT_SHARE", 0));
	s->speculate_req = !(flags & PFT_FLAG_NO_SPECULATION) && !env_int("PFT_NO_SPECULATION", 0);
	s->speculate = s->speculate_req && cs->nranks == 1;      // N > 1: enabled by the peer path (setup_p2p)
	s->profile = env_int("PFT_PROFILE_STAGES", 0) != 0;
	s->pdl = !s->profile && env_int("PFT_NO_PDL", 0) == 0;
	make_consts(s);

	Geom &g = s->g;
	g.n1 = cs->n1; g.n2 = cs->n2; g.n3 = cs->n3;
	g.px = (cs->n1 + 3) & ~3;
	g.plane = (long long)g.px * cs->n2;
	g.front = (cs->rank == 0);
	g.rear = (cs->rank == cs->nranks - 1);
	// every rank sizes its arrays for the deepest slab so that array offsets inside the slab are
	// the same on all ranks (peers address each other's arrays by offset)
	s->field = g.plane * ((cs->total_n3 + cs->nranks - 1) / cs->nranks + 2);
	s->N1 = cs->n1 + 2 * PFT_BT; s->N2 = cs->n2 + 2 * PFT_BT; s->N3 = cs->n3 + 2 * PFT_BT;
	s->padded = (long long)PFT_NVAR * s->N1 * s->N2 * s->N3;

	s->use_tma = s->heat && !(flags & PFT_FLAG_LEGACY_LOADS) && !env_int("PFT_LEGACY", 0);
	s->rows2 = s->use_tma && !s->share && env_int("PFT_ROWS2", 0) != 0;    // measured slower than one cell per thread (profiles/round1_tuning.md)
	const int rpt = s->rows2 ? 2 : 1;        // cell rows per thread
	int max_threads = env_int("PFT_THREADS", s->rows2 ? 160 : 256);     // two or three CTAs per SM overlap each other's barrier waits
	if (max_threads < 32) max_threads = 32;
	if (max_threads > 512) max_threads = 512;
	if (s->use_tma && max_threads > 256) max_threads = 256;      // __launch_bounds__ of the TMA kernels
	choose_tile(cs->n1, (cs->n2 + rpt - 1) / rpt, max_threads, &s->TX, &s->TY);
	if (env_int("PFT_TX", 0) > 0) s->TX = env_int("PFT_TX", 0);
	if (env_int("PFT_TY", 0) > 0) s->TY = env_int("PFT_TY", 0);
	if (s->TX & 1) s->TX++;                  // tile origins must be even (TMA inner-offset alignment)
	if (s->TX * s->TY > (s->use_tma ? 256 : 512) || rpt * s->TX * s->TY < 2 * s->TX + 2 * rpt * s->TY) {
		int tx_ = s->TX, ty_ = s->TY;
		delete s;
		return fail(PFT_EINVAL, "tile %dx%d unusable (needs TX*TY <= 512 and enough threads for the halo ring)", tx_, ty_);
	}
	const int tile_rows = rpt * s->TY;
	int tiles_xy = ((cs->n1 + s->TX - 1) / s->TX) * ((cs->n2 + tile_rows - 1) / tile_rows);
	s->nzc = choose_chunks(tiles_xy, cs->n3, prop.multiProcessorCount);
	if (env_int("PFT_ZCHUNKS", 0) > 0) s->nzc = std::min(cs->n3, env_int("PFT_ZCHUNKS", 0));
	g.zc = (cs->n3 + s->nzc - 1) / s->nzc;
	s->nzc = (cs->n3 + g.zc - 1) / g.zc;
	s->grid = dim3((cs->n1 + s->TX - 1) / s->TX, (cs->n2 + tile_rows - 1) / tile_rows, s->nzc);
	s->block = dim3(s->TX, s->TY, 1);
	{
		int NF = s->heat ? 3 : 1;
		size_t ts = (size_t)(s->TX + 2) * (s->TY + 2);
		size_t fl = s->share ? (size_t)s->TY * (s->TX + 1) + (size_t)(s->TY + 1) * s->TX : 0;
		s->smem = (2 * NF * ts + fl) * sizeof(double);
	}

	CUS(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
	CUS(cudaEventCreate(&s->ev0)); CU(cudaEventCreate(&s->ev1));
	CUS(cudaEventCreateWithFlags(&s->ev_red, cudaEventDisableTiming));
	const size_t fb = ((size_t)s->field * sizeof(double) + 255) & ~(size_t)255;
	{
		// layout: identical on every rank of a run with equal slab depth; offsets are exchanged anyway
		size_t off = (sizeof(SyncBlock) + 255) & ~(size_t)255;
		for (int a = 0; a < A_COUNT; a++)
			for (int q = 0; q < 2; q++) {
				// calc_mode 10/11: du/dt == 0 (equation.c:708), so u is static: no K/aux arrays for it,
				// and the candidate state shares the one temperature array
				if (!s->heat && q == 0 && a != A_X) { s->off[a][q] = 0; continue; }
				s->off[a][q] = off; off += fb;
			}
		s->off_gl = off; off += fb;
		s->slab_bytes = off;
		e = cudaMalloc(&s->slab, s->slab_bytes);
		if (e != cudaSuccess) { s->slab = nullptr; pft_destroy(s); return fail(PFT_ENOMEM, "cudaMalloc(%zu): %s", s->slab_bytes, cudaGetErrorString(e)); }
		CUS(cudaMemsetAsync(s->slab, 0, s->slab_bytes, s->stream));
		for (int a = 0; a < A_COUNT; a++)
			for (int q = 0; q < 2; q++) s->arr[a][q] = s->off[a][q] ? (double *)(s->slab + s->off[a][q]) : nullptr;
		s->gl = (double *)(s->slab + s->off_gl);
		if (!s->heat) { s->arr[A_XN][0] = s->arr[A_X][0]; s->off[A_XN][0] = s->off[A_X][0]; }
	}
	CUS(cudaMalloc(&s->halo_tickets, 2 * sizeof(unsigned int)));
	CUS(cudaMemsetAsync(s->halo_tickets, 0, 2 * sizeof(unsigned int), s->stream));

	s->split = s->use_tma && !s->share && env_int("PFT_SPLIT", 0) != 0;
	if (s->use_tma) {
		s->bx = s->TX + 4;                   // box starts at x = i0-2 (16-byte aligned inner offset), TX is even
		s->by = (s->rows2 ? 2 : 1) * s->TY + 2;
		s->as = (s->bx * s->by + 15) & ~15;  // each array buffer starts 128-byte aligned
		size_t fl = s->share ? (size_t)s->TY * (s->TX + 1) + (size_t)(s->TY + 1) * s->TX : 0;
		size_t worst = 0;
		for (int st = 1; st <= 5; st++) {
			int na = 2 * stage_np(st) + 1;
			bool form = !(st == 1 || st == 3);
			s->smem_tma[st - 1] = ((size_t)TMA_STAGES * na * s->as + (form ? (s->split && !s->rows2 ? 6 : 4) * (size_t)s->as : 0) + fl) * sizeof(double);
			worst = std::max(worst, s->smem_tma[st - 1]);
		}
		if (worst > 225 * 1024) { pft_destroy(s); return fail(PFT_EINVAL, "tile %dx%d needs %zu bytes of shared memory", s->TX, s->TY, worst); }
		for (int a = 0; a < A_COUNT; a++)
			for (int q = 0; q < 2; q++) {
				int rc = make_tensor_map(s, &s->tmap[a][q], s->arr[a][q]);
				if (rc) { pft_destroy(s); return rc; }
			}
		int rc = make_tensor_map(s, &s->tmap_gl, s->gl);
		if (rc) { pft_destroy(s); return rc; }
	}
	CUS(cudaMalloc(&s->ctl, sizeof(Ctl)));
	CUS(cudaMemsetAsync(s->ctl, 0, sizeof(Ctl), s->stream));
	CUS(cudaMalloc(&s->red, 2 * sizeof(double)));
	CUS(cudaHostAlloc(&s->red_host, 2 * sizeof(double), cudaHostAllocDefault));
	CUS(cudaHostAlloc(&s->slot, sizeof(HostSlot), cudaHostAllocMapped));
	memset(s->slot, 0, sizeof(HostSlot));
	CUS(cudaHostGetDevicePointer(&s->slot_dev, s->slot, 0));
	CUS(cudaStreamSynchronize(s->stream));

	memset(&s->st, 0, sizeof s->st);
	s->st.tile_x = s->TX; s->st.tile_y = s->TY; s->st.z_chunk = g.zc; s->st.threads = s->TX * s->TY;
	s->st.grid_blocks = s->grid.x * s->grid.y * s->grid.z;
	s->st.bytes_per_field = (long)fb;
	*out = s;
	return 0;
}

extern "C" int pft_destroy(pft_solver *s)
{
	if (!s) return 0;
	cudaSetDevice(s->device);
	if (s->stream) cudaStreamSynchronize(s->stream);
	drain_timers(s);
	if (s->have_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
	for (int r = 0; r < PFT_MAX_RANKS; r++)
		if (s->peer_base[r] && s->peer_base[r] != s->slab) cudaIpcCloseMemHandle(s->peer_base[r]);
	if (s->slab) cudaFree(s->slab);
	if (s->halo_tickets) cudaFree(s->halo_tickets);
	if (s->trace) cudaFree(s->trace);
	if (s->trace_cta) cudaFree(s->trace_cta);
	if (s->stage_buf) cudaFree(s->stage_buf);
	if (s->ctl) cudaFree(s->ctl);
	if (s->red) cudaFree(s->red);
	if (s->red_host) cudaFreeHost(s->red_host);
	if (s->slot) cudaFreeHost(s->slot);
	if (s->ev0) cudaEventDestroy(s->ev0);
	if (s->ev1) cudaEventDestroy(s->ev1);
	if (s->ev_red) cudaEventDestroy(s->ev_red);
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
	return 0;
}

// ------------------------------------------------------------------------------------------
// data movement

static int ensure_stage(pft_solver *s)
{
	if (!s->stage_buf) {
		cudaError_t e = cudaMalloc(&s->stage_buf, (size_t)s->padded * sizeof(double));
		if (e != cudaSuccess) return fail(PFT_ENOMEM, "cudaMalloc(staging %lld doubles): %s", s->padded, cudaGetErrorString(e));
	}
	return 0;
}

static void repack(pft_solver *s, double *padded_var, double *compact, int dir)
{
	const Geom &g = s->g;
	long long total = (long long)g.n1 * g.n2 * g.n3;
	int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
	PFT_LAUNCH(repack_kernel, dim3(blocks), dim3(256), 0, s->stream, padded_var, compact, g.n1, g.n2, g.n3, g.px, g.plane,
	           s->N1, (long long)s->N1 * s->N2, PFT_BT, dir);
	s->st.kernel_launches++;
}

static int refresh_halos(pft_solver *s)
{
	if (s->cs.nranks <= 1) return 0;
	int rc = exchange_pair(s, s->arr[A_X][0], s->arr[A_X][1]);
	if (rc) return rc;
	return exchange_pair(s, s->gl, nullptr);
}

extern "C" int pft_upload(pft_solver *s, const double *x)
{
	if (!s || !x) return fail(PFT_EINVAL, "null argument");
	CU(cudaSetDevice(s->device));
	int rc = ensure_stage(s);
	if (rc) return rc;
	const long long SIZE = s->padded / PFT_NVAR;
	CU(cudaMemcpyAsync(s->stage_buf, x, (size_t)s->padded * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	repack(s, s->stage_buf, s->arr[A_X][0], 0);
	repack(s, s->stage_buf + SIZE, s->arr[A_X][1], 0);
	repack(s, s->stage_buf + 2 * SIZE, s->gl, 0);
	CU(cudaGetLastError());
	rc = refresh_halos(s);
	if (rc) return rc;
	CU(cudaStreamSynchronize(s->stream));
	s->k1_valid = false;
	return 0;
}

extern "C" int pft_download(pft_solver *s, double *x)
{
	if (!s || !x) return fail(PFT_EINVAL, "null argument");
	CU(cudaSetDevice(s->device));
	int rc = ensure_stage(s);
	if (rc) return rc;
	const long long SIZE = s->padded / PFT_NVAR;
	// the staging image still holds the ghost shell of the last upload (or of this host array)
	CU(cudaMemcpyAsync(s->stage_buf, x, (size_t)s->padded * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	repack(s, s->stage_buf, s->arr[A_X][0], 1);
	repack(s, s->stage_buf + SIZE, s->arr[A_X][1], 1);
	repack(s, s->stage_buf + 2 * SIZE, s->gl, 1);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(x, s->stage_buf, (size_t)s->padded * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return 0;
}

static int dense_io(pft_solver *s, double *host, double *compact, int dir)
{
	const Geom &g = s->g;
	long long total = (long long)g.n1 * g.n2 * g.n3;
	int rc = ensure_stage(s);
	if (rc) return rc;
	int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
	if (dir == 0) {
		CU(cudaMemcpyAsync(s->stage_buf, host, (size_t)total * sizeof(double), cudaMemcpyHostToDevice, s->stream));
		PFT_LAUNCH(dense_kernel, dim3(blocks), dim3(256), 0, s->stream, s->stage_buf, compact, g.n1, g.n2, g.n3, g.px, g.plane, 0);
	} else {
		PFT_LAUNCH(dense_kernel, dim3(blocks), dim3(256), 0, s->stream, s->stage_buf, compact, g.n1, g.n2, g.n3, g.px, g.plane, 1);
		CU(cudaMemcpyAsync(host, s->stage_buf, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	}
	s->st.kernel_launches++;
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(s->stream));
	return 0;
}

extern "C" int pft_upload_interior(pft_solver *s, const double *u, const double *p, const double *gl)
{
	if (!s) return fail(PFT_EINVAL, "null solver");
	CU(cudaSetDevice(s->device));
	int rc;
	if (u && (rc = dense_io(s, const_cast<double *>(u), s->arr[A_X][0], 0))) return rc;
	if (p && (rc = dense_io(s, const_cast<double *>(p), s->arr[A_X][1], 0))) return rc;
	if (gl && (rc = dense_io(s, const_cast<double *>(gl), s->gl, 0))) return rc;
	rc = refresh_halos(s);
	if (rc) return rc;
	CU(cudaStreamSynchronize(s->stream));
	s->k1_valid = false;
	return 0;
}

extern "C" int pft_download_interior(pft_solver *s, double *u, double *p, double *gl)
{
	if (!s) return fail(PFT_EINVAL, "null solver");
	CU(cudaSetDevice(s->device));
	int rc;
	if (u && (rc = dense_io(s, u, s->arr[A_X][0], 1))) return rc;
	if (p && (rc = dense_io(s, p, s->arr[A_X][1], 1))) return rc;
	if (gl && (rc = dense_io(s, gl, s->gl, 1))) return rc;
	return 0;
}

extern "C" int pft_debug_fetch(pft_solver *s, int which, double *u, double *p)
{
	if (!s || which < 0 || which >= A_COUNT) return fail(PFT_EINVAL, "bad array id %d", which);
	CU(cudaSetDevice(s->device));
	int rc;
	if (u) {
		if (s->arr[which][0]) { if ((rc = dense_io(s, u, s->arr[which][0], 1))) return rc; }
		else memset(u, 0, sizeof(double) * (size_t)s->g.n1 * s->g.n2 * s->g.n3);     // K_u == 0 in calc_mode 10/11
	}
	if (p && (rc = dense_io(s, p, s->arr[which][1], 1))) return rc;
	return 0;
}

extern "C" int pft_debug_trace(pft_solver *s, unsigned long long *records, int *stages, int max_records)
{
	if (!s || !records || !stages) return fail(PFT_EINVAL, "null argument");
#ifdef PFT_TRACE
	CU(cudaSetDevice(s->device));
	CU(cudaStreamSynchronize(s->stream));
	const int n = std::min((int)s->trace_stage.size(), max_records);
	if (n > 0) CU(cudaMemcpy(records, s->trace, sizeof(unsigned long long) * TRACE_REC * n, cudaMemcpyDeviceToHost));
	for (int q = 0; q < n; q++) stages[q] = s->trace_stage[q];
	s->trace_stage.clear();
	if (s->trace) CU(cudaMemset(s->trace, 0, sizeof(unsigned long long) * TRACE_CAP * TRACE_REC));
	return n;
#else
	(void)max_records;
	return fail(PFT_EINVAL, "this library was built without -DPFT_TRACE (make libpft_b200_trace.so)");
#endif
}

extern "C" int pft_debug_trace_cta(pft_solver *s, unsigned long long *out, int launches)
{
	if (!s || !out) return fail(PFT_EINVAL, "null argument");
#ifdef PFT_TRACE
	CU(cudaSetDevice(s->device));
	CU(cudaStreamSynchronize(s->stream));
	const int n = std::min(launches, TRACE_CTA_LAUNCHES);
	if (!s->trace_cta) return 0;
	CU(cudaMemcpy(out, s->trace_cta, sizeof(unsigned long long) * 4096 * n, cudaMemcpyDeviceToHost));
	CU(cudaMemset(s->trace_cta, 0, sizeof(unsigned long long) * TRACE_CTA_LAUNCHES * 4096));
	return n;
#else
	(void)launches;
	return fail(PFT_EINVAL, "this library was built without -DPFT_TRACE (make libpft_b200_trace.so)");
#endif
}

extern "C" int pft_rasterise(pft_solver *s, const double *xyz_raw, int count)
{
	if (!s || (count > 0 && !xyz_raw) || count < 0) return fail(PFT_EINVAL, "bad sphere list");
	if (count == 0) return 0;
	CU(cudaSetDevice(s->device));
	const pft_case &cs = s->cs;
	std::vector<double> b(3 * (size_t)count);
	for (int q = 0; q < count; q++) {   // equation.c:480-482
		b[q] = xyz_raw[3 * q] * cs.param[P_beads_scaling] + cs.param[P_beads_offset_x];
		b[count + q] = xyz_raw[3 * q + 1] * cs.param[P_beads_scaling] + cs.param[P_beads_offset_y];
		b[2 * (size_t)count + q] = xyz_raw[3 * q + 2] * cs.param[P_beads_scaling] + cs.param[P_beads_offset_z];
	}
	double *d = nullptr;
	CU(cudaMalloc(&d, b.size() * sizeof(double)));
	CU(cudaMemcpyAsync(d, b.data(), b.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	const Geom &g = s->g;
	long long total = (long long)g.n1 * g.n2 * g.n3;
	const int threads = 256;
	long long blocks = (total + threads - 1) / threads;
	PFT_LAUNCH(rasterise_kernel, dim3((unsigned)blocks), dim3(threads), 3 * threads * sizeof(double), s->stream,
		s->gl, d, d + count, d + 2 * (size_t)count, count, g.n1, g.n2, g.n3, g.px, g.plane,
		cs.first_row, cs.total_n3, cs.L1, cs.L2, cs.L3, cs.param[P_xi_gl], cs.param[P_ball_radius]);
	s->st.kernel_launches++;
	CU(cudaGetLastError());
	int rc = (s->cs.nranks > 1) ? exchange_pair(s, s->gl, nullptr) : 0;
	CU(cudaStreamSynchronize(s->stream));
	cudaFree(d);
	return rc;
}

extern "C" int pft_ice_fraction(pft_solver *s, double *fraction)
{
	if (!s || !fraction) return fail(PFT_EINVAL, "null argument");
	CU(cudaSetDevice(s->device));
	unsigned long long *d = nullptr, h = 0;
	CU(cudaMalloc(&d, sizeof h));
	CU(cudaMemsetAsync(d, 0, sizeof h, s->stream));
	const Geom &g = s->g;
	PFT_LAUNCH(ice_count_kernel, dim3(148 * 4), dim3(256), 0, s->stream, s->arr[A_X][1], g.n1, g.n2, g.n3, g.px, g.plane, d);
	s->st.kernel_launches++;
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	cudaFree(d);
	*fraction = (double)h / ((double)g.n1 * g.n2 * g.n3);
	return 0;
}

// ------------------------------------------------------------------------------------------
// RHS with host buffers: the reference's RK_RightHandSide signature

extern "C" int pft_rhs_host(pft_solver *s, double t, const double *w, double *dw)
{
	if (!s || !w || !dw) return fail(PFT_EINVAL, "null argument");
	int rc = pft_upload(s, w);
	if (rc) return rc;
	pft_system dummy;
	memset(&dummy, 0, sizeof dummy);
	rc = launch_stage(s, 1, t, 0.0, 0.0, A_X, A_K1, nullptr, &dummy, 0.0, 1.0);
	if (rc) return rc;
	s->st.rhs_evals++;
	rc = ensure_stage(s);
	if (rc) return rc;
	const long long SIZE = s->padded / PFT_NVAR;
	CU(cudaMemcpyAsync(s->stage_buf, dw, (size_t)s->padded * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	if (s->heat) repack(s, s->stage_buf, s->arr[A_K1][0], 1);
	repack(s, s->stage_buf + SIZE, s->arr[A_K1][1], 1);
	CU(cudaMemcpyAsync(dw, s->stage_buf, (size_t)s->padded * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	// du/dt == 0 in calc_mode 10/11 and dgl/dt == 0 always (equation.c:708,731): interior rows only
	{
		const int N1 = s->N1, N2 = s->N2;
		const long long rowsize = (long long)N1 * N2;
		for (int q = 0; q < PFT_NVAR; q++) {
			if (q == 1 || (q == 0 && s->heat)) continue;
			for (int k = 0; k < s->g.n3; k++)
				for (int j = 0; j < s->g.n2; j++)
					memset(dw + q * SIZE + (k + PFT_BT) * rowsize + (long long)(j + PFT_BT) * N1 + PFT_BT, 0, sizeof(double) * (size_t)s->g.n1);
		}
	}
	s->k1_valid = false;
	return 0;
}

// ------------------------------------------------------------------------------------------
// the integrator

#define CMD_h_TOO_SMALL 1
#define CMD_NAN         2
#define CMD_UPDATE      4
#define CMD_FINISHED    8
#define CMD_NEXTFINISH  16
#define CMD_BREAK       32

static inline double host_ms()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int wait_slot(pft_solver *s, unsigned long long seq)
{
	volatile unsigned long long *p = &s->slot->seq;
	unsigned long spins = 0;
	while (*p != seq) {
		_mm_pause();
		if ((++spins & 0xfffff) == 0) {
			cudaError_t e = cudaStreamQuery(s->stream);
			if (e != cudaSuccess && e != cudaErrorNotReady)
				return fail(PFT_ECUDA, "device fault while waiting for stage 5: %s", cudaGetErrorString(e));
			if (e == cudaSuccess && *p != seq)
				return fail(PFT_ESTATE, "stream idle but stage 5 never published its result (seq %llu)", seq);
		}
	}
	return 0;
}

extern "C" int pft_solve(pft_solver *s, pft_system *sys, double final_time)
{
	if (!s || !sys) { fail(PFT_EINVAL, "null argument"); return PFT_SOLVE_BADSYSTEM; }
	if (sys->delta <= 0) { fail(PFT_EINVAL, "delta <= 0"); return PFT_SOLVE_BADSYSTEM; }        // hybrid2.c:286
	if (cudaSetDevice(s->device) != cudaSuccess) return fail(PFT_ECUDA, "cudaSetDevice failed");

	double t = sys->t, h = sys->h, new_h = 0.0;
	const double h_min = sys->h_min, delta = sys->delta;
	int command = 0;
	long attempts = 0;
	int xc = A_X, xn = A_XN;              // current / candidate state pair (swapped on accept)
	const bool multi = s->cs.nranks > 1;
	int rc = 0, ret = PFT_SOLVE_DONE;
	double t_host = 0.0;

	sys->last_nan = 0;                    // :316

	// :319-325
	if ((final_time > t && h < 0) || (final_time < t && h > 0)) h *= -1;
	if (h == 0 || fabs(final_time - t) <= fabs(h)) { h = final_time - t; command |= CMD_FINISHED; }

	cudaEventRecord(s->ev0, s->stream);
	t_host = host_ms();

	if (!(s->k1_valid && s->k1_t == t) || (s->flags & PFT_FLAG_RECOMPUTE_K1)) {
		if ((rc = launch_stage(s, 1, t, 0.0, 0.0, xc, A_K1, nullptr, sys, 0.0, 1.0))) goto device_error;   // :373
		sys->rhs_evals++; s->st.rhs_evals++;
	}
	s->k1_valid = false;

	for (;;) {
		const double h2 = h / 2.0, h3 = h / 3.0, h6 = h / 6.0, h8 = h / 8.0;    // :355
		const double eps_scale = fabs(h3);

		if ((rc = launch_stage(s, 2, t + h3, h3, h6, xc, A_A3, nullptr, sys, 0.0, 1.0))) goto device_error;   // :378-409
		if ((rc = launch_stage(s, 3, t + h3, 0.0, 0.0, xc, A_K3, nullptr, sys, 0.0, 1.0))) goto device_error; // :412
		if ((rc = launch_stage(s, 4, t + h2, h8, 0.0, xc, A_K4, nullptr, sys, 0.0, 1.0))) goto device_error;  // :417-432
		if ((rc = launch_stage(s, 5, t + h, h, h3, xc, xn, nullptr, sys, fabs(h), eps_scale))) goto device_error; // :437-524,657-668
		sys->rhs_evals += 4; s->st.rhs_evals += 4;
		const unsigned long long seq = s->seq;

		double eps; int nonfinite;
		if (!multi || s->p2p) {
			// (peer path: stage 5's last CTA has formed the global max + decision on the device, the role of
			// MPI_Allreduce(MAX), :572)
			if (s->speculate)          // K1 of the next step, gated on the device-side accept flag
				if ((rc = launch_stage(s, 1, t + h, 0.0, 0.0, xn, A_K1, &s->ctl->accept, sys, 0.0, 1.0))) goto device_error;
			const double tw = host_ms();
			s->st.host_launch_ms += tw - t_host;
			if ((rc = wait_slot(s, seq))) goto device_error;
			t_host = host_ms();
			s->st.host_wait_ms += t_host - tw;
			eps = s->slot->eps; nonfinite = s->slot->nonfinite;
		} else {
			// NCCL path (PFT_COMM=nccl, or no peer access): global max over the slabs through the host
			if ((rc = wait_slot(s, seq))) goto device_error;
			s->red_host[0] = s->slot->eps; s->red_host[1] = (double)s->slot->nonfinite;
			if (cudaMemcpyAsync(s->red, s->red_host, 2 * sizeof(double), cudaMemcpyHostToDevice, s->stream) != cudaSuccess) { rc = fail(PFT_ECUDA, "memcpy"); goto device_error; }
			{
				ncclResult_t r_ = g_nccl.AllReduce(s->red, s->red, 2, ncclDouble, ncclMax, s->comm, s->stream);
				if (r_ != ncclSuccess) { rc = fail(PFT_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r_)); goto device_error; }
			}
			if (cudaMemcpyAsync(s->red_host, s->red, 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) { rc = fail(PFT_ECUDA, "memcpy"); goto device_error; }
			if (cudaStreamSynchronize(s->stream) != cudaSuccess) { rc = fail(PFT_ECUDA, "sync after allreduce: %s", cudaGetErrorString(cudaGetLastError())); goto device_error; }
			eps = s->red_host[0]; nonfinite = s->red_host[1] != 0.0;
		}
		if (s->profile) drain_timers(s);

		sys->steps_total++; s->st.attempts++;                                    // :460
		attempts++;

		double max_eps = eps;
		if (sys->delta_local) max_eps *= eps_scale;                              // :578
		new_h = ((max_eps > 0.0) ? pow((delta / max_eps), 0.2) * 0.8 : 2.0) * h; // :580

		if (sys->handle_nan && nonfinite) {                                      // :493-503
			command |= CMD_NAN;
			if (h / (final_time - t) < 1e-11) command |= CMD_h_TOO_SMALL;
		}
		if (max_eps < delta || fabs(h) < h_min) {                                // :599-611
			command |= CMD_UPDATE;
			if (fabs(final_time - (t + h)) <= fabs(new_h)) command |= CMD_NEXTFINISH;
		}
		if (!multi || s->p2p) {
			const int host_accept = ((command & CMD_UPDATE) && !(command & CMD_NAN)) ? 1 : 0;
			if (host_accept != s->slot->accept) {
				rc = fail(PFT_ESTATE, "device/host acceptance mismatch (eps %.17g delta %.17g h %.17g)", eps, delta, h);
				goto device_error;
			}
		}

		if (command & CMD_NAN) {                                                 // :626-646
			sys->last_nan = 1;
			if (command & CMD_h_TOO_SMALL) { sys->t = t; s->k1_valid = true; s->k1_t = t; ret = PFT_SOLVE_NAN; goto done; }
			h /= 10;
			command = 0;
			if (s->flags & PFT_FLAG_RECOMPUTE_K1) {
				if ((rc = launch_stage(s, 1, t, 0.0, 0.0, xc, A_K1, nullptr, sys, 0.0, 1.0))) goto device_error;
				sys->rhs_evals++; s->st.rhs_evals++;
			}
		} else {
			if (command & CMD_UPDATE) {                                          // :651-734
				t += h;
				std::swap(xc, xn);
				sys->steps++; s->st.accepted++;
				if ((!multi || s->p2p) && s->speculate) { sys->rhs_evals++; s->st.rhs_evals++; }     // the gated stage 1 ran
				else {
					if ((rc = launch_stage(s, 1, t, 0.0, 0.0, xc, A_K1, nullptr, sys, 0.0, 1.0))) goto device_error;
					sys->rhs_evals++; s->st.rhs_evals++;
				}
				if (sys->service) {
					sys->t = t; sys->h = h;
					if (sys->service(final_time, t, h, sys->steps, sys->steps_total, sys->service_ctx)) command |= CMD_BREAK;
				}
				if (command & CMD_FINISHED) break;
				if (command & CMD_BREAK) { sys->t = t; sys->h = new_h; ret = PFT_SOLVE_INTERRUPTED; goto done_k1; }
			} else if (s->flags & PFT_FLAG_RECOMPUTE_K1) {
				if ((rc = launch_stage(s, 1, t, 0.0, 0.0, xc, A_K1, nullptr, sys, 0.0, 1.0))) goto device_error;
				sys->rhs_evals++; s->st.rhs_evals++;
			}
			if (command & CMD_NEXTFINISH) {                                      // :745-760
				sys->h = new_h;
				h = final_time - t;
				command = CMD_FINISHED;
			} else {
				command = 0;
				h = new_h;
			}
		}

		if (sys->max_attempts > 0 && attempts >= sys->max_attempts) {
			sys->t = t;
			if (!(command & CMD_FINISHED)) sys->h = h;
			ret = PFT_SOLVE_BUDGET;
			goto done_k1;
		}
	}
	sys->t = t;                                                                  // :768

done_k1:
	s->k1_valid = true; s->k1_t = t;      // K1 == f(t, x) holds at every exit that reaches here
done:
	cudaEventRecord(s->ev1, s->stream);
	if (cudaStreamSynchronize(s->stream) != cudaSuccess) { rc = fail(PFT_ECUDA, "stream sync at solve exit: %s", cudaGetErrorString(cudaGetLastError())); goto device_error; }
	if (s->profile) drain_timers(s);
	{
		float ms = 0.f;
		cudaEventElapsedTime(&ms, s->ev0, s->ev1);
		s->st.device_ms_total = ms;
	}
	// keep the accepted state in pair A_X so that upload/download/debug see a fixed place
	if (xc != A_X) {
		std::swap(s->arr[A_X][0], s->arr[A_XN][0]);
		std::swap(s->arr[A_X][1], s->arr[A_XN][1]);
		std::swap(s->tmap[A_X][0], s->tmap[A_XN][0]);
		std::swap(s->tmap[A_X][1], s->tmap[A_XN][1]);
		std::swap(s->off[A_X][0], s->off[A_XN][0]);
		std::swap(s->off[A_X][1], s->off[A_XN][1]);
	}
	return ret;

device_error:
	cudaStreamSynchronize(s->stream);
	drain_timers(s);
	s->k1_valid = false;
	return rc;
}

extern "C" int pft_get_stats(pft_solver *s, pft_stats *out)
{
	if (!s || !out) return fail(PFT_EINVAL, "null argument");
	*out = s->st;
	return 0;
}

extern "C" int pft_reset_stats(pft_solver *s)
{
	if (!s) return fail(PFT_EINVAL, "null solver");
	pft_stats keep = s->st;
	memset(&s->st, 0, sizeof s->st);
	s->st.tile_x = keep.tile_x; s->st.tile_y = keep.tile_y; s->st.z_chunk = keep.z_chunk;
	s->st.threads = keep.threads; s->st.grid_blocks = keep.grid_blocks; s->st.bytes_per_field = keep.bytes_per_field;
	return 0;
}
