// pft_b200.cu -- C ABI (include/pft_b200.h) over the sm_100a kernels of pft_kernels.cuh.
//
// Host side of the Runge-Kutta-Merson integrator
// (modules/RK_MPI_SAsolver_hybrid2/RK_MPI_SAsolver_hybrid2.c:215-770 of the reference) for a
// device-resident state:
//   * the five RHS evaluations and all vector combinations run as five fused stage kernels;
//   * the step-size controller runs on the DEVICE too (pft_ctl.h: one thread of stage 5's last CTA),
//     with a twin of the host C library's pow() that is checked against that library at start-up
//     (pft_pow.h), so the accepted-step sequence is the reference's bit for bit;
//   * every scalar a stage needs is read from the device-resident control block, so the launches of
//     an attempt do not depend on the previous outcome: the host enqueues attempts ahead of the
//     device and sleeps on a progress slot -- no host round trip and no host spin in the time loop;
//   * after a rejected attempt K1 is reused (x and t did not change), which the reference
//     recomputes needlessly (:373): stage 1 of the next attempt returns at once.
//
// This file is compiled with -fmad=false and the host compiler with -ffp-contract=off: every
// floating-point expression keeps the reference's association.
#include "../../include/pft_b200.h"
#include "pft_kernels.cuh"
#include "pft_kernels_tma.cuh"
#include "pft_kernels_v2.cuh"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <thread>
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
extern "C" const char *pft_version(void) { return "pft-b200 0.2 (sm_100a, float64, FAST variant: FMA contraction allowed, not bit-exact)"; }
#else
extern "C" const char *pft_version(void) { return "pft-b200 0.2 (sm_100a, float64, strict IEEE association)"; }
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

enum { A_X = 0, A_K1, A_A3, A_K3, A_K4, A_XN, A_K1B, A_COUNT };   // array pairs (u,p); order = pft_debug_fetch `which`
// (A_K1B: speculative stage 1 keeps one K1 buffer per state pair -- A_K1 goes with A_X, A_K1B with A_XN)

struct StageTimer { cudaEvent_t a, b; int stage; };

// how one stage is launched.  ver 0: merson_stage (per-thread loads), 1: merson_stage_tma, 2: merson_stage_v2
struct StageCfg {
	int ver, R;           // kernel generation; cells per thread (rows)
	bool nobar;           // v2, stages 1 and 3: no CTA barrier in the plane loop (the last warp to leave a plane refills its buffer)
	bool fixed;           // v2: an instantiation with this tile shape as compile-time constants exists
	int TX, TY;           // threads; the tile is TX x (TY*R) cells
	int bx, by, as;       // TMA box (TX+4) x (TY*R+2) and the doubles per array buffer
	int nzc, zc;          // z-chunks and planes per chunk
	dim3 grid, block;
	size_t smem;
};

struct pft_solver {
	pft_case cs;
	unsigned flags;
	int device;
	Consts c;
	Geom g;
	int model;            // 0,1,2
	bool heat;
	StageCfg cfg[5];      // launch geometry and kernel variant of each of the five stages
	long long field;      // doubles per field array
	double *arr[A_COUNT][2];   // [pair][u|p]
	double *gl;
	double *stage_buf;    // padded-layout staging (3*SIZE doubles)
	bool stage_shell[PFT_NVAR];  // the staging image of that variable holds the ghost shell of the caller's host array
	long long padded;     // 3*SIZE
	int N1, N2, N3;
	// device-resident step control (pft_ctl.h)
	StepCtl *ctl;         // device
	StepCtl *ctl_host;    // pinned: staging of the block the host writes at solve entry
	HostSlot *slot;       // pinned + mapped: progress slot written by the controller thread
	HostSlot *slot_dev;   // device alias
	PowTables *tab_dev;   // device copy of the host library's pow tables (nullptr: host pow)
	bool dev_pow;
	int cur;              // which state pair holds x at rest (mirror of StepCtl.cur)
	unsigned long long aseq;   // attempts started on this solver so far (mirror of StepCtl.aseq at rest)
	int depth;            // attempts the host keeps enqueued ahead of the device
	cudaStream_t stream;
	cudaEvent_t ev0, ev1;
	bool k1_valid;
	double k1_t;
	// speculative stage 1 (pft_ctl.h): attempts are launched as [S2 S3 S4 S5 S1'], two control blocks alternate
	bool spec_ok;         // the build, the kernel variant and the flags allow it
	bool spec;            // ... and so does the transport (set by build_params)
	bool spec_earlyout;   // S1' CTAs that start after a rejecting decision return at once (PFT_SPEC_EARLYOUT=0: A/B)
	bool k1_uncounted;    // K1 of the state at rest came from an S1' that no counter has seen yet (rhs_evals)
	unsigned long long hbase;   // halo-flag sequence base at rest (mirror of StepCtl.hbase)
	ncclComm_t comm;
	bool have_comm;
	bool profile;
	bool pdl;             // stage kernels use programmatic dependent launch (off while profiling stages)
	std::vector<StageTimer> timers;
	pft_stats st;
	// one device slab: [SyncBlock][15 field arrays (13 without speculative stage 1)]; peers map it (CUDA IPC between processes, peer access inside one)
	char *slab;
	size_t slab_bytes;
	size_t off[A_COUNT][2], off_gl;            // byte offsets of the arrays inside the slab (same on every rank)
	// peer path (multi-GPU without NCCL in the time loop)
	bool p2p;
	bool ipc;                                  // peer_base[] came from cudaIpcOpenMemHandle
	char *peer_base[PFT_MAX_RANKS];            // every rank's slab in this process' address space
	int peer_n3[PFT_MAX_RANKS];
	unsigned int *halo_tickets;
	unsigned long long *halo_published;
	pft_group *group;                          // member of an in-process group (pft_group_*)
	// TMA path (heat modes): one tensor map per array, box = tile + halo ring
	bool use_tma;
	// launch arguments of the five stages: independent of the step length and of earlier outcomes (pft_ctl.h)
	bool params_ok;
	KParams kp[5];
	TmaParams tp[5];
	// stage-to-stage data flow (pft_kernels.cuh, dep_wait): completion words of the CTAs of stages 1-4
	bool dataflow;
	unsigned long long *done[5];
	unsigned long long lseq;                   // attempts enqueued so far (host count; the flags carry it)
};

// one host process, several GPUs: slab r of the block lives on devices[r]
struct pft_group {
	int n;
	pft_case whole;
	unsigned flags;
	std::vector<int> devices, planes;
	bool balance_tried;       // pft_group_solve rebalances once by itself
	std::vector<pft_solver *> member;
	// service callback agreement: slab 0's worker calls the user's callback, the others take its verdict
	pft_service_fn user_service;
	void *user_ctx;
	std::atomic<long> cb_seq;
	std::atomic<int> cb_val;
};

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
	// small grids (50x50x100: 10 tiles): with fewer than two CTAs per SM the launch is latency-bound, and
	// thinner chunks pay -- 4 planes per chunk measured +34% there, while 100x100x200 (1000 CTAs at 8
	// planes) loses 4-20% when cut thinner (profiles/r02_tune2_small.txt)
	if (tiles_xy * nzc < 2L * sms) {
		const long want = (2L * sms + tiles_xy - 1) / tiles_xy;
		nzc = std::max(nzc, std::min(want, (long)std::max(1, n3 / 4)));
	}
	return (int)nzc;
}

// ------------------------------------------------------------------------------------------

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D tensor map over one field array: (px, n2, n3+2) doubles, box (bx, by, 1)
static int make_tensor_map(pft_solver *s, CUtensorMap *m, double *base, int bx, int by)
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
	cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1u};
	cuuint32_t estr[3] = {1u, 1u, 1u};
	CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dim, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) return fail(PFT_ECUDA, "cuTensorMapEncodeTiled failed (%d) for box %dx%d", (int)r, bx, by);
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

// ------------------------------------------------------------------------------------------
// the host library's pow tables (pft_pow.h): loaded and checked once per process

struct PowGate {
	PowTables tab;
	bool usable = false;
	char why[256] = "";
};

static const PowGate &pow_gate()
{
	static PowGate g;
	static std::once_flag once;
	std::call_once(once, []() {
		if (env_int("PFT_HOST_POW", 0)) { snprintf(g.why, sizeof g.why, "disabled by PFT_HOST_POW"); return; }
		if (pft_pow_load_tables(&g.tab, g.why, sizeof g.why)) return;
		const long bad = pft_pow_selftest(g.tab, 0.2, 400000, 0x5eedull);
		if (bad) { snprintf(g.why, sizeof g.why, "the twin differs from the C library's pow() on %ld of 400000 operands", bad); return; }
		g.usable = true;
	});
	return g;
}

// ------------------------------------------------------------------------------------------
// stage launch

template <class K, class PARAMS>
static cudaError_t launch_kernel(pft_solver *s, K fn, bool *attr_set_dev, const StageCfg &cf, const PARAMS &kp)
{
	bool &attr_set = attr_set_dev[s->device & 63];        // the attribute is per device (and per template instance)
	if (!attr_set) {
		cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
		if (e != cudaSuccess) return e;
		attr_set = true;
	}
	PFT_LAUNCH_PDL(s->pdl, fn, cf.grid, cf.block, cf.smem, s->stream, kp);
	return cudaGetLastError();
}

template <int STAGE, int MODEL, bool HEAT>
static cudaError_t launch_t(pft_solver *s, const KParams &kp)
{
	static bool attr[64] = {};
	return launch_kernel(s, merson_stage<STAGE, MODEL, HEAT>, attr, s->cfg[STAGE - 1], kp);
}

template <int STAGE>
static cudaError_t launch_stage_t(pft_solver *s, const KParams &kp)
{
	if (s->model == 2) return launch_t<STAGE, 2, true>(s, kp);
	if (s->model == 0) return s->heat ? launch_t<STAGE, 0, true>(s, kp) : launch_t<STAGE, 0, false>(s, kp);
	return s->heat ? launch_t<STAGE, 1, true>(s, kp) : launch_t<STAGE, 1, false>(s, kp);
}

template <int STAGE, int MODEL>
static cudaError_t launch_tma_t(pft_solver *s, const TmaParams &kp)
{
	static bool attr[64] = {};
	return launch_kernel(s, merson_stage_tma<STAGE, MODEL>, attr, s->cfg[STAGE - 1], kp);
}

// second-generation kernel: R cells per thread, optional compile-time tile shape, optional barrier-free plane loop
template <int STAGE, int MODEL, int R, int CTX, int CTY, bool NOBAR>
static cudaError_t launch_v2_t(pft_solver *s, const TmaParams &kp)
{
	static bool attr[64] = {};
	return launch_kernel(s, merson_stage_v2<STAGE, MODEL, R, CTX, CTY, NOBAR>, attr, s->cfg[STAGE - 1], kp);
}

// the tile shapes that exist as compile-time instantiations (threads TX x TY; cells TX x TY*R)
#define PFT_FIXED_TX 50
#define PFT_FIXED_TY 5
static bool have_fixed_shape(int R, int TX, int TY) { return (R == 1 || R == 2) && TX == PFT_FIXED_TX && TY == PFT_FIXED_TY; }

template <int STAGE, int MODEL, int R>
static cudaError_t launch_v2_shape(pft_solver *s, const TmaParams &kp)
{
	const StageCfg &cf = s->cfg[STAGE - 1];
	constexpr bool CAN_NOBAR = (STAGE == 1 || STAGE == 3);
	// forming stages with the split (arrive / wait) barrier exist for the default geometry only: fixed shape, one row per thread
	constexpr bool CAN_SPLIT = !CAN_NOBAR && R == 1;
	if (cf.fixed) {
		if (CAN_NOBAR && cf.nobar) return launch_v2_t<STAGE, MODEL, R, PFT_FIXED_TX, PFT_FIXED_TY, CAN_NOBAR>(s, kp);
		if (CAN_SPLIT && cf.nobar) return launch_v2_t<STAGE, MODEL, R, PFT_FIXED_TX, PFT_FIXED_TY, CAN_SPLIT>(s, kp);
		return launch_v2_t<STAGE, MODEL, R, PFT_FIXED_TX, PFT_FIXED_TY, false>(s, kp);
	}
	if (CAN_NOBAR && cf.nobar) return launch_v2_t<STAGE, MODEL, R, 0, 0, CAN_NOBAR>(s, kp);
	return launch_v2_t<STAGE, MODEL, R, 0, 0, false>(s, kp);
}

template <int STAGE, int MODEL>
static cudaError_t launch_v2_rows(pft_solver *s, const TmaParams &kp)
{
	return s->cfg[STAGE - 1].R == 2 ? launch_v2_shape<STAGE, MODEL, 2>(s, kp) : launch_v2_shape<STAGE, MODEL, 1>(s, kp);
}

template <int STAGE>
static cudaError_t launch_stage_tma_t(pft_solver *s, const TmaParams &kp)
{
	if (s->cfg[STAGE - 1].ver == 2) {
		if (s->model == 2) return launch_v2_rows<STAGE, 2>(s, kp);
		if (s->model == 0) return launch_v2_rows<STAGE, 0>(s, kp);
		return launch_v2_rows<STAGE, 1>(s, kp);
	}
	if (s->model == 2) return launch_tma_t<STAGE, 2>(s, kp);
	if (s->model == 0) return launch_tma_t<STAGE, 0>(s, kp);
	return launch_tma_t<STAGE, 1>(s, kp);
}

// peer path: where a stage pushes its boundary planes and whom it signals
static void fill_peer_args(pft_solver *s, PeerArgs *pr, int stage)
{
	const int r = s->cs.rank, R = s->cs.nranks;
	const size_t plane_b = (size_t)s->g.plane * sizeof(double);
	const int outs[5] = {A_K1, A_A3, A_K3, A_K4, -1};
	pr->rank = r; pr->nranks = R;
	pr->nzc = s->cfg[stage - 1].nzc;
	pr->my_flags = ((SyncBlock *)s->slab)->halo_flag;
	pr->tickets = s->halo_tickets;
	pr->published = s->halo_published;
	pr->wait_lo = (r > 0); pr->wait_hi = (r < R - 1);
	for (int q = 0; q < R; q++) pr->ranks[q] = (SyncBlock *)s->peer_base[q];
	for (int oi = 0; oi < 2; oi++) {
		// stage 5 writes the candidate state pair: index [cur ^ 1], i.e. entry 0 = pair A_X, entry 1 = pair A_XN
		// (speculative stage 1: K1 likewise has a buffer per pair, see out_ptrs)
		const int out = (stage == 5) ? (oi == 0 ? A_X : A_XN) : (stage == 1 && s->spec && oi == 1) ? A_K1B : outs[stage - 1];
		if (r > 0) {        // the lower neighbour's UPPER ghost plane (index n3'+1) and its halo_flag[1]
			char *pb = s->peer_base[r - 1];
			const size_t gp = (size_t)(s->peer_n3[r - 1] + 1) * plane_b;
			pr->lo_p[oi] = (double *)(pb + s->off[out][1] + gp);
			pr->lo_u[oi] = s->heat ? (double *)(pb + s->off[out][0] + gp) : nullptr;
			pr->lo_flag = &((SyncBlock *)pb)->halo_flag[1];
		}
		if (r < R - 1) {    // the upper neighbour's LOWER ghost plane (index 0) and its halo_flag[0]
			char *pb = s->peer_base[r + 1];
			pr->hi_p[oi] = (double *)(pb + s->off[out][1]);
			pr->hi_u[oi] = s->heat ? (double *)(pb + s->off[out][0]) : nullptr;
			pr->hi_flag = &((SyncBlock *)pb)->halo_flag[0];
		}
	}
}

// The argument blocks of the five stages.  Built once (and again when the transport changes):
// nothing in them depends on the step length, the time level or the outcome of an attempt.
static int build_params(pft_solver *s)
{
	const bool separate_nccl = s->cs.nranks > 1 && !s->p2p;
	s->spec = s->spec_ok && !separate_nccl;
	for (int stage = 1; stage <= 5; stage++) {
		const StageCfg &cf = s->cfg[stage - 1];
		KParams &kp = s->kp[stage - 1];
		kp.c = s->c;
		kp.g = s->g;
		kp.g.zc = cf.zc;
		kp.smem_bytes = (unsigned)cf.smem;
		StageArgs &a = kp.a;
		memset(&a, 0, sizeof a);
		a.ctl = s->ctl;
		a.xu[0] = s->arr[A_X][0]; a.xp[0] = s->arr[A_X][1];
		a.xu[1] = s->arr[A_XN][0]; a.xp[1] = s->arr[A_XN][1];
		a.gl = s->gl;
		a.k1u = s->arr[A_K1][0]; a.k1p = s->arr[A_K1][1];
		a.k3u = s->arr[A_K3][0]; a.k3p = s->arr[A_K3][1];
		a.k4u = s->arr[A_K4][0]; a.k4p = s->arr[A_K4][1];
		a.au = s->arr[A_A3][0]; a.ap = s->arr[A_A3][1];
		const int outs[5] = {A_K1, A_A3, A_K3, A_K4, A_XN};
		a.ou = s->arr[outs[stage - 1]][0]; a.op = s->arr[outs[stage - 1]][1];
		if (s->p2p) fill_peer_args(s, &a.pr, stage);
		else a.pr.nzc = cf.nzc;
		a.pr.rank = s->cs.rank;
		a.pr.nranks = s->p2p ? s->cs.nranks : 1;       // the in-kernel error-max exchange exists on the peer path only
		a.dataflow = s->dataflow ? 1 : 0;
		a.done_mine = (s->dataflow && (stage <= 4 || s->spec)) ? s->done[stage - 1] : nullptr;
		a.done_prev = (s->dataflow && stage >= 2) ? s->done[stage - 2] : (s->spec && stage == 1) ? s->done[4] : nullptr;
		a.spec = s->spec ? 1 : 0;
		a.k1u2[0] = s->arr[A_K1][0]; a.k1p2[0] = s->arr[A_K1][1];
		a.k1u2[1] = s->arr[A_K1B][0]; a.k1p2[1] = s->arr[A_K1B][1];
		if (stage == 5) {
			a.slot = s->slot_dev;
			a.tab = s->tab_dev;
			a.nblocks = cf.grid.x * cf.grid.y * cf.grid.z;
			a.fin_mode = separate_nccl ? FIN_STORE_LOCAL : (s->dev_pow ? FIN_DECIDE : FIN_PUBLISH_EPS);
		}
		if (s->use_tma) {
			TmaParams &tp = s->tp[stage - 1];
			tp.c = kp.c; tp.g = kp.g; tp.a = kp.a;
			tp.bx = cf.bx; tp.by = cf.by; tp.as = cf.as; tp.tx = cf.TX; tp.ty = cf.TY; tp.smem_bytes = (unsigned)cf.smem;
			// array order of a stage: p-type arrays, u-type arrays, gl (pft_kernels_tma.cuh, stage_np); one tensor
			// map per array with this stage's box
			const int src[5][4] = {{A_X, -1, -1, -1}, {A_X, A_K1, -1, -1}, {A_A3, -1, -1, -1}, {A_X, A_K1, A_K3, -1}, {A_X, A_K1, A_K3, A_K4}};
			const int np = stage_np(stage);
			int rc = 0;
			for (int m = 0; m < np && !rc; m++) {
				rc = make_tensor_map(s, &tp.maps[m], s->arr[src[stage - 1][m]][1], cf.bx, cf.by);
				if (!rc) rc = make_tensor_map(s, &tp.maps[np + m], s->arr[src[stage - 1][m]][0], cf.bx, cf.by);
			}
			if (!rc) rc = make_tensor_map(s, &tp.maps[2 * np], s->gl, cf.bx, cf.by);
			if (!rc) rc = make_tensor_map(s, &tp.xalt[0], s->arr[A_XN][1], cf.bx, cf.by);
			if (!rc) rc = make_tensor_map(s, &tp.xalt[1], s->arr[A_XN][0], cf.bx, cf.by);
			if (!rc && s->spec && (stage == 2 || stage >= 4)) {
				rc = make_tensor_map(s, &tp.k1alt[0], s->arr[A_K1B][1], cf.bx, cf.by);
				if (!rc) rc = make_tensor_map(s, &tp.k1alt[1], s->arr[A_K1B][0], cf.bx, cf.by);
			}
			if (rc) return rc;
		}
	}
	s->params_ok = true;
	return 0;
}

static int exchange_pair(pft_solver *s, double *u, double *p);

// one stage launch.  standalone != nullptr: a single right-hand-side evaluation outside the time loop
static int launch_stage(pft_solver *s, int stage, const double *standalone_dir = nullptr, bool s1spec = false)
{
	if (!s->params_ok) { int rc_ = build_params(s); if (rc_) return rc_; }
	StageTimer tm{};
	if (s->profile) {
		cudaEventCreate(&tm.a); cudaEventCreate(&tm.b); tm.stage = stage;
		cudaEventRecord(tm.a, s->stream);
	}
	cudaError_t e;
	s->tp[stage - 1].a.lseq = s->lseq;
	s->kp[stage - 1].a.lseq = s->lseq;
	s->tp[stage - 1].a.s1spec = s1spec ? (s->spec_earlyout ? 1 : 2) : 0;
	if (s->use_tma) {
		const TmaParams *tp = &s->tp[stage - 1];
		TmaParams alone;
		if (standalone_dir) {
			alone = *tp;
			memset(&alone.a.pr, 0, sizeof alone.a.pr);
			alone.a.pr.nzc = s->cfg[stage - 1].nzc; alone.a.pr.nranks = 1;
			alone.a.standalone = 1; alone.a.standalone_cur = s->cur; alone.a.dir_standalone = *standalone_dir;
			alone.a.dataflow = 0; alone.a.spec = 0; alone.a.s1spec = 0;
			tp = &alone;
		}
		switch (stage) {
		case 1: e = launch_stage_tma_t<1>(s, *tp); break;
		case 2: e = launch_stage_tma_t<2>(s, *tp); break;
		case 3: e = launch_stage_tma_t<3>(s, *tp); break;
		case 4: e = launch_stage_tma_t<4>(s, *tp); break;
		default: e = launch_stage_tma_t<5>(s, *tp); break;
		}
	} else {
		const KParams *kp = &s->kp[stage - 1];
		KParams alone;
		if (standalone_dir) {
			alone = *kp;
			memset(&alone.a.pr, 0, sizeof alone.a.pr);
			alone.a.pr.nzc = s->cfg[stage - 1].nzc; alone.a.pr.nranks = 1;
			alone.a.standalone = 1; alone.a.standalone_cur = s->cur; alone.a.dir_standalone = *standalone_dir;
			kp = &alone;
		}
		switch (stage) {
		case 1: e = launch_stage_t<1>(s, *kp); break;
		case 2: e = launch_stage_t<2>(s, *kp); break;
		case 3: e = launch_stage_t<3>(s, *kp); break;
		case 4: e = launch_stage_t<4>(s, *kp); break;
		default: e = launch_stage_t<5>(s, *kp); break;
		}
	}
	if (e != cudaSuccess) return fail(PFT_ECUDA, "stage %d launch: %s", stage, cudaGetErrorString(e));
	if (s->profile) { cudaEventRecord(tm.b, s->stream); s->timers.push_back(tm); }
	s->st.kernel_launches++;
	s->st.stage_launches[stage - 1]++;
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
	if (s->group) return fail(PFT_ESTATE, "members of an in-process group exchange their boundary planes through pft_group_upload");
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

// Peer path between processes: every rank exports its slab with CUDA IPC, the handles travel once
// through the NCCL communicator, and from then on the time loop contains no collective call at all:
// boundary planes are stored straight into the neighbours' ghost planes by the stage kernels, flags
// and the per-slab error maxima travel as peer stores (pft_kernels.cuh, "peer-memory halo exchange").
static int setup_p2p(pft_solver *s)
{
	const int R = s->cs.nranks, r = s->cs.rank;
	if (R > PFT_MAX_RANKS) return fail(PFT_EINVAL, "peer path supports at most %d ranks", PFT_MAX_RANKS);
	struct Export { cudaIpcMemHandle_t h; int n3, first_row; int ok; unsigned long long bytes; };
	Export mine;
	memset(&mine, 0, sizeof mine);
	mine.ok = (cudaIpcGetMemHandle(&mine.h, s->slab) == cudaSuccess);
	if (!mine.ok) cudaGetLastError();
	mine.n3 = s->cs.n3;
	mine.first_row = s->cs.first_row;
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
	int row = all[0].first_row;
	for (int q = 0; q < R; q++) {
		ok = ok && all[q].ok;
		if (all[q].bytes != s->slab_bytes) return fail(PFT_EINVAL, "rank %d sized its slab differently (%llu vs %zu bytes): every rank must pass the same max_n3", q, all[q].bytes, s->slab_bytes);
		if (all[q].first_row != row || all[q].n3 < 1) return fail(PFT_EINVAL, "the slabs do not tile the grid: rank %d starts at plane %d, expected %d", q, all[q].first_row, row);
		row += all[q].n3;
		s->peer_n3[q] = all[q].n3;
	}
	if (row - all[0].first_row != s->cs.total_n3) return fail(PFT_EINVAL, "the slabs hold %d planes, the grid has %d", row - all[0].first_row, s->cs.total_n3);
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
	s->ipc = s->p2p;
	if (!s->p2p)
		for (int q = 0; q < R; q++) {
			if (s->peer_base[q] && s->peer_base[q] != s->slab) cudaIpcCloseMemHandle(s->peer_base[q]);
			s->peer_base[q] = nullptr;
		}
	return 0;
}

extern "C" int pft_comm_init(pft_solver *s, const void *id128)
{
	if (!s || !id128) return fail(PFT_EINVAL, "null argument");
	if (s->cs.nranks <= 1) return 0;
	if (s->group) return fail(PFT_ESTATE, "this solver belongs to an in-process group");
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
	s->params_ok = false;
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
	if (cs->max_n3 != 0 && cs->max_n3 < cs->n3) return fail(PFT_EINVAL, "max_n3 %d is smaller than this slab (%d planes)", cs->max_n3, cs->n3);
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
	s->profile = env_int("PFT_PROFILE_STAGES", 0) != 0;
	s->pdl = !s->profile && env_int("PFT_NO_PDL", 0) == 0;
	s->depth = (flags & PFT_FLAG_NO_RUNAHEAD) ? 1 : std::max(1, std::min(64, env_int("PFT_RUNAHEAD", 6)));
	make_consts(s);

	Geom &g = s->g;
	g.n1 = cs->n1; g.n2 = cs->n2; g.n3 = cs->n3;
	g.px = (cs->n1 + 3) & ~3;
	g.plane = (long long)g.px * cs->n2;
	g.front = (cs->rank == 0);
	g.rear = (cs->rank == cs->nranks - 1);
	// every rank sizes its arrays for the deepest slab so that array offsets inside the slab are
	// the same on all ranks (peers address each other's arrays by offset)
	s->field = g.plane * ((cs->max_n3 > 0 ? cs->max_n3 : (cs->total_n3 + cs->nranks - 1) / cs->nranks) + 2);
	if (s->field < g.plane * (cs->n3 + 2)) { delete s; return fail(PFT_EINVAL, "slab of %d planes exceeds the run's deepest slab; pass max_n3", cs->n3); }
	s->N1 = cs->n1 + 2 * PFT_BT; s->N2 = cs->n2 + 2 * PFT_BT; s->N3 = cs->n3 + 2 * PFT_BT;
	s->padded = (long long)PFT_NVAR * s->N1 * s->N2 * s->N3;

	s->use_tma = s->heat && !(flags & PFT_FLAG_LEGACY_LOADS) && !env_int("PFT_LEGACY", 0);
	{
		// per-stage kernel variant: PFT_KERNEL = 1 | 2 (generation), PFT_ROWS = cells per thread, one digit per stage
		// ("2" or "21211"), PFT_NOBAR = 0 | 1 (stages 1 and 3 without CTA barrier), PFT_FIXED_SHAPE = 0 | 1
		const int kver = s->use_tma ? std::max(1, std::min(2, env_int("PFT_KERNEL", 2))) : 0;
		const char *rows = getenv("PFT_ROWS");
		const int nrows = (rows && *rows) ? (int)strlen(rows) : 0;
		const bool want_nobar = env_int("PFT_NOBAR", 1) != 0, want_fixed = env_int("PFT_FIXED_SHAPE", 1) != 0;
		for (int st = 0; st < 5; st++) {
			StageCfg &cf = s->cfg[st];
			cf.ver = kver;
			cf.R = 1;
			if (kver == 2 && nrows > 0) cf.R = std::max(1, std::min(2, (int)(rows[std::min(st, nrows - 1)] - '0')));
			// stages 1, 3: barrier-free plane loop (default); stages 2, 4, 5: split barrier (PFT_SPLITBAR=1, fixed shape and R == 1 only)
			// (PFT_SPLITBAR = 1: all three; 2 | 4 | 5: that stage only)
			const int sb = env_int("PFT_SPLITBAR", 0);
			cf.nobar = kver == 2 && ((st == 0 || st == 2) ? want_nobar : (sb == 1 || sb == st + 1));
			int max_threads = env_int("PFT_THREADS", 256);     // two or three CTAs per SM overlap each other's barrier waits
			if (max_threads < 32) max_threads = 32;
			if (max_threads > 512) max_threads = 512;
			if (s->use_tma && max_threads > 256) max_threads = 256;      // __launch_bounds__ of the TMA kernels
			choose_tile(cs->n1, (cs->n2 + cf.R - 1) / cf.R, max_threads, &cf.TX, &cf.TY);
			if (env_int("PFT_TX", 0) > 0) cf.TX = env_int("PFT_TX", 0);
			if (env_int("PFT_TY", 0) > 0) cf.TY = env_int("PFT_TY", 0);
			if (cf.TX & 1) cf.TX++;                  // tile origins must be even (TMA inner-offset alignment)
			if (cf.R > 1 && cf.TX * cf.TY < 2 * cf.TX + 2 * cf.TY * cf.R) cf.R = 1;    // a forced tile too small for the taller ring
			if (cf.TX * cf.TY > (s->use_tma ? 256 : 512) || cf.TX * cf.TY < 2 * cf.TX + 2 * cf.TY * cf.R) {
				int tx_ = cf.TX, ty_ = cf.TY;
				delete s;
				return fail(PFT_EINVAL, "tile %dx%d unusable (needs TX*TY <= 512 and enough threads for the halo ring)", tx_, ty_);
			}
			cf.fixed = kver == 2 && want_fixed && have_fixed_shape(cf.R, cf.TX, cf.TY);
			const int rows_t = cf.TY * cf.R;
			const int tiles_xy = ((cs->n1 + cf.TX - 1) / cf.TX) * ((cs->n2 + rows_t - 1) / rows_t);
			cf.nzc = choose_chunks(tiles_xy, cs->n3, prop.multiProcessorCount);
			if (env_int("PFT_ZCHUNKS", 0) > 0) cf.nzc = std::min(cs->n3, env_int("PFT_ZCHUNKS", 0));
			cf.zc = (cs->n3 + cf.nzc - 1) / cf.nzc;
			cf.nzc = (cs->n3 + cf.zc - 1) / cf.zc;
			cf.grid = dim3((cs->n1 + cf.TX - 1) / cf.TX, (cs->n2 + rows_t - 1) / rows_t, cf.nzc);
			if (kver == 2) cf.block = dim3(cf.TX * cf.TY, 1, 1);
			else cf.block = dim3(cf.TX, cf.TY, 1);
			if (!s->use_tma) {
				int NF = s->heat ? 3 : 1;
				size_t ts = (size_t)(cf.TX + 2) * (cf.TY + 2);
				cf.smem = (2 * NF * ts) * sizeof(double);
				cf.bx = cf.by = cf.as = 0;
			} else {
				cf.bx = cf.TX + 4;                   // box starts at x = i0-2 (16-byte aligned inner offset), TX is even
				cf.by = rows_t + 2;
				cf.as = (cf.bx * cf.by + 15) & ~15;  // each array buffer starts 128-byte aligned
				const int na = 2 * stage_np(st + 1) + 1;
				const bool form = !(st == 0 || st == 2);
				const int ntile = kver == 2 ? 6 : 4;  // state tiles: double buffered (v1), one per raw buffer (v2)
				cf.smem = ((size_t)TMA_STAGES * na * cf.as + (form ? ntile * (size_t)cf.as : 0)) * sizeof(double);
				if (cf.smem > 225 * 1024) { size_t b_ = cf.smem; int tx_ = cf.TX, ty_ = cf.TY; delete s; return fail(PFT_EINVAL, "tile %dx%d needs %zu bytes of shared memory", tx_, ty_, b_); }
			}
		}
		g.zc = s->cfg[0].zc;
	}

	CUS(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
	CUS(cudaEventCreate(&s->ev0)); CUS(cudaEventCreate(&s->ev1));
	const size_t fb = ((size_t)s->field * sizeof(double) + 255) & ~(size_t)255;
	// speculative stage 1 needs the second-generation kernels with stage-to-stage data flow and the device controller
	// (checked again below, once those are known to be on); PFT_SPEC=0 keeps the classic launch order
	const bool want_spec = s->cfg[0].ver == 2 && s->pdl && env_int("PFT_DATAFLOW", 1) != 0 && env_int("PFT_SPEC", 1) != 0 &&
	                       pow_gate().usable && !(flags & (PFT_FLAG_HOST_POW | PFT_FLAG_RECOMPUTE_K1));
	{
		// layout: identical on every rank of a run with equal slab depth; offsets are exchanged anyway
		size_t off = (sizeof(SyncBlock) + 255) & ~(size_t)255;
		for (int a = 0; a < A_COUNT; a++)
			for (int q = 0; q < 2; q++) {
				// calc_mode 10/11: du/dt == 0 (equation.c:708), so u is static: no K/aux arrays for it,
				// and the candidate state shares the one temperature array
				if (!s->heat && q == 0 && a != A_X) { s->off[a][q] = 0; continue; }
				if (a == A_K1B && !want_spec) { s->off[a][q] = 0; continue; }
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
	CUS(cudaMalloc(&s->halo_tickets, 16 * sizeof(unsigned int)));
	CUS(cudaMemsetAsync(s->halo_tickets, 0, 16 * sizeof(unsigned int), s->stream));
	CUS(cudaMalloc(&s->halo_published, 2 * sizeof(unsigned long long)));
	CUS(cudaMemsetAsync(s->halo_published, 0, 2 * sizeof(unsigned long long), s->stream));
	{
		// data flow between stages needs one geometry for all five stages (second-generation kernels, same R) and
		// programmatic dependent launch (without it the kernels serialise anyway)
		bool same = s->cfg[0].ver == 2 && s->pdl && env_int("PFT_DATAFLOW", 1) != 0;
		for (int st = 1; st < 5; st++)
			same = same && s->cfg[st].ver == 2 && s->cfg[st].grid.x == s->cfg[0].grid.x && s->cfg[st].grid.y == s->cfg[0].grid.y &&
			       s->cfg[st].grid.z == s->cfg[0].grid.z && s->cfg[st].zc == s->cfg[0].zc;
		s->dataflow = same;
		if (same) {
			const size_t n = (size_t)s->cfg[0].grid.x * s->cfg[0].grid.y * s->cfg[0].grid.z;
			for (int st = 0; st < 5; st++) {
				CUS(cudaMalloc(&s->done[st], n * sizeof(unsigned long long)));
				CUS(cudaMemsetAsync(s->done[st], 0, n * sizeof(unsigned long long), s->stream));
			}
		}
	}

	CUS(cudaMalloc(&s->ctl, 2 * sizeof(StepCtl)));             // two blocks: speculative stage 1 alternates between them
	CUS(cudaMemsetAsync(s->ctl, 0, 2 * sizeof(StepCtl), s->stream));
	CUS(cudaHostAlloc(&s->ctl_host, 2 * sizeof(StepCtl), cudaHostAllocDefault));
	CUS(cudaHostAlloc(&s->slot, sizeof(HostSlot), cudaHostAllocMapped));
	memset(s->slot, 0, sizeof(HostSlot));
	CUS(cudaHostGetDevicePointer(&s->slot_dev, s->slot, 0));
	{
		const PowGate &pg = pow_gate();
		s->dev_pow = pg.usable && !(flags & PFT_FLAG_HOST_POW);
		if (pg.usable) {     // (the tables are uploaded even with PFT_FLAG_HOST_POW: pft_debug_pow checks the device twin)
			CUS(cudaMalloc(&s->tab_dev, sizeof(PowTables)));
			CUS(cudaMemcpyAsync(s->tab_dev, &pg.tab, sizeof(PowTables), cudaMemcpyHostToDevice, s->stream));
		} else if (env_int("PFT_VERBOSE", 0))
			fprintf(stderr, "pft_b200: device step-size controller off, the host supplies pow() per attempt: %s\n", pg.why);
	}
	CUS(cudaStreamSynchronize(s->stream));

	memset(&s->st, 0, sizeof s->st);
	{
		const StageCfg &cf = s->cfg[4];          // the dominant stage
		s->st.tile_x = cf.TX; s->st.tile_y = cf.TY * cf.R; s->st.z_chunk = cf.zc; s->st.threads = cf.TX * cf.TY;
		s->st.grid_blocks = cf.grid.x * cf.grid.y * cf.grid.z;
		s->st.kernel_generation = cf.ver; s->st.rows_per_thread = cf.R;
	}
	s->st.bytes_per_field = (long)fb;
	s->spec_ok = want_spec && s->dataflow && s->dev_pow;
	s->spec_earlyout = env_int("PFT_SPEC_EARLYOUT", 1) != 0;
	s->st.device_pow = s->dev_pow ? 1 : 0;
	s->st.runahead = s->depth;
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
	if (s->ipc)
		for (int r = 0; r < PFT_MAX_RANKS; r++)
			if (s->peer_base[r] && s->peer_base[r] != s->slab) cudaIpcCloseMemHandle(s->peer_base[r]);
	if (s->slab) cudaFree(s->slab);
	if (s->halo_tickets) cudaFree(s->halo_tickets);
	if (s->halo_published) cudaFree(s->halo_published);
	for (int st = 0; st < 5; st++) if (s->done[st]) cudaFree(s->done[st]);
	if (s->stage_buf) cudaFree(s->stage_buf);
	if (s->ctl) cudaFree(s->ctl);
	if (s->tab_dev) cudaFree(s->tab_dev);
	if (s->ctl_host) cudaFreeHost(s->ctl_host);
	if (s->slot) cudaFreeHost(s->slot);
	if (s->ev0) cudaEventDestroy(s->ev0);
	if (s->ev1) cudaEventDestroy(s->ev1);
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

// planes klo .. khi-1 (klo = -1 / khi = n3+1 reach into the z-ghost planes of both layouts)
static void repack(pft_solver *s, double *padded_var, double *compact, int dir, int klo, int khi)
{
	const Geom &g = s->g;
	long long total = (long long)g.n1 * g.n2 * (khi - klo);
	int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
	PFT_LAUNCH(repack_kernel, dim3(blocks), dim3(256), 0, s->stream, padded_var, compact, g.n1, g.n2, klo, khi, g.px, g.plane,
	           s->N1, (long long)s->N1 * s->N2, PFT_BT, dir);
	s->st.kernel_launches++;
}

static int refresh_halos(pft_solver *s, unsigned mask)
{
	if (s->cs.nranks <= 1 || s->group) return 0;
	int rc = 0;
	if (mask & (PFT_FIELD_U | PFT_FIELD_P))
		rc = exchange_pair(s, (mask & PFT_FIELD_U) ? s->arr[s->cur ? A_XN : A_X][0] : nullptr, (mask & PFT_FIELD_P) ? s->arr[s->cur ? A_XN : A_X][1] : nullptr);
	if (!rc && (mask & PFT_FIELD_GL)) rc = exchange_pair(s, s->gl, nullptr);
	return rc;
}

static double *field_ptr(pft_solver *s, int q)
{
	return q == 2 ? s->gl : s->arr[s->cur ? A_XN : A_X][q];
}

// One variable of a padded host array -> device.  `host_var` points at padded plane 0 of that variable as this
// slab sees it; lo / hi: also take the plane below / above the slab (a neighbouring slab's boundary rows of the
// SAME host array: in-process groups) into the z-ghost plane.  Only the planes that are used travel.
static int upload_var(pft_solver *s, const double *host_var, int q, int lo, int hi)
{
	const long long rowsize = (long long)s->N1 * s->N2, SIZE = s->padded / PFT_NVAR;
	const long long p0 = PFT_BT - lo, np = s->g.n3 + lo + hi;
	double *st = s->stage_buf + q * SIZE;
	CU(cudaMemcpyAsync(st + p0 * rowsize, host_var + p0 * rowsize, (size_t)(np * rowsize) * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	repack(s, st, field_ptr(s, q), 0, -lo, s->g.n3 + hi);
	s->stage_shell[q] = true;
	return 0;
}

// device -> the interior nodes of one variable of a padded host array.  The x/y ghost nodes inside the copied
// planes come from the staging image: the ones of the caller's last upload of that variable if there was one
// (one-shot), otherwise they are fetched from the host array first, so that ghost nodes are never altered.
static int download_var(pft_solver *s, double *host_var, int q)
{
	const long long rowsize = (long long)s->N1 * s->N2, SIZE = s->padded / PFT_NVAR;
	const long long p0 = PFT_BT, np = s->g.n3;
	double *st = s->stage_buf + q * SIZE;
	if (!s->stage_shell[q])
		CU(cudaMemcpyAsync(st + p0 * rowsize, host_var + p0 * rowsize, (size_t)(np * rowsize) * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	repack(s, st, field_ptr(s, q), 1, 0, s->g.n3);
	CU(cudaMemcpyAsync(host_var + p0 * rowsize, st + p0 * rowsize, (size_t)(np * rowsize) * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	s->stage_shell[q] = false;
	return 0;
}

extern "C" int pft_upload_fields(pft_solver *s, const double *x, unsigned mask)
{
	if (!s || !x) return fail(PFT_EINVAL, "null argument");
	if (s->group) return fail(PFT_ESTATE, "members of an in-process group are filled through pft_group_upload");
	CU(cudaSetDevice(s->device));
	int rc = ensure_stage(s);
	if (rc) return rc;
	const long long SIZE = s->padded / PFT_NVAR;
	for (int q = 0; q < PFT_NVAR; q++)
		if ((mask >> q) & 1u)
			if ((rc = upload_var(s, x + q * SIZE, q, 0, 0))) return rc;
	CU(cudaGetLastError());
	rc = refresh_halos(s, mask);
	if (rc) return rc;
	CU(cudaStreamSynchronize(s->stream));
	s->k1_valid = false;
	return 0;
}

extern "C" int pft_download_fields(pft_solver *s, double *x, unsigned mask)
{
	if (!s || !x) return fail(PFT_EINVAL, "null argument");
	CU(cudaSetDevice(s->device));
	int rc = ensure_stage(s);
	if (rc) return rc;
	const long long SIZE = s->padded / PFT_NVAR;
	for (int q = 0; q < PFT_NVAR; q++)
		if ((mask >> q) & 1u)
			if ((rc = download_var(s, x + q * SIZE, q))) return rc;
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(s->stream));
	return 0;
}

extern "C" int pft_upload(pft_solver *s, const double *x) { return pft_upload_fields(s, x, PFT_FIELD_ALL); }
extern "C" int pft_download(pft_solver *s, double *x) { return pft_download_fields(s, x, PFT_FIELD_ALL); }

extern "C" int pft_host_register(void *ptr, unsigned long long bytes)
{
	if (!ptr || !bytes) return fail(PFT_EINVAL, "null argument");
#ifndef PFT_CUDA_EMU
	cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
	if (e != cudaSuccess) { cudaGetLastError(); return fail(PFT_ECUDA, "cudaHostRegister(%llu bytes): %s", bytes, cudaGetErrorString(e)); }
#endif
	return 0;
}

extern "C" int pft_host_unregister(void *ptr)
{
	if (!ptr) return fail(PFT_EINVAL, "null argument");
#ifndef PFT_CUDA_EMU
	cudaError_t e = cudaHostUnregister(ptr);
	if (e != cudaSuccess) { cudaGetLastError(); return fail(PFT_ECUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
#endif
	return 0;
}

static int dense_io(pft_solver *s, double *host, double *compact, int dir)
{
	const Geom &g = s->g;
	long long total = (long long)g.n1 * g.n2 * g.n3;
	int rc = ensure_stage(s);
	if (rc) return rc;
	s->stage_shell[0] = false;      // the dense image goes through the head of the staging buffer
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
	if (u && (rc = dense_io(s, const_cast<double *>(u), field_ptr(s, 0), 0))) return rc;
	if (p && (rc = dense_io(s, const_cast<double *>(p), field_ptr(s, 1), 0))) return rc;
	if (gl && (rc = dense_io(s, const_cast<double *>(gl), s->gl, 0))) return rc;
	rc = refresh_halos(s, (u ? PFT_FIELD_U : 0u) | (p ? PFT_FIELD_P : 0u) | (gl ? PFT_FIELD_GL : 0u));
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
	if (u && (rc = dense_io(s, u, field_ptr(s, 0), 1))) return rc;
	if (p && (rc = dense_io(s, p, field_ptr(s, 1), 1))) return rc;
	if (gl && (rc = dense_io(s, gl, s->gl, 1))) return rc;
	return 0;
}

// which: 0 x, 1 K1, 2 aux3, 3 K3, 4 K4, 5 x_new (the candidate of the last attempt)
extern "C" int pft_debug_fetch(pft_solver *s, int which, double *u, double *p)
{
	if (!s || which < 0 || which >= A_COUNT) return fail(PFT_EINVAL, "bad array id %d", which);
	CU(cudaSetDevice(s->device));
	int pair = which;
	if (which == A_X) pair = s->cur ? A_XN : A_X;
	if (which == A_XN) pair = s->cur ? A_X : A_XN;
	if (which == A_K1B && !s->arr[A_K1B][1]) return fail(PFT_EINVAL, "bad array id %d", which);
	if (s->spec && which == A_K1) pair = s->cur ? A_K1B : A_K1;       // K1 of the state at rest (one buffer per pair)
	int rc;
	if (u) {
		if (s->arr[pair][0]) { if ((rc = dense_io(s, u, s->arr[pair][0], 1))) return rc; }
		else memset(u, 0, sizeof(double) * (size_t)s->g.n1 * s->g.n2 * s->g.n3);     // K_u == 0 in calc_mode 10/11
	}
	if (p && (rc = dense_io(s, p, s->arr[pair][1], 1))) return rc;
	return 0;
}

extern "C" int pft_debug_pow(pft_solver *s, const double *x, double y, double *out, int n)
{
	if (!s || !x || !out || n < 0) return fail(PFT_EINVAL, "bad argument");
	if (!s->tab_dev) return fail(PFT_ESTATE, "the pow tables of the host C library are not available: %s", pow_gate().why);
	CU(cudaSetDevice(s->device));
	double *d = nullptr;
	CU(cudaMalloc(&d, sizeof(double) * 2 * (size_t)std::max(n, 1)));
	CU(cudaMemcpyAsync(d, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
	PFT_LAUNCH(pow_twin_kernel, dim3(148 * 4), dim3(256), 0, s->stream, s->tab_dev, d, y, d + n, n);
	s->st.kernel_launches++;
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out, d + n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	cudaFree(d);
	return 0;
}

extern "C" int pft_rasterise(pft_solver *s, const double *xyz_raw, int count)
{
	if (!s || (count > 0 && !xyz_raw) || count < 0) return fail(PFT_EINVAL, "bad sphere list");
	if (count == 0) return 0;
	CU(cudaSetDevice(s->device));
	const pft_case &cs = s->cs;
	const Geom &g = s->g;
	// centres in grid coordinates (equation.c:480-482), sorted into uniform bins of the culling radius:
	// a node only visits the 27 bins around it.  Bit-identical to testing every sphere: a sphere that is
	// skipped lies further away than R + 19.5/s, where its profile is exactly 0 (see rasterise_kernel).
	std::vector<double> b(3 * (size_t)count);
	for (int q = 0; q < count; q++) {
		b[3 * (size_t)q] = xyz_raw[3 * q] * cs.param[P_beads_scaling] + cs.param[P_beads_offset_x];
		b[3 * (size_t)q + 1] = xyz_raw[3 * q + 1] * cs.param[P_beads_scaling] + cs.param[P_beads_offset_y];
		b[3 * (size_t)q + 2] = xyz_raw[3 * q + 2] * cs.param[P_beads_scaling] + cs.param[P_beads_offset_z];
	}
	const double sfac = 0.5 / cs.param[P_xi_gl];
	const double cut = cs.param[P_ball_radius] + 19.5 / sfac;
	RasterBins rb;
	double lo[3] = {0.0, 0.0, cs.L3 * (0.5 + cs.first_row) / cs.total_n3}, hi[3] = {cs.L1, cs.L2, cs.L3 * (0.5 + cs.first_row + g.n3 - 1) / cs.total_n3};
	// bin edge >= cut; at most 64 bins per direction
	for (int d = 0; d < 3; d++) {
		rb.org[d] = lo[d];
		const double ext = std::max(hi[d] - lo[d], 1e-300);
		int nb = (int)std::min(64.0, std::max(1.0, floor(ext / cut)));
		if (env_int("PFT_RASTER_ONE_BIN", 0)) nb = 1;      // A/B and test: every node tests every sphere
		rb.nb[d] = nb;
		rb.inv[d] = nb / ext;
	}
	const int nbins = rb.nb[0] * rb.nb[1] * rb.nb[2];
	auto bin_of = [&](const double *c, int d) {
		double f = (c[d] - rb.org[d]) * rb.inv[d];
		int i = (f < 0) ? 0 : (f >= rb.nb[d] ? rb.nb[d] - 1 : (int)f);      // spheres outside the block go to the edge bins ...
		return i;
	};
	// ... but only if they can reach the block at all
	std::vector<int> keep;
	for (int q = 0; q < count; q++) {
		bool reach = true;
		for (int d = 0; d < 3; d++) reach = reach && b[3 * (size_t)q + d] >= lo[d] - cut && b[3 * (size_t)q + d] <= hi[d] + cut;
		if (reach) keep.push_back(q);
	}
	std::vector<int> start(nbins + 1, 0);
	std::vector<int> which(keep.size());
	for (size_t m = 0; m < keep.size(); m++) {
		const double *c = &b[3 * (size_t)keep[m]];
		which[m] = (bin_of(c, 2) * rb.nb[1] + bin_of(c, 1)) * rb.nb[0] + bin_of(c, 0);
		start[which[m] + 1]++;
	}
	for (int i = 0; i < nbins; i++) start[i + 1] += start[i];
	std::vector<double> sorted(3 * std::max<size_t>(keep.size(), 1));
	{
		std::vector<int> fill(start.begin(), start.end() - 1);
		for (size_t m = 0; m < keep.size(); m++) {
			const int dst = fill[which[m]]++;
			for (int d = 0; d < 3; d++) sorted[3 * (size_t)dst + d] = b[3 * (size_t)keep[m] + d];
		}
	}
	double *dc = nullptr; int *ds = nullptr;
	CU(cudaMalloc(&dc, sorted.size() * sizeof(double)));
	CU(cudaMalloc(&ds, start.size() * sizeof(int)));
	CU(cudaMemcpyAsync(dc, sorted.data(), sorted.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	CU(cudaMemcpyAsync(ds, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
	long long total = (long long)g.n1 * g.n2 * g.n3;
	const int threads = 256;
	long long blocks = (total + threads - 1) / threads;
	PFT_LAUNCH(rasterise_kernel, dim3((unsigned)blocks), dim3(threads), 0, s->stream,
		s->gl, dc, ds, rb, g.n1, g.n2, g.n3, g.px, g.plane,
		cs.first_row, cs.total_n3, cs.L1, cs.L2, cs.L3, cs.param[P_xi_gl], cs.param[P_ball_radius]);
	s->st.kernel_launches++;
	CU(cudaGetLastError());
	int rc = (s->cs.nranks > 1 && !s->group) ? exchange_pair(s, s->gl, nullptr) : 0;
	CU(cudaStreamSynchronize(s->stream));
	cudaFree(dc); cudaFree(ds);
	s->k1_valid = false;            // K1 = f(t, x) was evaluated with the old glass field
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
	PFT_LAUNCH(ice_count_kernel, dim3(148 * 4), dim3(256), 0, s->stream, field_ptr(s, 1), g.n1, g.n2, g.n3, g.px, g.plane, d);
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
	// equation.c:96-111
	const double dir = t < s->cs.param[P_phase_switch_time] ? s->cs.param[P_top_temp1] : s->cs.param[P_top_temp2];
	rc = launch_stage(s, 1, &dir);
	if (rc) return rc;
	s->st.rhs_evals++;
	const long long SIZE = s->padded / PFT_NVAR, rowsize = (long long)s->N1 * s->N2;
	// interior nodes of dw: through the staging image, ghost nodes of dw untouched
	for (int q = 0; q < 2; q++) {
		if (q == 0 && !s->heat) continue;
		double *st = s->stage_buf + q * SIZE;
		CU(cudaMemcpyAsync(st + PFT_BT * rowsize, dw + q * SIZE + PFT_BT * rowsize, (size_t)(s->g.n3 * rowsize) * sizeof(double), cudaMemcpyHostToDevice, s->stream));
		repack(s, st, s->arr[A_K1][q], 1, 0, s->g.n3);
		CU(cudaMemcpyAsync(dw + q * SIZE + PFT_BT * rowsize, st + PFT_BT * rowsize, (size_t)(s->g.n3 * rowsize) * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
		s->stage_shell[q] = false;
	}
	CU(cudaStreamSynchronize(s->stream));
	// du/dt == 0 in calc_mode 10/11 and dgl/dt == 0 always (equation.c:708,731): interior rows only
	{
		const int N1 = s->N1;
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

static inline double host_ms()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
	__builtin_ia32_pause();
#elif defined(__aarch64__)
	asm volatile("yield");
#endif
}

// Waits until the device has advanced the sequence word to `target`.  Polls briefly, then SLEEPS between looks:
// with attempts enqueued ahead the host has nothing to do, and a core spinning per GPU buys nothing.
// spin_first: polls before the first sleep (depth-1 modes wait for every single attempt and want the low latency).
static int wait_seq(pft_solver *s, const volatile unsigned long long *p, unsigned long long target, unsigned spin_first)
{
	unsigned long polls = 0;
	while (*p < target) {
		if (++polls <= spin_first) { cpu_relax(); continue; }
		struct timespec ts = {0, 20000};
		nanosleep(&ts, nullptr);
		if ((polls & 0xff) == 0) {
			cudaError_t e = cudaStreamQuery(s->stream);
			if (e != cudaSuccess && e != cudaErrorNotReady)
				return fail(PFT_ECUDA, "device fault while waiting for an attempt to finish: %s", cudaGetErrorString(e));
			if (e == cudaSuccess && *p < target)
				return fail(PFT_ESTATE, "stream idle but the attempt never published its result (sequence %llu < %llu)", (unsigned long long)*p, target);
		}
	}
	__atomic_thread_fence(__ATOMIC_ACQUIRE);
	return 0;
}

// the launches of ONE attempt: [S1 (skipped on the device while K1 is valid), S2, S3, S4, S5 (+ the decision)];
// with speculative stage 1: [S1 only at solve entry and only if K1 is not valid], S2, S3, S4, S5 (+ the decision), S1'
static int enqueue_attempt(pft_solver *s, int live_cur, bool entry_s1)
{
	const bool nccl_mode = s->cs.nranks > 1 && !s->p2p;
	const int outs[5] = {A_K1, A_A3, A_K3, A_K4, live_cur ? A_X : A_XN};
	int rc;
	s->lseq++;
	if (s->spec) {
		if (entry_s1 && (rc = launch_stage(s, 1))) return rc;
		for (int stage = 2; stage <= 5; stage++)
			if ((rc = launch_stage(s, stage))) return rc;
		return launch_stage(s, 1, nullptr, true);
	}
	for (int stage = 1; stage <= 5; stage++) {
		if ((rc = launch_stage(s, stage))) return rc;
		if (nccl_mode && (rc = exchange_pair(s, s->heat ? s->arr[outs[stage - 1]][0] : nullptr, s->arr[outs[stage - 1]][1]))) return rc;
	}
	if (nccl_mode) {
		// the global error maximum: MPI_Allreduce(MAX), RK_MPI_SAsolver_hybrid2.c:572
		double *red = (double *)((char *)s->ctl + offsetof(StepCtl, red));
		ncclResult_t r_ = g_nccl.AllReduce(red, red, 2, ncclDouble, ncclMax, s->comm, s->stream);
		if (r_ != ncclSuccess) return fail(PFT_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r_));
		if (s->dev_pow) PFT_LAUNCH(ctl_decide_kernel, dim3(1), dim3(1), 0, s->stream, s->ctl, s->slot_dev, (const PowTables *)s->tab_dev, 0.0, 0);
		else PFT_LAUNCH(ctl_publish_eps_kernel, dim3(1), dim3(1), 0, s->stream, s->ctl, s->slot_dev);
		s->st.kernel_launches++;
	}
	return 0;
}

extern "C" int pft_solve(pft_solver *s, pft_system *sys, double final_time)
{
	if (!s || !sys) { fail(PFT_EINVAL, "null argument"); return PFT_SOLVE_BADSYSTEM; }
	if (sys->delta <= 0) { fail(PFT_EINVAL, "delta <= 0"); return PFT_SOLVE_BADSYSTEM; }        // hybrid2.c:286
	if (cudaSetDevice(s->device) != cudaSuccess) return fail(PFT_ECUDA, "cudaSetDevice failed");
	if (s->cs.nranks > 1 && !s->p2p && !s->have_comm) return fail(PFT_ESTATE, "nranks > 1 but pft_comm_init was not called");
	if (!s->params_ok && build_params(s)) return PFT_ECUDA;

	const bool nccl_mode = s->cs.nranks > 1 && !s->p2p;
	const bool host_pow = !s->dev_pow;
	const bool callback = sys->service != nullptr;
	const int depth = (nccl_mode || host_pow || callback || s->profile) ? 1 : s->depth;
	const unsigned spin_first = depth == 1 ? 200000u : 64u;
	int rc = 0, ret = PFT_SOLVE_DONE;
	bool interrupted = false;

	// ---- the control block of this call (RK_MPI_SAsolver_hybrid2.c:316-325), written by the host while the device is idle
	// (speculative stage 1: the first attempt of this call is number lseq+1 and reads block (lseq+1) & 1)
	const int blk = s->spec ? (int)((s->lseq + 1ull) & 1ull) : 0;
	StepCtl &c = s->ctl_host[blk];
	memset(s->ctl_host, 0, 2 * sizeof(StepCtl));
	c.t = sys->t; c.h = sys->h;
	c.final_time = final_time; c.h_min = sys->h_min; c.delta = sys->delta;
	c.delta_local = sys->delta_local ? 1 : 0; c.handle_nan = sys->handle_nan ? 1 : 0;
	c.recompute_k1 = (s->flags & PFT_FLAG_RECOMPUTE_K1) ? 1 : 0;
	c.steps = sys->steps; c.steps_total = sys->steps_total; c.rhs_evals = sys->rhs_evals;
	c.max_attempts = sys->max_attempts;
	c.switch_time = s->cs.param[P_phase_switch_time]; c.top_temp1 = s->cs.param[P_top_temp1]; c.top_temp2 = s->cs.param[P_top_temp2];
	c.cur = s->cur; c.aseq = s->aseq; c.hbase = s->hbase;
	c.spec = s->spec ? 1 : 0;
	c.skip1 = (s->k1_valid && s->k1_t == sys->t && !c.recompute_k1) ? 1 : 0;                      // :373
	ctl_begin(c);
	if (c.skip1 && s->k1_uncounted) c.rhs_evals++;         // that K1 was evaluated by the S1' that closed the previous call
	s->k1_uncounted = false;
	c.ready = s->lseq + 1ull; c.closed_from = ~0ull;
	s->ctl_host[blk ^ 1].closed_from = ~0ull;
	sys->last_nan = 0;                                                                           // :316
	const long rhs0 = sys->rhs_evals, steps0 = sys->steps;

	if (cudaMemcpyAsync(s->ctl, s->ctl_host, 2 * sizeof(StepCtl), cudaMemcpyHostToDevice, s->stream) != cudaSuccess) return fail(PFT_ECUDA, "control block upload failed");
	cudaEventRecord(s->ev0, s->stream);

	const unsigned long long seq0 = s->aseq;
	unsigned long long done = seq0, enq = seq0;
	const unsigned long long budget = sys->max_attempts > 0 ? (unsigned long long)sys->max_attempts : ~0ull;
	int live_cur = s->cur;                 // depth-1 modes: the host follows StepCtl.cur attempt by attempt
	int status = ST_RUNNING;
	HostSlot last;
	memset(&last, 0, sizeof last);
	double t_mark = host_ms();

	// a call whose first step already reaches final_time (fixed-step mode: :324 set CMD_FINISHED) most likely needs ONE
	// attempt: its window opens up from one.  Any other call starts with the full window -- launches beyond the end
	// return at once, and the device never waits for the host while the window opens.
	// (PFT_SLOW_OPEN=1: every call opens its window from one -- for A/B measurements)
	const unsigned long long open0 = ((c.command & CMD_FINISHED) || env_int("PFT_SLOW_OPEN", 0)) ? 1ull : (unsigned long long)depth;
	while (status == ST_RUNNING) {
		// ---- keep the launch queue filled
		const unsigned long long window = std::min<unsigned long long>((unsigned long long)depth, open0 + (done - seq0));
		while (enq - done < window && enq - seq0 < budget) {
			if ((rc = enqueue_attempt(s, live_cur, enq == seq0 && !c.skip1))) goto device_error;
			enq++;
		}
		if (enq == done) { rc = fail(PFT_ESTATE, "attempt budget exhausted but the device did not stop"); goto device_error; }
		{ const double now = host_ms(); s->st.host_launch_ms += now - t_mark; t_mark = now; }

		if (host_pow) {
			// the device published the global error maximum and waits for pow() (pft_pow.h explains when)
			if ((rc = wait_seq(s, &s->slot->seq_eps, done + 1, spin_first))) goto device_error;
			double max_eps = s->slot->eps;
			if (s->slot->delta_local) max_eps *= s->slot->h3_attempt;                              // :578
			const double pw = (max_eps > 0.0) ? pow(sys->delta / max_eps, 0.2) : 0.0;              // :580
			PFT_LAUNCH(ctl_decide_kernel, dim3(1), dim3(1), 0, s->stream, s->ctl, s->slot_dev, (const PowTables *)nullptr, pw, 1);
			s->st.kernel_launches++;
		}
		if ((rc = wait_seq(s, &s->slot->seq, done + 1, spin_first))) goto device_error;
		{ const double now = host_ms(); s->st.host_wait_ms += now - t_mark; t_mark = now; }
		const unsigned long long newdone = s->slot->seq;
		__atomic_thread_fence(__ATOMIC_ACQUIRE);
		status = *(volatile int *)&s->slot->status;

		if (depth == 1) {
			// nothing is in flight: the slot is stable.  Follow the attempt the way the reference's master does.
			last = *s->slot;
			live_cur = last.cur;
			status = last.status;
			if (callback && last.accept) {                                                       // :676-705
				sys->t = last.t; sys->h = last.h_attempt;
				sys->steps = (long)last.steps; sys->steps_total = (long)last.steps_total;
				const int stop = sys->service(final_time, last.t, last.h_attempt, (long)last.steps, (long)last.steps_total, sys->service_ctx);
				if (last.status == ST_DONE) { /* FINISHED wins over BREAK (:695 before :697) */ }
				else if (stop) { interrupted = true; done = newdone; break; }
			}
			if (callback && (last.command & CMD_NEXTFINISH) && !(last.command & CMD_NAN) && last.status != ST_DONE) sys->h = last.new_h;   // :754
			if (callback && last.status == ST_BUDGET && !(last.command_next & CMD_FINISHED)) sys->h = last.h_next;
		}
		done = newdone;
	}

	// ---- exit: drain the queue (launches beyond the end return at once), read the final state
	cudaEventRecord(s->ev1, s->stream);
	if (cudaStreamSynchronize(s->stream) != cudaSuccess) { rc = fail(PFT_ECUDA, "stream sync at solve exit: %s", cudaGetErrorString(cudaGetLastError())); goto device_error; }
	if (s->profile) drain_timers(s);
	{
		float ms = 0.f;
		cudaEventElapsedTime(&ms, s->ev0, s->ev1);
		s->st.device_ms_total = ms;
	}
	last = *s->slot;
	{
		const long n_att = (long)(last.seq - seq0);
		s->st.attempts += n_att;
		s->st.accepted += (long)last.steps - steps0;
		s->st.noop_launches += (long)(enq - last.seq) * 5;
		// (stage launches that returned at once are not right-hand-side evaluations)
		long rhs = (long)last.rhs_evals;
		// stage 1 of the next attempt was counted but never enqueued (with speculative stage 1 it has run: S1')
		if (interrupted && last.status == ST_RUNNING && !s->spec) rhs -= 1;
		s->st.rhs_evals += rhs - rhs0;
		sys->rhs_evals = rhs;
	}
	if (!callback) { if (last.sys_h_set) sys->h = last.sys_h; }
	else if (interrupted) sys->h = last.new_h;                                                   // :701
	sys->t = last.t;                                                                             // :768 / :700 / :634
	sys->steps = (long)last.steps; sys->steps_total = (long)last.steps_total;
	sys->last_nan = last.last_nan;
	s->cur = last.cur;
	s->aseq = last.seq;
	if (last.seq > seq0) s->hbase = last.hbase;
	else s->hbase = c.hbase;                                                // (no attempt ran: the entry launch took its number)
	if (s->spec && last.seq > seq0) {
		// K1 of the state at rest always exists: kept after a rejected attempt, evaluated by S1' after an accepted one
		s->k1_valid = true;
		s->k1_uncounted = last.accept && last.status != ST_RUNNING;         // (ctl_prepare counts it only when the run goes on)
	} else
		s->k1_valid = (last.seq > seq0) ? (last.accept == 0) : s->k1_valid;     // after a rejected attempt K1 == f(t, x) still holds
	s->k1_t = last.t;
	if (interrupted) ret = PFT_SOLVE_INTERRUPTED;
	else if (last.status == ST_BUDGET) ret = PFT_SOLVE_BUDGET;
	else if (last.status == ST_NAN) ret = PFT_SOLVE_NAN;
	return ret;

device_error:
	cudaStreamSynchronize(s->stream);
	drain_timers(s);
	// what the device did to the control block is unknown: bring the host's mirror back in step with it
	if (s->slot->seq >= s->aseq) { s->aseq = s->slot->seq; s->cur = s->slot->cur; if (s->slot->seq > seq0) s->hbase = s->slot->hbase; else s->hbase = c.hbase; }
	s->k1_valid = false; s->k1_uncounted = false;
	return rc;
}

// ------------------------------------------------------------------------------------------
// one host process, several GPUs

extern "C" int pft_group_size(const pft_group *g) { return g ? g->n : 0; }
extern "C" pft_solver *pft_group_member(pft_group *g, int r) { return (g && r >= 0 && r < g->n) ? g->member[r] : nullptr; }

static void group_free_members(pft_group *g)
{
	for (pft_solver *m : g->member) pft_destroy(m);
	g->member.clear();
}

extern "C" int pft_group_destroy(pft_group *g)
{
	if (!g) return 0;
	group_free_members(g);
	delete g;
	return 0;
}

// create the member solvers for g->planes and wire them to each other
static int group_build_members(pft_group *g)
{
	const pft_case *whole = &g->whole;
	const int ngpus = g->n;
	int first = whole->first_row, deepest = 0;
	for (int r = 0; r < ngpus; r++) deepest = std::max(deepest, g->planes[r]);
	for (int r = 0; r < ngpus; r++) {
		pft_case c = *whole;
		c.rank = r; c.nranks = ngpus;
		c.n3 = g->planes[r];
		c.first_row = first;
		c.max_n3 = deepest;
		first += c.n3;
		// (total_n3 stays the whole grid's: the z grid spacing comes from it, equation.c:607)
		pft_solver *m = nullptr;
		int rc = pft_create(&c, g->devices[r], g->flags, &m);
		if (rc) { group_free_members(g); return rc; }
		m->group = g;
		g->member.push_back(m);
	}
	if (ngpus > 1) {
#ifndef PFT_CUDA_EMU
		for (int r = 0; r < ngpus; r++)
			for (int q = 0; q < ngpus; q++) {
				if (q == r || g->member[q]->device == g->member[r]->device) continue;
				int can = 0;
				cudaDeviceCanAccessPeer(&can, g->member[r]->device, g->member[q]->device);
				if (!can) { int a_ = g->member[r]->device, b_ = g->member[q]->device; group_free_members(g); return fail(PFT_ENODEVICE, "device %d cannot access device %d: the in-process group needs peer access (NVLink)", a_, b_); }
				cudaSetDevice(g->member[r]->device);
				cudaError_t e = cudaDeviceEnablePeerAccess(g->member[q]->device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { group_free_members(g); return fail(PFT_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); }
				cudaGetLastError();
			}
#endif
		for (int r = 0; r < ngpus; r++) {
			pft_solver *m = g->member[r];
			for (int q = 0; q < ngpus; q++) { m->peer_base[q] = g->member[q]->slab; m->peer_n3[q] = g->member[q]->cs.n3; }
			m->p2p = true; m->ipc = false;
			m->params_ok = false;
		}
	}
	return 0;
}

extern "C" int pft_group_create_planes(const pft_case *whole, int ngpus, const int *devices, const int *planes, unsigned flags, pft_group **out)
{
	if (!whole || !out || ngpus < 1) return fail(PFT_EINVAL, "bad argument");
	*out = nullptr;
	if (whole->nranks != 1 || whole->rank != 0) return fail(PFT_EINVAL, "pft_group_create takes the whole block of a single-rank run");
	if (ngpus > PFT_MAX_RANKS) return fail(PFT_EINVAL, "at most %d GPUs per group", PFT_MAX_RANKS);
	if (whole->n3 / ngpus < 1) return fail(PFT_EINVAL, "%d planes cannot be cut into %d slabs", whole->n3, ngpus);
	pft_group *g = new (std::nothrow) pft_group();
	if (!g) return fail(PFT_ENOMEM, "out of host memory");
	g->n = ngpus; g->whole = *whole; g->flags = flags;
	g->user_service = nullptr; g->user_ctx = nullptr;
	g->cb_seq.store(-1); g->cb_val.store(0);
	g->balance_tried = (ngpus == 1) || env_int("PFT_BALANCE", 1) == 0 || planes != nullptr;
	int sum = 0;
	for (int r = 0; r < ngpus; r++) {
		g->devices.push_back(devices ? devices[r] : r);
		g->planes.push_back(planes ? planes[r] : slab_n3(whole->n3, ngpus, r));          // intertrack.c:1780-1787
		sum += g->planes[r];
		if (g->planes[r] < 2 && ngpus > 1) { delete g; return fail(PFT_EINVAL, "slab %d would hold fewer than 2 planes", r); }
	}
	if (sum != whole->n3) { delete g; return fail(PFT_EINVAL, "the slabs hold %d planes, the block has %d", sum, whole->n3); }
#ifndef PFT_CUDA_EMU
	// one slab per GPU: the CTAs of a slab spin on flags that the neighbouring slab's CTAs write, which must not compete for the same SMs
	for (int r = 0; r < ngpus; r++)
		for (int q = 0; q < r; q++)
			if (g->devices[q] == g->devices[r]) { int d_ = g->devices[r]; delete g; return fail(PFT_EINVAL, "device %d appears twice in the group: one slab per GPU", d_); }
#endif
	int rc = group_build_members(g);
	if (rc) { delete g; return rc; }
	*out = g;
	return 0;
}

extern "C" int pft_group_create(const pft_case *whole, int ngpus, const int *devices, unsigned flags, pft_group **out)
{
	return pft_group_create_planes(whole, ngpus, devices, nullptr, flags, out);
}

extern "C" int pft_group_planes(const pft_group *g, int *planes)
{
	if (!g || !planes) return fail(PFT_EINVAL, "null argument");
	for (int r = 0; r < g->n; r++) planes[r] = g->planes[r];
	return g->n;
}

// the slabs of ONE padded host array: slab r's padded plane 0 is the whole block's padded plane (first_row_r - first_row_0)
extern "C" int pft_group_upload(pft_group *g, const double *x, unsigned mask)
{
	if (!g || !x) return fail(PFT_EINVAL, "null argument");
	const pft_case &w = g->whole;
	const long long rowsize = (long long)(w.n1 + 2 * PFT_BT) * (w.n2 + 2 * PFT_BT), SIZEW = rowsize * (w.n3 + 2 * PFT_BT);
	for (int r = 0; r < g->n; r++) {
		pft_solver *m = g->member[r];
		CU(cudaSetDevice(m->device));
		int rc = ensure_stage(m);
		if (rc) return rc;
		const long long row0 = m->cs.first_row - w.first_row;
		for (int q = 0; q < PFT_NVAR; q++)
			if ((mask >> q) & 1u)
				if ((rc = upload_var(m, x + q * SIZEW + row0 * rowsize, q, r > 0, r < g->n - 1))) return rc;
		CU(cudaGetLastError());
		m->k1_valid = false;
	}
	for (int r = 0; r < g->n; r++) { CU(cudaSetDevice(g->member[r]->device)); CU(cudaStreamSynchronize(g->member[r]->stream)); }
	return 0;
}

extern "C" int pft_group_download(pft_group *g, double *x, unsigned mask)
{
	if (!g || !x) return fail(PFT_EINVAL, "null argument");
	const pft_case &w = g->whole;
	const long long rowsize = (long long)(w.n1 + 2 * PFT_BT) * (w.n2 + 2 * PFT_BT), SIZEW = rowsize * (w.n3 + 2 * PFT_BT);
	for (int r = 0; r < g->n; r++) {
		pft_solver *m = g->member[r];
		CU(cudaSetDevice(m->device));
		int rc = ensure_stage(m);
		if (rc) return rc;
		const long long row0 = m->cs.first_row - w.first_row;
		for (int q = 0; q < PFT_NVAR; q++)
			if ((mask >> q) & 1u)
				if ((rc = download_var(m, x + q * SIZEW + row0 * rowsize, q))) return rc;
		CU(cudaGetLastError());
	}
	for (int r = 0; r < g->n; r++) { CU(cudaSetDevice(g->member[r]->device)); CU(cudaStreamSynchronize(g->member[r]->stream)); }
	return 0;
}

// Service_Callback in a group: the callback belongs to the master (intertrack.c:2173 installs it on rank 0 only);
// every slab stops after the same accepted step, so slab 0's worker asks the user and the others take its answer
// (the reference re-broadcasts the command word after the callback, RK_MPI_SAsolver_hybrid2.c:690).
static int group_service_leader(double final_time, double t, double h, long steps, long steps_total, void *ctx)
{
	pft_group *g = (pft_group *)ctx;
	const int v = g->user_service(final_time, t, h, steps, steps_total, g->user_ctx) ? 1 : 0;
	g->cb_val.store(v, std::memory_order_relaxed);
	g->cb_seq.store(steps_total, std::memory_order_release);
	return v;
}

static int group_service_follower(double, double, double, long, long steps_total, void *ctx)
{
	pft_group *g = (pft_group *)ctx;
	while (g->cb_seq.load(std::memory_order_acquire) != steps_total) cpu_relax();
	return g->cb_val.load(std::memory_order_relaxed);
}

// one worker thread per slab, each the ordinary pft_solve host loop on its own device
static int group_run(pft_group *g, pft_system *sys, double final_time)
{
	if (g->n == 1) return pft_solve(g->member[0], sys, final_time);
	std::vector<pft_system> ss(g->n, *sys);
	std::vector<int> rcs(g->n, 0);
	std::vector<std::string> errs(g->n);
	g->user_service = sys->service; g->user_ctx = sys->service_ctx;
	g->cb_seq.store(-1);
	for (int r = 0; r < g->n; r++) {
		if (sys->service) { ss[r].service = (r == 0) ? group_service_leader : group_service_follower; ss[r].service_ctx = g; }
	}
	std::vector<std::thread> th;
	for (int r = 1; r < g->n; r++)
		th.emplace_back([&, r]() { rcs[r] = pft_solve(g->member[r], &ss[r], final_time); if (rcs[r] < 0) errs[r] = pft_last_error(); });
	rcs[0] = pft_solve(g->member[0], &ss[0], final_time);
	if (rcs[0] < 0) errs[0] = pft_last_error();
	for (auto &t : th) t.join();
	for (int r = 0; r < g->n; r++)
		if (rcs[r] != rcs[0] || ss[r].t != ss[0].t || ss[r].steps_total != ss[0].steps_total) {
			if (rcs[r] < 0 && rcs[r] != PFT_SOLVE_NAN) return fail(rcs[r], "slab %d: %s", r, errs[r].c_str());
			if (rcs[0] < 0 && rcs[0] != PFT_SOLVE_NAN) return fail(rcs[0], "slab 0: %s", errs[0].c_str());
			return fail(PFT_ESTATE, "the slabs of the group disagree (slab %d: code %d t %.17g attempts %ld; slab 0: code %d t %.17g attempts %ld)",
			            r, rcs[r], ss[r].t, ss[r].steps_total, rcs[0], ss[0].t, ss[0].steps_total);
		}
	if (rcs[0] < 0 && rcs[0] != PFT_SOLVE_NAN && rcs[0] != PFT_SOLVE_BADSYSTEM) return fail(rcs[0], "slab 0: %s", errs[0].c_str());
	const pft_service_fn keep_fn = sys->service; void *keep_ctx = sys->service_ctx;
	*sys = ss[0];
	sys->service = keep_fn; sys->service_ctx = keep_ctx;
	return rcs[0];
}

// planes in inverse proportion to the measured cost per plane (every slab keeps >= 2 planes, the sum is exact)
static std::vector<int> balanced_planes(const std::vector<int> &planes, const std::vector<double> &cost)
{
	const int n = (int)planes.size();
	int total = 0;
	double wsum = 0.0;
	std::vector<double> w(n);
	for (int r = 0; r < n; r++) { total += planes[r]; w[r] = planes[r] / cost[r]; wsum += w[r]; }
	std::vector<int> out(n);
	int given = 0;
	for (int r = 0; r < n; r++) { out[r] = std::max(2, (int)floor(total * w[r] / wsum)); given += out[r]; }
	// deal the remainder (or take back an excess) one plane at a time, largest fractional part first
	while (given != total) {
		int best = -1;
		double bestv = 0.0;
		for (int r = 0; r < n; r++) {
			const double frac = total * w[r] / wsum - out[r];
			const double v = (given < total) ? frac : -frac;
			if ((given < total || out[r] > 2) && (best < 0 || v > bestv)) { best = r; bestv = v; }
		}
		if (best < 0) break;
		out[best] += (given < total) ? 1 : -1;
		given += (given < total) ? 1 : -1;
	}
	return out;
}

// *solve_rc: what the calibration attempts did to the integration (PFT_SOLVE_*; the caller's budget is accounted for)
static int group_balance(pft_group *g, pft_system *sys, double final_time, int attempts, int *solve_rc)
{
	*solve_rc = PFT_SOLVE_BUDGET;
	g->balance_tried = true;
	if (g->n < 2 || attempts < 1) return 0;
	// ---- `attempts` attempts with per-stage device timing (events around every launch, one attempt in flight)
	std::vector<pft_stats> before(g->n);
	for (int r = 0; r < g->n; r++) { pft_solver *m = g->member[r]; before[r] = m->st; m->profile = true; m->pdl = false; }
	const long keep_budget = sys->max_attempts;
	sys->max_attempts = (keep_budget > 0) ? std::min<long>(keep_budget, attempts) : attempts;
	const long t0 = sys->steps_total;
	const int rc = group_run(g, sys, final_time);
	const long used = sys->steps_total - t0;
	sys->max_attempts = keep_budget;
	for (int r = 0; r < g->n; r++) { pft_solver *m = g->member[r]; m->profile = env_int("PFT_PROFILE_STAGES", 0) != 0; m->pdl = !m->profile && env_int("PFT_NO_PDL", 0) == 0; }
	*solve_rc = rc;
	if (rc == PFT_SOLVE_BUDGET && !(keep_budget > 0 && used >= keep_budget)) *solve_rc = -1000;      // only the calibration's own budget ran out: go on
	if (rc < 0) return rc == PFT_SOLVE_NAN ? 0 : rc;
	// stages 1-4 (stage 5's time contains the wait for every slab's error maximum): device ms per attempt
	std::vector<double> cost(g->n, 0.0);
	double mean = 0.0, worst = 0.0;
	for (int r = 0; r < g->n; r++) {
		const pft_stats &a = g->member[r]->st, &b0 = before[r];
		for (int st = 0; st < 4; st++) {
			const long nl = a.stage_launches[st] - b0.stage_launches[st];
			if (nl > 0) cost[r] += (a.stage_ms[st] - b0.stage_ms[st]) / nl;
		}
	}
	if (const char *forced = getenv("PFT_BALANCE_COSTS")) {      // test hook: "c0,c1,..." instead of the measured costs
		for (int r = 0; r < g->n && forced && *forced; r++) {
			cost[r] = atof(forced);
			forced = strchr(forced, ',');
			if (forced) forced++;
		}
	}
	for (int r = 0; r < g->n; r++) {
		if (!(cost[r] > 0.0)) return 0;
		mean += cost[r] / g->n;
		worst = std::max(worst, cost[r]);
	}
	if (worst <= 1.03 * mean) return 0;
	const std::vector<int> planes = balanced_planes(g->planes, cost);
	if (planes == g->planes) return 0;
	// ---- deal the planes out again: the state travels through one padded host array of the whole block
	const pft_case &w = g->whole;
	const size_t nd = (size_t)PFT_NVAR * (w.n1 + 2 * PFT_BT) * (w.n2 + 2 * PFT_BT) * (size_t)(w.n3 + 2 * PFT_BT);
	double *x = (double *)calloc(nd, sizeof(double));
	if (!x) return fail(PFT_ENOMEM, "out of host memory for the rebalancing copy (%zu doubles)", nd);
	int rc2 = pft_group_download(g, x, PFT_FIELD_ALL);
	if (!rc2) {
		group_free_members(g);
		g->planes = planes;
		rc2 = group_build_members(g);
	}
	if (!rc2) rc2 = pft_group_upload(g, x, PFT_FIELD_ALL);
	free(x);
	if (rc2) return rc2;
	return 1;
}

extern "C" int pft_group_balance(pft_group *g, pft_system *sys, double final_time, int attempts)
{
	if (!g || !sys) return fail(PFT_EINVAL, "null argument");
	int solve_rc;
	return group_balance(g, sys, final_time, attempts, &solve_rc);
}

extern "C" int pft_group_solve(pft_group *g, pft_system *sys, double final_time)
{
	if (!g || !sys) { fail(PFT_EINVAL, "null argument"); return PFT_SOLVE_BADSYSTEM; }
	if (!g->balance_tried) {
		// the first 30 attempts of the first solve call double as the calibration run (through the ordinary loop)
		const long t0 = sys->steps_total, budget = sys->max_attempts;
		int solve_rc;
		const int rb = group_balance(g, sys, final_time, 30, &solve_rc);
		if (rb < 0) return rb;
		if (solve_rc != -1000) return solve_rc;              // finished, interrupted, NaN abort or the caller's own budget
		if (budget > 0) sys->max_attempts = budget - (sys->steps_total - t0);
		const int rc = group_run(g, sys, final_time);
		sys->max_attempts = budget;
		return rc;
	}
	return group_run(g, sys, final_time);
}

extern "C" int pft_get_stats(pft_solver *s, pft_stats *out)
{
	if (!s || !out) return fail(PFT_EINVAL, "null argument");
	*out = s->st;
	out->speculative_stage1 = (s->params_ok ? s->spec : (s->spec_ok && !(s->cs.nranks > 1 && !s->p2p))) ? 1 : 0;
	return 0;
}

extern "C" int pft_reset_stats(pft_solver *s)
{
	if (!s) return fail(PFT_EINVAL, "null solver");
	pft_stats keep = s->st;
	memset(&s->st, 0, sizeof s->st);
	s->st.tile_x = keep.tile_x; s->st.tile_y = keep.tile_y; s->st.z_chunk = keep.z_chunk;
	s->st.threads = keep.threads; s->st.grid_blocks = keep.grid_blocks; s->st.bytes_per_field = keep.bytes_per_field;
	s->st.device_pow = keep.device_pow; s->st.runahead = keep.runahead;
	s->st.kernel_generation = keep.kernel_generation; s->st.rows_per_thread = keep.rows_per_thread;
	return 0;
}
