// pft_kernels.cuh -- sm_100a device code of the PorousFreezeThaw hot path (float64).
//
// One kernel template, `merson_stage<STAGE, MODEL, HEAT, SHARE>`, evaluates the right-hand
// side f (reference: apps/intertrack-hybrid-S-freezing/equation.c:566-741 for MODEL 0/1,
// :745-884 for MODEL 2) on the stage input of Merson stage STAGE and fuses the stage
// combination that FOLLOWS it (reference: modules/RK_MPI_SAsolver_hybrid2/
// RK_MPI_SAsolver_hybrid2.c:378-450 combinations, :508-524 error estimate, :657-668 update):
//
//   STAGE 1:  K1 = f(x)                                               -> K1
//   STAGE 2:  K2 = f(K1*h3 + x)              (input formed on the fly) -> aux3 = (K1+K2)*h6 + x
//   STAGE 3:  K3 = f(aux3)                                            -> K3
//   STAGE 4:  K4 = f((K1 + 3.0*K3)*h8 + x)   (on the fly)              -> K4
//   STAGE 5:  K5 = f((0.5*K1 - 1.5*K3 + 2.0*K4)*h + x) (on the fly)    -> eps = max|0.2K1-0.9K3+0.8K4-0.1K5|,
//                                                                        x_new = x + h3*(0.5*(K1+K5) + 2.0*K4)
//
// Arithmetic is strict IEEE-754 double in the reference's association order: this file must be
// compiled with -fmad=false (the build defines PFT_STRICT_FP together with that flag).  Values
// shared between neighbouring cells (face fluxes lambda_face*(u_b-u_a)) are bit-identical to what
// each cell would compute on its own, because the reference evaluates a face from the lower
// cell's '+' side and from the upper cell's '-' side with the same operand order.
//
// Data layout in HBM (per field, per array): (n3+2) z-planes of n2 rows of `px` doubles,
// px = n1 rounded up to 4; plane 0 and plane n3+1 are z-ghost planes that hold the neighbouring
// slab's boundary plane (multi-GPU).  No x/y ghosts are stored: zero-Neumann mirroring is a
// clamped neighbour index (value-identical, equation.c:142-160), the z-front mirror likewise, and
// the z-rear Dirichlet temperature is a kernel argument (equation.c:96-111,175-183).
//
// Thread mapping: a CTA owns a TX x TY tile of (x,y) and marches a chunk of z-planes; each thread
// keeps the z-1/z/z+1 values of its column in registers, the current plane's in-plane neighbours
// come from a double-buffered shared-memory tile with a one-cell halo ring.
#pragma once

#if !defined(PFT_STRICT_FP) && !defined(PFT_FAST_FP)
#error "pft_kernels.cuh must be compiled with -fmad=false -DPFT_STRICT_FP (strict IEEE association); -DPFT_FAST_FP builds the opt-in FMA variant"
#endif

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "pft_ctl.h"

// dynamic shared memory (tests/cuda_emu/ redefines this for its CPU emulation of the kernels)
#ifndef PFT_DYN_SMEM
#define PFT_DYN_SMEM(type, name) extern __shared__ type name[]
#define PFT_DYN_SMEM_ALIGNED(type, name) extern __shared__ __align__(128) type name[]
#endif

namespace pft {

struct Consts {
	double h1_2, h2_2, h3_2, h1_d2, h2_d2, h3_d2;      // equation.c:605-612
	double xi_2_inv_a, bam, sam, eps2_3, eps3_2;       // :442-447; bam=(b*alpha)*mu, sam=(xi_inv_b_sqrt_a2*alpha)*mu
	double p_eps0, p_eps1;
	double alpha, zeta, L, u_star, gamma;
	double g_rho, i_rho, w_rho, g_cp, i_cp, w_cp, g_lam, i_lam, w_lam;
};

struct Geom {
	int n1, n2, n3, px;
	long long plane;          // n2 * px
	int front, rear;          // this slab owns the physical z-front / z-rear boundary
	int zc;                   // planes per z-chunk (gridDim.z chunks)
};

constexpr int PFT_MAX_RANKS = 16;

// one rank's local error maximum of an attempt, as seen by every rank (peer-written)
struct EpsSlot {
	unsigned long long seq;
	double eps;
	int nonfinite, pad;
	unsigned long long pad2;
};

// head of every rank's device slab; peers map it through CUDA IPC
struct SyncBlock {
	unsigned long long halo_flag[2];     // [0] written by the lower neighbour, [1] by the upper: stage sequence numbers
	unsigned long long pad[2];
	EpsSlot slot[2][PFT_MAX_RANKS];      // [attempt & 1][rank]
};

// peer-memory halo exchange (one process per GPU with CUDA IPC, or one process driving several GPUs with
// peer access); all zero on a single GPU and on the NCCL path
struct PeerArgs {
	// the neighbours' ghost planes of this stage's output pair.  Stage 5 writes the candidate pair,
	// which alternates with StepCtl.cur: index [cur ^ 1]; the other stages use index 0.
	double *lo_u[2], *lo_p[2];           // lower neighbour's upper ghost plane
	double *hi_u[2], *hi_p[2];           // upper neighbour's lower ghost plane
	unsigned long long *lo_flag, *hi_flag;   // the neighbours' halo_flag entries this rank signals
	const unsigned long long *my_flags;  // this rank's halo_flag[2]
	unsigned int *tickets;               // finished boundary CTAs, [launch sequence & 7][lower, upper]
	unsigned long long *published;       // [lower, upper]: launch sequence of the last flag this rank sent that way
	SyncBlock *ranks[PFT_MAX_RANKS];     // stage 5: every rank's sync block (error-max broadcast)
	int rank, nranks;
	int wait_lo, wait_hi;                // this slab has a lower / an upper neighbour
	int nzc;                             // number of z-chunks (boundary chunks are scheduled first)
};

// how stage 5 ends
enum { FIN_DECIDE = 0,      // the last CTA takes the step decision on the device (pft_ctl.h) and publishes it
       FIN_PUBLISH_EPS = 1, // ... publishes the global error maximum and leaves the decision to ctl_decide_kernel
                            //     (the host supplies pow(): the library's pow could not be reproduced, pft_pow.h)
       FIN_STORE_LOCAL = 2 };// ... stores this slab's maximum in StepCtl.red (NCCL path: an all-reduce follows)

// Launch arguments of a stage.  Nothing in here depends on the step length or on the outcome of an
// earlier attempt: those scalars are read from the device-resident StepCtl, so the same argument
// block serves every attempt.
struct StageArgs {
	PeerArgs pr;
	StepCtl *ctl;
	double *xu[2], *xp[2];                      // the two state pairs; StepCtl.cur says which one holds x
	const double *gl;                           // the static glass field
	const double *k1u, *k1p, *k3u, *k3p, *k4u, *k4p;
	const double *au, *ap;                      // stage 3: the stored stage input (aux3)
	double *ou, *op;                            // output pair of stages 1-4
	// stage 5 only
	HostSlot *slot;
	const PowTables *tab;                       // device copy of the host library's pow tables (FIN_DECIDE)
	unsigned int nblocks;
	int fin_mode;
	int standalone, standalone_cur;             // a single right-hand-side evaluation outside the time loop (pft_rhs_host)
	double dir_standalone;
	// stage-to-stage data flow (merson_stage_v2): a CTA of stage s+1 starts as soon as the CTAs of stage s it reads from
	// are done, instead of waiting for the whole stage-s grid (see dep_wait)
	int dataflow;
	unsigned long long lseq;                    // number of the attempt these launches belong to, counted by the HOST
	unsigned long long *done_mine;              // [chunk][tile_y][tile_x]: this stage's CTAs write lseq when their output is stored
	const unsigned long long *done_prev;        // the same array of the stage before
	// speculative stage 1 (pft_ctl.h): `ctl` points at two blocks, the attempt reads block lseq & 1; K1 has two
	// buffers that go with the two state pairs
	int spec;
	int s1spec;                                 // this stage-1 launch is S1' (evaluates the candidate of attempt lseq);
	                                            // 2: ... and does not look whether the decision has overtaken it (A/B)
	double *k1u2[2], *k1p2[2];
};

// what a stage kernel reads from the control block when it starts
struct StageScalars {
	double ha, hb, dir_u;
	int cur;
	unsigned long long hseq;             // sequence number of this launch in the halo-flag protocol (same on every rank)
};

#ifndef PFT_CUDA_EMU
template <class T> __device__ __forceinline__ T ctl_ld(const T *p) { return __ldcg(p); }
#else
template <class T> inline T ctl_ld(const T *p) { return *p; }
#endif

// the control block an attempt's launches read
__device__ __forceinline__ StepCtl *ctl_block(const StageArgs &a) { return a.ctl + (a.spec ? (int)(a.lseq & 1ull) : 0); }

template <int STAGE>
__device__ __forceinline__ bool stage_begin(const StageArgs &a, StageScalars &sc)
{
	if (a.standalone) { sc.ha = sc.hb = 0.0; sc.dir_u = a.dir_standalone; sc.cur = a.standalone_cur; sc.hseq = 0; return true; }
	// (L2 loads: the block was written by the previous attempt's controller thread while CTAs of this
	//  launch may already have been resident on the SM -- programmatic dependent launch)
	const StepCtl *c = ctl_block(a);
	if (ctl_ld(&c->status) != ST_RUNNING) return false;   // enqueued beyond the end of the integration: nothing to do
	sc.cur = ctl_ld(&c->cur);
	if (a.spec) {
		// launch order of an attempt: [S1 at solve entry] S2 S3 S4 S5 S1'
		sc.hseq = ctl_ld(&c->hbase) + (STAGE == 1 ? (a.s1spec ? 5 : 0) : STAGE - 1);
		sc.dir_u = ctl_ld(&c->dir[(STAGE == 1 && a.s1spec) ? 4 : STAGE - 1]);     // S1' evaluates at t + h
	} else {
		sc.dir_u = ctl_ld(&c->dir[STAGE - 1]);
		sc.hseq = 5ull * ctl_ld(&c->aseq) + STAGE;
	}
	sc.ha = sc.hb = 0.0;
	if (STAGE == 2) { sc.ha = ctl_ld(&c->h3); sc.hb = ctl_ld(&c->h6); }
	if (STAGE == 4) sc.ha = ctl_ld(&c->h8);
	if (STAGE == 5) { sc.ha = ctl_ld(&c->h); sc.hb = ctl_ld(&c->h3); }
	return true;
}

struct KParams {
	Consts c;
	Geom g;
	StageArgs a;
	unsigned smem_bytes;      // dynamic shared memory of the launch
};

// ----------------------------------------------------------------------------------------
// constitutive laws (equation.c:341-388, 410-421)

__device__ __forceinline__ double mix3(double p, double gl, double xg, double xi, double xw)
{
	return gl * xg + (1.0 - gl) * (p * xi + (1.0 - p) * xw);
}

// lambda(0.5*(pa+pb), 0.5*(ga+gb)) * (-ua + ub): one face of div(lambda grad u), a = lower cell
__device__ __forceinline__ double face_flux(const Consts &c, double pa, double pb, double ga, double gb, double ua, double ub)
{
	double pf = 0.5 * (pa + pb);
	double gf = 0.5 * (ga + gb);
	double lam = mix3(pf, gf, c.g_lam, c.i_lam, c.w_lam);
	return lam * (-ua + ub);
}

// a / b for b > 0.  A zero numerator (exact in the bulk of the domain: uniform temperature, p == 0
// or 1) sends nvcc's IEEE division through its out-of-range slow path (a ~70-instruction
// subroutine, measured at a third of all executed instructions).  (+-0)/b == +-0 for positive b,
// so divide a harmless 1.0 instead and select the zero afterwards: every lane stays on the
// correctly-rounded fast path, non-zero operands are untouched.
// (Numerators below ~2^-960 -- the far tail of the phase field, 1e-290 ... 1e-323 -- still take the
// subroutine.  Two exact remedies were built, verified bit for bit and rejected on measurement: a rescaling
// in front of nvcc's division (round 1, -3 %) and a hand-written division with its own tiny/subnormal path
// (round 2: 58.6 vs 59.0 Gcell-RHS/s early in the 400 case, 57.2 vs 57.5 after 1500 attempts when the tail has
// spread over the domain; profiles/README.md).)
__device__ __forceinline__ double div_pos(double a, double b)
{
	const bool z = (a == 0.0);
	double an = z ? 1.0 : a;
#ifndef PFT_CUDA_EMU
	asm volatile("" : "+d"(an));      // opaque to the optimiser, which would otherwise divide the original a again
#endif
	const double q = an / b;
	return z ? a : q;
}

__device__ __forceinline__ double sshape(const Consts &c, double x)
{
	double y = x - c.p_eps0;
	double r = y * y * (c.eps2_3 - c.eps3_2 * y);
	r = (x >= c.p_eps1) ? 1.0 : r;
	r = (x <= c.p_eps0) ? 0.0 : r;
	return r;
}

// ----------------------------------------------------------------------------------------
// stage input ("state") formation

template <int STAGE, bool HEAT>
struct Raw {      // what has to be remembered of a centre cell until its outputs are formed
	double xu, xp, k1u, k1p, k3u, k3p, k4u, k4p;
};

// Loads the stage input (su, sp) at linear index `idx`; fills `r` with the operands the output
// combination needs (only touched members are kept alive by the compiler).
template <int STAGE, bool HEAT, bool KEEP>
__device__ __forceinline__ void load_state(const StageArgs &a, const double *xu_, const double *xp_, double ha, long long idx,
                                           double &su, double &sp, Raw<STAGE, HEAT> &r)
{
	if (STAGE == 1) {
		su = xu_[idx];
		sp = xp_[idx];
	} else if (STAGE == 3) {
		sp = a.ap[idx];
		su = HEAT ? a.au[idx] : xu_[idx];
	} else if (STAGE == 2) {
		double xp = xp_[idx], k1p = a.k1p[idx];
		sp = k1p * ha + xp;                                     // :388  q = v*h3 + w
		double xu = xu_[idx];
		if (HEAT) { double k1u = a.k1u[idx]; su = k1u * ha + xu; if (KEEP) r.k1u = k1u; }
		else su = xu;
		if (KEEP) { r.xu = xu; r.xp = xp; r.k1p = k1p; }
	} else if (STAGE == 4) {
		double xp = xp_[idx], k1p = a.k1p[idx], k3p = a.k3p[idx];
		sp = (k1p + 3.0 * k3p) * ha + xp;                       // :428
		double xu = xu_[idx];
		if (HEAT) { double k1u = a.k1u[idx], k3u = a.k3u[idx]; su = (k1u + 3.0 * k3u) * ha + xu; }
		else su = xu;
	} else {
		double xp = xp_[idx], k1p = a.k1p[idx], k3p = a.k3p[idx], k4p = a.k4p[idx];
		sp = (0.5 * k1p - 1.5 * k3p + 2.0 * k4p) * ha + xp;     // :449
		double xu = xu_[idx];
		if (HEAT) {
			double k1u = a.k1u[idx], k3u = a.k3u[idx], k4u = a.k4u[idx];
			su = (0.5 * k1u - 1.5 * k3u + 2.0 * k4u) * ha + xu;
			if (KEEP) { r.k1u = k1u; r.k3u = k3u; r.k4u = k4u; }
		} else su = xu;
		if (KEEP) { r.xu = xu; r.xp = xp; r.k1p = k1p; r.k3p = k3p; r.k4p = k4p; }
	}
}

// ----------------------------------------------------------------------------------------

// ----------------------------------------------------------------------------------------
// peer-memory halo exchange protocol
//
//  * boundary-chunk CTAs of stage launch n first wait until the neighbour's flag shows n-1: the
//    neighbour's boundary CTAs finished launch n-1, i.e. its pushes into OUR ghost planes have
//    landed (RAW) and it no longer reads ITS ghost planes of launch n-1 (WAR for our pushes);
//  * they store their output planes k = 0 / n3-1 also into the neighbour's ghost planes;
//  * the last boundary CTA of a side (ticket) publishes flag = n to that neighbour.
// Flags are monotonic, launches are identical on every rank.  Interior CTAs never wait.

#ifndef PFT_CUDA_EMU
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#else
inline unsigned long long ld_acquire_sys(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#endif

#ifndef PFT_CUDA_EMU
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// orders what this thread has observed through ordinary loads before its later TMA requests (asynchronous proxy)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#else
inline unsigned long long ld_acquire_gpu(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_gpu(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void fence_proxy_async() {}
#endif

// programmatic dependent launch (stage kernels are launched with programmatic stream serialization):
// pdl_launch() lets the NEXT launch of the stream start filling SMs as this launch's CTAs retire -- it
// fires once every CTA of this launch has issued it, so the successor never competes with this launch
// for SM slots; pdl_wait() blocks until the PREVIOUS launch has completed and its writes are visible.
// What a CTA does before pdl_wait() (index arithmetic, barrier set-up) overlaps the predecessor's tail.
#ifndef PFT_CUDA_EMU
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
inline void pdl_wait() {}
inline void pdl_launch() {}
#endif

// chunk processed by this CTA.  Launch order = zig-zag from both ends of the slab towards the middle
// (0, nzc-1, 1, nzc-2, ...): the two boundary chunks come first, so that their planes reach the neighbouring
// slabs while the interior is still being computed, and a chunk's z-neighbours are at most two positions away
// in launch order, so that the CTAs the next stage starts with do not depend on the last ones of this stage
// (stage-to-stage data flow, dep_wait).
__device__ __forceinline__ int chunk_of_block(int nzc)
{
	const int z = blockIdx.z;
	return (z & 1) ? nzc - 1 - (z >> 1) : (z >> 1);
}

// ---- stage-to-stage data flow --------------------------------------------------------------------------------
// Within an attempt every array is written by exactly one stage and read only by later ones, and a CTA reads its
// tile plus one ring of cells and one plane below and above its chunk.  So a CTA of stage s+1 may start once the
// 27 CTAs of stage s around its own position are done; it need not wait for the stragglers of the whole grid (on a
// 2 M-cell slab the SMs idle 15-22 % of every stage kernel in the tail of the last wave).  Every CTA of stages 1-4
// publishes `lseq` (the host's attempt number) in done_mine[] when its output is stored -- also when it has nothing
// to do -- and the CTAs of stages 2-5 wait on done_prev[] instead of griddepcontrol.wait.  In the classic launch
// order stage 1 keeps the full wait: it needs the step decision, which needs every CTA of stage 5.  With speculative
// stage 1 (pft_ctl.h) stage 5 publishes too, S1' follows it through the same 27-neighbour rule, and the one grid-wide
// wait of an attempt is stage 2's wait for the decision (ctl_dep_wait).  Launch order guarantees progress: a dependent
// grid becomes resident only when every CTA of its predecessor has started (pdl_launch at CTA entry).
__device__ __forceinline__ long long cta_slot(int chunk) { return ((long long)chunk * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x; }

__device__ __forceinline__ void dep_poll(const StageArgs &a, int chunk, int nzc, int n)
{
	const int X = (int)blockIdx.x + n % 3 - 1, Y = (int)blockIdx.y + (n / 3) % 3 - 1, Z = chunk + n / 9 - 1;
	if (X >= 0 && X < (int)gridDim.x && Y >= 0 && Y < (int)gridDim.y && Z >= 0 && Z < nzc) {
		const unsigned long long *f = &a.done_prev[((long long)Z * gridDim.y + Y) * gridDim.x + X];
		while (ld_acquire_gpu(f) < a.lseq) { }
	}
}

__device__ __forceinline__ void dep_wait(const StageArgs &a, int chunk, int nzc, int tid)
{
	if (tid < 27) dep_poll(a, chunk, nzc, tid);
	__syncthreads();
	if (tid == 0) fence_proxy_async();
}

__device__ __forceinline__ void dep_signal(const StageArgs &a, int chunk, int tid)
{
	if (!a.dataflow) return;
	__syncthreads();                         // every thread's stores precede thread 0's fence (CTA barrier)
	// (S1' produces K1 for the attempt that follows: stage 2 of attempt lseq+1 waits for it)
	if (tid == 0) { __threadfence(); st_release_gpu(&a.done_mine[cta_slot(chunk)], a.lseq + (a.s1spec ? 1ull : 0ull)); }
}

// speculative stage 1: the block of attempt lseq becomes valid when the decision of attempt lseq-1 is published
// (or the host wrote it at solve entry).  Returns false when the launch lies beyond the end of the integration.
// Only stage 2 ever waits here for real: it is the first launch of an attempt that needs the decision.
// The wait on the block (thread 0) and the 27 polls of dep_wait (second warp) run side by side and end in ONE CTA
// barrier: a CTA starts its first tensor load two L2 round trips after it became resident (the block's `ready`
// word next to the completion words, then the block's scalars), which is what a stage costs on a grid of a few
// hundred CTAs.
__device__ __forceinline__ bool ctl_dep_wait(const StageArgs &a, int chunk, int nzc, int tid)
{
	__shared__ int s_go;
	const int base = (blockDim.x * blockDim.y >= 64) ? 32 : 0;      // (tiles of fewer than 64 threads: one after the other)
	if (tid == 0) {
		const StepCtl *b = ctl_block(a);
		int go;
		for (;;) {
			if (ld_acquire_gpu(&b->ready) == a.lseq) { go = 1; break; }
			if (ld_acquire_gpu(&b->closed_from) <= a.lseq) { go = 0; break; }
		}
		s_go = go;
	}
	if (tid >= base && tid < base + 27) dep_poll(a, chunk, nzc, tid - base);
	__syncthreads();
	if (tid == 0) fence_proxy_async();
	return s_go != 0;
}

__device__ __forceinline__ void halo_wait(const PeerArgs &pr, unsigned long long hseq, bool lo, bool hi, int tid)
{
	const bool wl = lo && pr.wait_lo, wh = hi && pr.wait_hi;
	if (!(wl || wh) || hseq == 0) return;
	if (tid == 0) {
		if (wl) while (ld_acquire_sys(&pr.my_flags[0]) + 1 < hseq) { }
		if (wh) while (ld_acquire_sys(&pr.my_flags[1]) + 1 < hseq) { }
	}
	__syncthreads();
}

// one thread: the flag of launch hseq goes to a neighbour once the flag of launch hseq-1 has gone
__device__ __forceinline__ void halo_publish(const PeerArgs &pr, int side, unsigned long long *flag, unsigned long long hseq)
{
	while (ld_acquire_gpu(&pr.published[side]) + 1 < hseq) { }
	__threadfence_system();
	st_release_sys(flag, hseq);
	st_release_gpu(&pr.published[side], hseq);
}

__device__ __forceinline__ void halo_signal(const PeerArgs &pr, unsigned long long hseq, bool lo, bool hi, int tid)
{
	const bool wl = lo && pr.wait_lo, wh = hi && pr.wait_hi;
	if (!(wl || wh) || hseq == 0) return;
	__syncthreads();                     // every thread's peer stores happen before thread 0's ticket (CTA barrier) ...
	if (tid == 0) {
		// ... a device-scope release per CTA, and ONE system-scope fence per side and launch, by the CTA that
		// takes the last ticket: causality is transitive (barrier -> device-scope release/acquire on the ticket
		// -> system-scope release of the flag).  A system fence in every boundary CTA queued up for 5-10 us
		// when the boundary CTAs of a launch finish together (measured, profiles/round1_tuning.md).
		__threadfence();
		const unsigned int n = gridDim.x * gridDim.y;
		// With the stage-to-stage data flow a boundary CTA of launch n+1 may finish while one of launch n is still
		// running (certainly so when an S1' CTA returns at once, candidate_rejected).  So (a) every launch counts on a
		// ticket pair of its own, eight in rotation -- launches eight apart lie on different sides of a step decision,
		// which needs every CTA of the attempt before it -- and (b) flags go out in launch order: a neighbour reads
		// "flag >= n" as "the boundary planes of every launch up to n have landed".
		unsigned int *tk = pr.tickets + 2 * (int)(hseq & 7ull);
		if (wl && atomicAdd(&tk[0], 1u) == n - 1) { tk[0] = 0u; halo_publish(pr, 0, pr.lo_flag, hseq); }
		if (wh && atomicAdd(&tk[1], 1u) == n - 1) { tk[1] = 0u; halo_publish(pr, 1, pr.hi_flag, hseq); }
	}
}

// a skipped launch (stage 1 after a rejected attempt: K1 is still valid) still has to advance the flags
__device__ __forceinline__ void halo_skip(const PeerArgs &pr, unsigned long long hseq)
{
	if (hseq == 0) return;
	if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0) {
		if (pr.wait_lo) halo_publish(pr, 0, pr.lo_flag, hseq);
		if (pr.wait_hi) halo_publish(pr, 1, pr.hi_flag, hseq);
	}
}

// where a stage writes: stages 1-4 their fixed output pair, stage 5 the candidate state pair
struct OutPtrs {
	double *ou, *op;
	double *lo_u, *lo_p, *hi_u, *hi_p;   // the neighbours' ghost planes of the same pair (peer path), else null
};

template <int STAGE>
__device__ __forceinline__ OutPtrs out_ptrs(const StageArgs &a, int cur)
{
	OutPtrs o;
	// stage 5 writes the candidate pair; with speculative stage 1 K1 has a buffer per state pair: S1' fills the one
	// of the candidate pair, the stage 1 at solve entry the one of the live pair
	const int oi = (STAGE == 5) ? (cur ^ 1) : (STAGE == 1 && a.spec) ? (a.s1spec ? cur ^ 1 : cur) : 0;
	o.ou = (STAGE == 5) ? a.xu[oi] : (STAGE == 1 && a.spec) ? a.k1u2[oi] : a.ou;
	o.op = (STAGE == 5) ? a.xp[oi] : (STAGE == 1 && a.spec) ? a.k1p2[oi] : a.op;
	o.lo_u = a.pr.lo_u[oi]; o.lo_p = a.pr.lo_p[oi];
	o.hi_u = a.pr.hi_u[oi]; o.hi_p = a.pr.hi_p[oi];
	return o;
}

// store one output pair value, and push it into the neighbours' ghost planes from the boundary planes
template <bool HEAT>
__device__ __forceinline__ void store_out(const OutPtrs &w, long long o, long long col, int k, int n3, double vu, double vp)
{
	w.op[o] = vp;
	if (HEAT) w.ou[o] = vu;
	if (k == 0 && w.lo_p) { w.lo_p[col] = vp; if (HEAT) w.lo_u[col] = vu; }
	if (k == n3 - 1 && w.hi_p) { w.hi_p[col] = vp; if (HEAT) w.hi_u[col] = vu; }
}

// ---- stage 5 epilogue: grid-wide max of the error estimate, then the step decision.
// Several GPUs: the last CTA of every rank's launch hands its slab maximum to every rank (peer stores,
// one thread per rank), waits for the others' and forms the global maximum -- the role of
// MPI_Allreduce(MAX), RK_MPI_SAsolver_hybrid2.c:572 -- so that every rank takes the same decision
// without a further launch.  Peers never wait for THIS CTA (their boundary waits concern CTAs that
// are long finished), so the spin cannot deadlock.
// The decision itself (pft_ctl.h: accept test, new step length with the host library's pow, update of
// t / h / which pair holds x, end-of-interval logic, derived scalars of the next attempt) is taken by
// ONE thread here; the next launches read its result from the control block.
// scratch: the CTA's dynamic shared memory (the plane buffers are dead by now).  The decision is a chain of
// dependent accesses to the control block and to the pow tables; staged in shared memory by the whole CTA
// (two coalesced copies) it costs ~2 us instead of ~10 us of serial global-memory latency -- which is 5 % of an
// attempt on the thin slabs of an 8-GPU run.
// speculative stage 1: hand the freshly written block to attempt lseq+1; at the end of the integration close both
// blocks for every later attempt (launches enqueued ahead return at once)
__device__ __forceinline__ void ctl_release_next(StepCtl *cur_blk, StepCtl *nxt, unsigned long long lseq, int status)
{
	__threadfence();
	if (status != ST_RUNNING) {
		*(volatile unsigned long long *)&nxt->closed_from = lseq + 1ull;
		st_release_gpu(&cur_blk->closed_from, lseq + 1ull);
	}
	st_release_gpu(&nxt->ready, lseq + 1ull);
}

__device__ __forceinline__ void stage5_finish(const StageArgs &a, int tid, double emax, int nonfinite, void *scratch, unsigned scratch_bytes)
{
	// CTA max through a shared-memory atomic on the bit pattern (non-negative doubles order
	// like their uint64 patterns; the tile need not be a multiple of the warp size)
	__shared__ unsigned long long s_bits;
	__shared__ int s_nonfinite, s_last;
	if (tid == 0) { s_bits = 0ull; s_nonfinite = 0; }
	__syncthreads();
	if (emax > 0.0) atomicMax(&s_bits, (unsigned long long)__double_as_longlong(emax));
	if (nonfinite) atomicOr(&s_nonfinite, 1);
	__syncthreads();
	StepCtl *ctl = ctl_block(a);             // the block this attempt read
	StepCtl *nxt = a.spec ? a.ctl + (int)((a.lseq + 1ull) & 1ull) : ctl;      // the block the next attempt reads
	if (tid == 0) {
		if (s_bits) atomicMax(&ctl->eps_bits, s_bits);
		if (s_nonfinite) atomicOr(&ctl->nonfinite, 1);
		__threadfence();
		const unsigned int t = atomicAdd(&ctl->ticket, 1u);
		s_last = (t == a.nblocks - 1);
		if (s_last) {                        // last CTA of the launch: every other CTA's max is in
			__threadfence();
			s_bits = atomicMax(&ctl->eps_bits, 0ull);
			s_nonfinite = atomicOr(&ctl->nonfinite, 0);
			ctl->eps_bits = 0ull;
			ctl->nonfinite = 0;
			ctl->ticket = 0u;
		}
	}
	__syncthreads();
	if (!s_last) return;                     // CTA-uniform

	const unsigned long long seq = ctl->aseq + 1ull;      // number of this attempt
	if (a.pr.nranks > 1) {
		const int b = (int)(seq & 1ull);
		const int nthr = blockDim.x * blockDim.y;
		const double mine = __longlong_as_double((long long)s_bits);
		const int mine_nf = s_nonfinite;
		__syncthreads();                     // everyone has read the slab result before it is overwritten below
		for (int r = tid; r < a.pr.nranks; r += nthr) {
			EpsSlot *e = &a.pr.ranks[r]->slot[b][a.pr.rank];
			e->eps = mine;
			e->nonfinite = mine_nf;
			__threadfence_system();
			st_release_sys(&e->seq, seq);
		}
		if (tid == 0) { s_bits = 0ull; s_nonfinite = 0; }
		__syncthreads();
		for (int r = tid; r < a.pr.nranks; r += nthr) {
			const EpsSlot *e = &a.pr.ranks[a.pr.rank]->slot[b][r];
			while (ld_acquire_sys(&e->seq) != seq) { }
			const double v = *(volatile const double *)&e->eps;
			if (v > 0.0) atomicMax(&s_bits, (unsigned long long)__double_as_longlong(v));
			if (*(volatile const int *)&e->nonfinite) atomicOr(&s_nonfinite, 1);
		}
		__syncthreads();
	}
	if (a.fin_mode == FIN_DECIDE && scratch_bytes >= (unsigned)(sizeof(StepCtl) + sizeof(PowTables) + 32)) {
		const int nthr = blockDim.x * blockDim.y;
		unsigned long long *sc_ctl = (unsigned long long *)scratch;
		unsigned long long *sc_tab = sc_ctl + (sizeof(StepCtl) + 15) / 16 * 2;
		static_assert(sizeof(StepCtl) % 8 == 0 && sizeof(PowTables) % 8 == 0, "copied as 8-byte words");
		__syncthreads();                                  // everyone is done with the plane buffers
		for (unsigned i = tid; i < sizeof(StepCtl) / 8; i += nthr) sc_ctl[i] = ctl_ld((const unsigned long long *)ctl + i);
		for (unsigned i = tid; i < sizeof(PowTables) / 8; i += nthr) sc_tab[i] = ((const unsigned long long *)a.tab)[i];
		__syncthreads();
		if (tid == 0) {
			StepCtl &c = *(StepCtl *)sc_ctl;
			const double eps = __longlong_as_double((long long)s_bits);
			const double h_attempt = c.h;
			const int cmd = ctl_decide(c, eps, s_nonfinite, (const PowTables *)sc_tab, 0.0);
			ctl_publish(c, a.slot, eps, s_nonfinite, cmd, h_attempt);
		}
		__syncthreads();
		for (unsigned i = tid; i < PFT_CTL_COPY_WORDS; i += nthr) ((unsigned long long *)nxt)[i] = sc_ctl[i];
		__syncthreads();
		if (tid == 0) {
			if (a.spec) ctl_release_next(ctl, nxt, a.lseq, ((const StepCtl *)sc_ctl)->status);
			__threadfence_system();
			*(volatile unsigned long long *)&a.slot->seq = seq;
		}
		return;
	}
	if (tid == 0) {
		const double eps = __longlong_as_double((long long)s_bits);
		const int nfl = s_nonfinite;
		HostSlot *s = a.slot;
		if (a.fin_mode == FIN_STORE_LOCAL) {
			ctl->red[0] = eps; ctl->red[1] = nfl ? 1.0 : 0.0;
		} else if (a.fin_mode == FIN_PUBLISH_EPS) {
			ctl->red[0] = eps; ctl->red[1] = nfl ? 1.0 : 0.0;
			s->eps = eps; s->nonfinite = nfl;
			s->delta_local = ctl->delta_local; s->h3_attempt = fabs(ctl->h3);
			__threadfence_system();
			*(volatile unsigned long long *)&s->seq_eps = seq;
		} else {
			const double h_attempt = ctl->h;
			if (a.spec) for (unsigned i = 0; i < PFT_CTL_COPY_WORDS; i++) ((unsigned long long *)nxt)[i] = ((const unsigned long long *)ctl)[i];
			const int cmd = ctl_decide(*nxt, eps, nfl, a.tab, 0.0);
			__threadfence();
			ctl_publish(*nxt, s, eps, nfl, cmd, h_attempt);
			if (a.spec) ctl_release_next(ctl, nxt, a.lseq, nxt->status);
			__threadfence_system();
			*(volatile unsigned long long *)&s->seq = seq;
		}
	}
}

// S1' only: has the decision of this attempt been published already, and did it keep the old state (rejected
// attempt, or the NaN branch)?  Then nobody will read the K1 this launch is about to evaluate.
__device__ __forceinline__ bool candidate_rejected(const StageArgs &a, int cur, int tid)
{
	__shared__ int s_rej;
	if (tid == 0) {
		const StepCtl *nx = a.ctl + (int)((a.lseq + 1ull) & 1ull);
		s_rej = (ld_acquire_gpu(&nx->ready) == a.lseq + 1ull && ctl_ld(&nx->cur) == cur) ? 1 : 0;
	}
	__syncthreads();
	return s_rej != 0;
}

// the decision as a kernel of its own (one thread): after an NCCL all-reduce of StepCtl.red, and/or with
// pow(delta/max_eps, 0.2) supplied by the host (use_host_pow)
__global__ void ctl_decide_kernel(StepCtl *ctl, HostSlot *s, const PowTables *tab, double pw_host, int use_host_pow)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	if (ctl->status != ST_RUNNING) return;
	const double eps = ctl->red[0];
	const int nfl = ctl->red[1] != 0.0;
	const unsigned long long seq = ctl->aseq + 1ull;
	const double h_attempt = ctl->h;
	const int cmd = ctl_decide(*ctl, eps, nfl, use_host_pow ? nullptr : tab, pw_host);
	__threadfence();
	ctl_publish(*ctl, s, eps, nfl, cmd, h_attempt);
	__threadfence_system();
	*(volatile unsigned long long *)&s->seq = seq;
}

// NCCL path without a usable device pow: hand the all-reduced error maximum to the host
__global__ void ctl_publish_eps_kernel(StepCtl *ctl, HostSlot *s)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	if (ctl->status != ST_RUNNING) return;
	s->eps = ctl->red[0]; s->nonfinite = ctl->red[1] != 0.0;
	s->delta_local = ctl->delta_local; s->h3_attempt = fabs(ctl->h3);
	__threadfence_system();
	*(volatile unsigned long long *)&s->seq_eps = ctl->aseq + 1ull;
}

// MODEL: 0 GradP (calc_mode 0/10), 1 SigmaP1-P (1/11), 2 temperature model (2)
// HEAT : the heat equation is active (false for calc_mode 10/11: du/dt = 0, u is static)
// Register-staged predecessor of merson_stage_tma (per-thread global loads, halo-duty threads); it serves
// calc_mode 10/11 and, behind PFT_FLAG_LEGACY_LOADS, as an A/B reference for the TMA kernels.
template <int STAGE, int MODEL, bool HEAT>
__global__ void __launch_bounds__(512, 1) merson_stage(const __grid_constant__ KParams P)
{
	const Consts &c = P.c;
	const Geom &g = P.g;
	const StageArgs &a = P.a;

	pdl_launch();
	pdl_wait();
	StageScalars sc;
	if (!stage_begin<STAGE>(a, sc)) return;
	if (STAGE == 1 && !a.standalone && ctl_ld(&a.ctl->skip1)) { halo_skip(a.pr, sc.hseq); return; }
	const double *xu_ = a.xu[sc.cur], *xp_ = a.xp[sc.cur];
	const OutPtrs w = out_ptrs<STAGE>(a, sc.cur);
	const double ha = sc.ha, hb = sc.hb;

	constexpr int NF = HEAT ? 3 : 1;        // fields kept in the shared tile: p [, u, gl]

	const int TX = blockDim.x, TY = blockDim.y;
	const int tx = threadIdx.x, ty = threadIdx.y;
	const int tid = ty * TX + tx;
	const int SP = TX + 2;                  // tile pitch incl. halo ring
	const int TS = SP * (TY + 2);           // one field of one buffer

	PFT_DYN_SMEM(double, smem);
	double *tile = smem;                                    // [2][NF][TS]

	const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
	const int i = i0 + tx, j = j0 + ty;
	const bool active = (i < g.n1) && (j < g.n2);
	const int ic = min(i, g.n1 - 1), jc = min(j, g.n2 - 1);
	const long long col = (long long)jc * g.px + ic;        // in-plane offset of this thread's column

	const int k0 = chunk_of_block(a.pr.nzc) * g.zc;
	const int k1 = min(g.n3, k0 + g.zc);
	const bool chunk_lo = (k0 == 0), chunk_hi = (k1 == g.n3);
	halo_wait(a.pr, sc.hseq, chunk_lo, chunk_hi, tid);

	// halo-ring duty: the first 2*TX+2*TY threads each fetch one ring cell per plane
	const int nh = 2 * TX + 2 * TY;
	const bool hduty = tid < nh;
	int hs = 0;             // slot in the tile
	long long hcol = 0;     // in-plane offset of the ring cell (clamped = mirrored)
	if (hduty) {
		int hx, hy;
		if (tid < TX) { hx = tid; hy = -1; }
		else if (tid < 2 * TX) { hx = tid - TX; hy = TY; }
		else if (tid < 2 * TX + TY) { hx = -1; hy = tid - 2 * TX; }
		else { hx = TX; hy = tid - 2 * TX - TY; }
		int gi = min(max(i0 + hx, 0), g.n1 - 1), gj = min(max(j0 + hy, 0), g.n2 - 1);
		hcol = (long long)gj * g.px + gi;
		hs = (hy + 1) * SP + (hx + 1);
	}
	const int cs = (ty + 1) * SP + (tx + 1);  // this thread's own slot

	Raw<STAGE, HEAT> rc = {}, rn = {}, rdummy = {};
	double um, pm, gm;       // plane k-1
	double uc, pc, gc;       // plane k
	double un, pn, gn;       // plane k+1
	double fzm = 0.0;        // flux through the face between plane k-1 and k
	double emax = 0.0;
	int nonfinite = 0;

	// ---- prologue: planes k0-1 (registers) and k0 (registers + tile 0)
	{
		const long long base = (long long)(k0 + 1) * g.plane;
		load_state<STAGE, HEAT, true>(a, xu_, xp_, ha, base + col, uc, pc, rc);
		gc = a.gl[base + col];
		if (k0 == 0 && g.front) { um = uc; pm = pc; gm = gc; }          // z-front mirror (equation.c:164-174)
		else {
			load_state<STAGE, HEAT, false>(a, xu_, xp_, ha, base - g.plane + col, um, pm, rdummy);
			gm = a.gl[base - g.plane + col];
		}
		double *t0 = tile;
		t0[cs] = pc;
		if (HEAT) { t0[TS + cs] = uc; t0[2 * TS + cs] = gc; }
		if (hduty) {
			double hu, hp;
			load_state<STAGE, HEAT, false>(a, xu_, xp_, ha, base + hcol, hu, hp, rdummy);
			t0[hs] = hp;
			if (HEAT) { t0[TS + hs] = hu; t0[2 * TS + hs] = a.gl[base + hcol]; }
		}
		if (HEAT) fzm = face_flux(c, pm, pc, gm, gc, um, uc);
	}

	int cur = 0;
	for (int k = k0; k < k1; ++k) {
		const long long base_n = (long long)(k + 2) * g.plane;      // plane k+1
		const bool top = (k + 1 == g.n3) && g.rear;                   // plane k+1 is the physical z-rear ghost
		const bool more = (k + 1 < k1);

		// ---- fetch plane k+1 (kept in registers until the end of the iteration)
		double hu = 0.0, hp = 0.0, hg = 0.0;
		if (top) {
			un = sc.dir_u; pn = pc; gn = gc;                          // equation.c:175-183 / :251-261
		} else {
			load_state<STAGE, HEAT, true>(a, xu_, xp_, ha, base_n + col, un, pn, rn);
			gn = a.gl[base_n + col];
		}
		if (more && hduty) {
			load_state<STAGE, HEAT, false>(a, xu_, xp_, ha, base_n + hcol, hu, hp, rdummy);
			if (HEAT) hg = a.gl[base_n + hcol];
		}

		__syncthreads();                                             // tile[cur] complete
		const double *tp = tile + cur * NF * TS;
		const double *tu = tp + TS, *tg = tp + 2 * TS;

		const double pw = tp[cs - 1], pe = tp[cs + 1], ps = tp[cs - SP], pnn = tp[cs + SP];

		double sum = 0.0;        // h1_2*(...) + h2_2*(...) + h3_2*(...) of div(lambda grad u)
		if (HEAT) {
			const double ue = tu[cs + 1], ge = tg[cs + 1];
			const double unn = tu[cs + SP], gnn = tg[cs + SP];
			const double uw = tu[cs - 1], gw = tg[cs - 1];
			const double us = tu[cs - SP], gs = tg[cs - SP];
			double fe = face_flux(c, pc, pe, gc, ge, uc, ue);
			double fn = face_flux(c, pc, pnn, gc, gnn, uc, unn);
			double fw = face_flux(c, pw, pc, gw, gc, uw, uc);
			double fs = face_flux(c, ps, pc, gs, gc, us, uc);
			// the z+ face is the first consumer of the plane fetched at the top of this iteration:
			// keep it last so that the in-plane work covers the load latency
			double fzp = face_flux(c, pc, pn, gc, gn, uc, un);
			sum = c.h1_2 * (-fw + fe) + c.h2_2 * (-fs + fn) + c.h3_2 * (-fzm + fzp);   // equation.c:711-720
			fzm = fzp;
		}

		double du, dp;
		if (MODEL == 2) {
			// equation.c:835-867
			double rho = mix3(pc, gc, c.g_rho, c.i_rho, c.w_rho);
			double cp = mix3(pc, gc, c.g_cp, c.i_cp, c.w_cp);
			double ch = cosh(c.gamma * (uc - c.u_star));
			double dp_du = (-0.5 * c.gamma / (ch * ch)) * fmax(0.0, 1.0 - c.zeta * gc);
			du = div_pos(sum, rho * (cp - c.L * dp_du));
			dp = dp_du * du;
		} else {
			// equation.c:659-694
			dp = c.h1_2 * (-(-pw + pc) + (-pc + pe))
			   + c.h2_2 * (-(-ps + pc) + (-pc + pnn))
			   + c.h3_2 * (-(-pm + pc) + (-pc + pn));
			double one_m_p = 1.0 - pc;
			double well = c.xi_2_inv_a * pc * one_m_p * (pc - 0.5);
			double du_star = uc - c.u_star;
			if (MODEL == 0) {
				double gx = c.h1_d2 * (-pw + pe), gy = c.h2_d2 * (-ps + pnn), gz = c.h3_d2 * (-pm + pn);
				double nrm = sqrt(gx * gx + gy * gy + gz * gz) + 1E-10;               // :332-336
				dp += well - c.bam * nrm * du_star;                                       // :370
			} else {
				dp += well - c.sam * sshape(c, pc) * sshape(c, one_m_p) * fmax(pc * one_m_p, 0.0) * du_star;   // :387
			}
			dp = div_pos(dp, c.alpha);
			dp *= fmax(0.0, 1.0 - c.zeta * gc);
			if (HEAT) {
				double rho = mix3(pc, gc, c.g_rho, c.i_rho, c.w_rho);
				double cp = mix3(pc, gc, c.g_cp, c.i_cp, c.w_cp);
				du = div_pos(div_pos(sum, rho) + c.L * dp, cp);                           // :711-723
			} else du = 0.0;
		}

		// ---- fused combination that follows this stage
		if (active) {
			const long long o = (long long)(k + 1) * g.plane + col;
			if (STAGE == 1 || STAGE == 3 || STAGE == 4) {
				store_out<HEAT>(w, o, col, k, g.n3, du, dp);
			} else if (STAGE == 2) {
				const double vp = (rc.k1p + dp) * hb + rc.xp;                                 // :408
				const double vu = HEAT ? (rc.k1u + du) * hb + rc.xu : 0.0;
				store_out<HEAT>(w, o, col, k, g.n3, vu, vp);
			} else {
				double e = fabs(0.2 * rc.k1p - 0.9 * rc.k3p + 0.8 * rc.k4p - 0.1 * dp);   // :521
				if (!(e <= 1.7976931348623157e308)) nonfinite = 1;
				emax = (e > emax) ? e : emax;
				const double vp = rc.xp + hb * (0.5 * (rc.k1p + dp) + 2.0 * rc.k4p);          // :667
				double vu = 0.0;
				if (HEAT) {
					double eu = fabs(0.2 * rc.k1u - 0.9 * rc.k3u + 0.8 * rc.k4u - 0.1 * du);
					if (!(eu <= 1.7976931348623157e308)) nonfinite = 1;
					emax = (eu > emax) ? eu : emax;
					vu = rc.xu + hb * (0.5 * (rc.k1u + du) + 2.0 * rc.k4u);
				}
				store_out<HEAT>(w, o, col, k, g.n3, vu, vp);
			}
		}

		// ---- publish plane k+1 into the other tile buffer, rotate the register window
		if (more) {
			double *tn = tile + (cur ^ 1) * NF * TS;
			tn[cs] = pn;
			if (HEAT) { tn[TS + cs] = un; tn[2 * TS + cs] = gn; }
			if (hduty) {
				tn[hs] = hp;
				if (HEAT) { tn[TS + hs] = hu; tn[2 * TS + hs] = hg; }
			}
		}
		um = uc; pm = pc; gm = gc;
		uc = un; pc = pn; gc = gn;
		rc = rn;
		cur ^= 1;
	}

	halo_signal(a.pr, sc.hseq, chunk_lo, chunk_hi, tid);
	if (STAGE == 5) stage5_finish(a, tid, emax, nonfinite, smem, P.smem_bytes);
}

// ----------------------------------------------------------------------------------------
// layout conversion between the reference's padded host layout (staged in device memory) and the
// compact device layout

// dir 0: padded -> compact, dir 1: compact -> padded; planes klo .. khi-1 of the slab (interior: 0 .. n3-1;
// klo = -1 / khi = n3+1 include the z-ghost plane of BOTH layouts: a neighbouring slab's boundary plane)
__global__ void repack_kernel(double *padded, double *compact, int n1, int n2, int klo, int khi, int px, long long plane,
                              int N1, long long rowsize, int bt, int dir)
{
	const long long total = (long long)n1 * n2 * (khi - klo);
	for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		int i = (int)(t % n1);
		long long r = t / n1;
		int j = (int)(r % n2);
		int k = (int)(r / n2) + klo;
		long long ip = (long long)(k + bt) * rowsize + (long long)(j + bt) * N1 + bt + i;
		long long ic = (long long)(k + 1) * plane + (long long)j * px + i;
		if (dir == 0) compact[ic] = padded[ip];
		else padded[ip] = compact[ic];
	}
}

// dense [n3][n2][n1] <-> compact device layout
__global__ void dense_kernel(double *dense, double *compact, int n1, int n2, int n3, int px, long long plane, int dir)
{
	const long long total = (long long)n1 * n2 * n3;
	for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		int i = (int)(t % n1);
		long long r = t / n1;
		int j = (int)(r % n2);
		int k = (int)(r / n2);
		long long ic = (long long)(k + 1) * plane + (long long)j * px + i;
		if (dir == 0) compact[ic] = dense[t];
		else dense[t] = compact[ic];
	}
}

// uniform bins over the slab's bounding box, edge >= the culling radius (host side: pft_rasterise)
struct RasterBins {
	double org[3], inv[3];
	int nb[3];
};

// sphere rasteriser (equation.c:508-529): gl = max(gl, 0.5*(1 - tanh(0.5/xi_gl*(|x-c| + 1e-10 - R)))).
// tanh(a) rounds to exactly 1.0 for a > 19.07, i.e. the profile of a sphere further away than
// cut = R + 19.5/s is exactly 0 and cannot raise gl >= 0.  The centres are sorted into bins of edge >= cut
// (start[b] .. start[b+1]), so a node visits the 27 bins around it: every sphere within `cut` of the node is
// among them, the others would not have changed the result -- bit-identical to testing every sphere, and
// max() does not care about the order.
__global__ void rasterise_kernel(double *gl, const double *centres, const int *start, const RasterBins rb,
                                 int n1, int n2, int n3, int px, long long plane, int first_row, int total_n3,
                                 double L1, double L2, double L3, double xi_gl, double radius)
{
	const long long total = (long long)n1 * n2 * n3;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= total) return;
	const int i = (int)(t % n1);
	const long long r = t / n1;
	const int j = (int)(r % n2), k = (int)(r / n2);
	const double x = L1 * (0.5 + i) / n1;
	const double y = L2 * (0.5 + j) / n2;
	const double z = L3 * (0.5 + k + first_row) / total_n3;
	const long long ic = (long long)(k + 1) * plane + (long long)j * px + i;
	double v = gl[ic];
	const double s = 0.5 / xi_gl;
	const double cut = radius + 19.5 / s;
	const double cut2 = cut * cut;
	const double pos[3] = {x, y, z};
	int b0[3], b1[3];
	for (int d = 0; d < 3; d++) {
		const double f = (pos[d] - rb.org[d]) * rb.inv[d];
		const int c = (f < 0) ? 0 : (f >= rb.nb[d] ? rb.nb[d] - 1 : (int)f);
		b0[d] = max(c - 1, 0); b1[d] = min(c + 1, rb.nb[d] - 1);
	}
	for (int bz = b0[2]; bz <= b1[2]; bz++)
		for (int by = b0[1]; by <= b1[1]; by++) {
			const int row = (bz * rb.nb[1] + by) * rb.nb[0];
			const int q0 = start[row + b0[0]], q1 = start[row + b1[0] + 1];      // the x-neighbour bins are contiguous
			for (int q = q0; q < q1; q++) {
				const double dx = x - centres[3 * q], dy = y - centres[3 * q + 1], dz = z - centres[3 * q + 2];
				const double d2 = dx * dx + dy * dy + dz * dz;
				if (d2 > cut2) continue;
				const double d = sqrt(d2) + 1E-10;
				const double f = 0.5 * (1.0 - tanh(s * (d - radius)));
				v = (v < f) ? f : v;
			}
		}
	gl[ic] = v;
}

// test hook: the device instance of the controller's pow twin on an array (pft_debug_pow)
__global__ void pow_twin_kernel(const PowTables *tab, const double *x, double y, double *out, int n)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = pft_pow_twin(*tab, x[i], y);
}

// count of interior nodes with p > 0.5 (scripts/avg.sh:9-20)
__global__ void ice_count_kernel(const double *p, int n1, int n2, int n3, int px, long long plane, unsigned long long *count)
{
	const long long total = (long long)n1 * n2 * n3;
	unsigned long long local = 0;
	for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		int i = (int)(t % n1);
		long long r = t / n1;
		int j = (int)(r % n2);
		int k = (int)(r / n2);
		if (p[(long long)(k + 1) * plane + (long long)j * px + i] > 0.5) local++;
	}
	if (local) atomicAdd(count, local);
}

}  // namespace pft
