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

// device-resident control block of the error reduction / step acceptance
struct Ctl {
	unsigned long long eps_bits;   // running max of the error estimate (bit pattern of a non-negative double)
	unsigned int ticket;           // finished-CTA counter of the current stage-5 launch
	int accept;                    // decision of the last attempt (read by the speculative stage 1)
	int nonfinite;                 // a non-finite error component was seen (NaN handling)
	int pad;
};

// host-mapped result slot (pinned, zero-copy): written by the last CTA of stage 5
struct HostSlot {
	unsigned long long seq;
	double eps;
	int accept;
	int nonfinite;
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

// peer-memory halo exchange (one process per GPU, slabs mapped into each other with CUDA IPC);
// all zero on a single GPU and on the NCCL path
struct PeerArgs {
	double *lo_u, *lo_p;                 // lower neighbour's upper ghost plane of this stage's output pair
	double *hi_u, *hi_p;                 // upper neighbour's lower ghost plane
	unsigned long long *lo_flag, *hi_flag;   // the neighbours' halo_flag entries this rank signals
	const unsigned long long *my_flags;  // this rank's halo_flag[2]
	unsigned int *tickets;               // finished boundary CTAs, [lower, upper]
	unsigned long long hseq;             // sequence number of this stage launch (identical on every rank)
	SyncBlock *ranks[PFT_MAX_RANKS];     // stage 5: every rank's sync block (error-max broadcast)
	int rank, nranks;
	int wait_lo, wait_hi;                // this slab has a lower / an upper neighbour
	int nzc;                             // number of z-chunks (boundary chunks are scheduled first)
};

struct StageArgs {
	PeerArgs pr;
	const double *xu, *xp, *gl;                 // step-start state (u, p) and the static glass field
	const double *k1u, *k1p, *k3u, *k3p, *k4u, *k4p;
	const double *au, *ap;                      // stage 3: the stored stage input (aux3)
	double *ou, *op;                            // output pair
	double ha, hb;                              // stage 2: h3,h6   stage 4: h8   stage 5: h,h3
	double dir_u;                               // Dirichlet temperature at this stage's time level
	// stage 5 only
	double delta, h_abs, h_min, eps_scale;      // eps_scale = |h3| for DELTA_LOCAL else 1 (applied as a multiply only if != 1)
	int scale_eps, handle_nan;
	Ctl *ctl;
	HostSlot *slot;
	unsigned long long seq;
	unsigned int nblocks;
	// stage 1 only: run only if *cond != 0 (speculative launch before the host's decision)
	const int *cond;
#ifdef PFT_TRACE
	unsigned long long *trace;                  // development build: one timeline record per launch
	unsigned long long *trace_cta;              // and (start, end of body, end, smid<<32|chunk) per CTA of the first launches
#endif
};

struct KParams {
	Consts c;
	Geom g;
	StageArgs a;
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
// subroutine.  An exact rescaling of those was built and verified bit for bit, and rejected: its
// extra branches cost 3 % on every division, more than the tail costs; profiles/round1_tuning.md.)
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
__device__ __forceinline__ void load_state(const StageArgs &a, long long idx, double &su, double &sp, Raw<STAGE, HEAT> &r)
{
	if (STAGE == 1) {
		su = a.xu[idx];
		sp = a.xp[idx];
	} else if (STAGE == 3) {
		sp = a.ap[idx];
		su = HEAT ? a.au[idx] : a.xu[idx];
	} else if (STAGE == 2) {
		double xp = a.xp[idx], k1p = a.k1p[idx];
		sp = k1p * a.ha + xp;                                   // :388  q = v*h3 + w
		double xu = a.xu[idx];
		if (HEAT) { double k1u = a.k1u[idx]; su = k1u * a.ha + xu; if (KEEP) r.k1u = k1u; }
		else su = xu;
		if (KEEP) { r.xu = xu; r.xp = xp; r.k1p = k1p; }
	} else if (STAGE == 4) {
		double xp = a.xp[idx], k1p = a.k1p[idx], k3p = a.k3p[idx];
		sp = (k1p + 3.0 * k3p) * a.ha + xp;                     // :428
		double xu = a.xu[idx];
		if (HEAT) { double k1u = a.k1u[idx], k3u = a.k3u[idx]; su = (k1u + 3.0 * k3u) * a.ha + xu; }
		else su = xu;
	} else {
		double xp = a.xp[idx], k1p = a.k1p[idx], k3p = a.k3p[idx], k4p = a.k4p[idx];
		sp = (0.5 * k1p - 1.5 * k3p + 2.0 * k4p) * a.ha + xp;   // :449
		double xu = a.xu[idx];
		if (HEAT) {
			double k1u = a.k1u[idx], k3u = a.k3u[idx], k4u = a.k4u[idx];
			su = (0.5 * k1u - 1.5 * k3u + 2.0 * k4u) * a.ha + xu;
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
inline unsigned long long ld_acquire_sys(const unsigned long long *p) { return *p; }
inline void st_release_sys(unsigned long long *p, unsigned long long v) { *p = v; }
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

// chunk processed by this CTA: the two boundary chunks come first in launch order, so that
// their planes reach the neighbours while the interior is still being computed
__device__ __forceinline__ int chunk_of_block(int nzc)
{
	const int z = blockIdx.z;
	if (nzc < 3) return z;
	return z == 0 ? nzc - 1 : (z == 1 ? 0 : z - 1);      // the top chunk pushes its plane last: give it the head start
}

// ---- development build (-DPFT_TRACE): per-launch timeline from %globaltimer -----------------------
// record: [0] ~first CTA start  [1] last CTA start  [2] last CTA end  [3] last lower-boundary CTA past
// its wait  [4] same, upper  [5] longest wait (ns)  [6] lower flag published  [7] upper flag published
// [8] stage 5: error-max exchange starts  [9] decision published  [10] longest halo_signal (ns)
// [11..13] longest CTA body (ns) among lower-boundary / upper-boundary / interior chunks
#ifdef PFT_TRACE
__device__ __forceinline__ unsigned long long gtime()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
#define PFT_TRACE_MAX(a, i, v) do { if ((a).trace) atomicMax(&(a).trace[i], (unsigned long long)(v)); } while (0)
#else
#define PFT_TRACE_MAX(a, i, v) do { } while (0)
#endif

__device__ __forceinline__ void halo_wait(const PeerArgs &pr, bool lo, bool hi, int tid)
{
	const bool wl = lo && pr.wait_lo, wh = hi && pr.wait_hi;
	if (!(wl || wh)) return;
	if (tid == 0) {
		if (wl) while (ld_acquire_sys(&pr.my_flags[0]) + 1 < pr.hseq) { }
		if (wh) while (ld_acquire_sys(&pr.my_flags[1]) + 1 < pr.hseq) { }
	}
	__syncthreads();
}

__device__ __forceinline__ void halo_signal(const PeerArgs &pr, bool lo, bool hi, int tid)
{
	const bool wl = lo && pr.wait_lo, wh = hi && pr.wait_hi;
	if (!(wl || wh)) return;
	__syncthreads();                     // every thread's peer stores happen before thread 0's ticket (CTA barrier) ...
	if (tid == 0) {
		// ... a device-scope release per CTA, and ONE system-scope fence per side and launch, by the CTA that
		// takes the last ticket: causality is transitive (barrier -> device-scope release/acquire on the ticket
		// -> system-scope release of the flag).  A system fence in every boundary CTA queued up for 5-10 us
		// when the boundary CTAs of a launch finish together (measured, profiles/round1_tuning.md).
		__threadfence();
		const unsigned int n = gridDim.x * gridDim.y;
		if (wl && atomicAdd(&pr.tickets[0], 1u) == n - 1) { pr.tickets[0] = 0u; __threadfence_system(); st_release_sys(pr.lo_flag, pr.hseq); }
		if (wh && atomicAdd(&pr.tickets[1], 1u) == n - 1) { pr.tickets[1] = 0u; __threadfence_system(); st_release_sys(pr.hi_flag, pr.hseq); }
	}
}

// a skipped launch (speculative stage 1 of a rejected step) still has to advance the flags
__device__ __forceinline__ void halo_skip(const PeerArgs &pr)
{
	if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0) {
		if (pr.wait_lo) st_release_sys(pr.lo_flag, pr.hseq);
		if (pr.wait_hi) st_release_sys(pr.hi_flag, pr.hseq);
	}
}

// store one output pair value, and push it into the neighbours' ghost planes from the boundary planes
template <bool HEAT>
__device__ __forceinline__ void store_out(const StageArgs &a, long long o, long long col, int k, int n3, double vu, double vp)
{
	a.op[o] = vp;
	if (HEAT) a.ou[o] = vu;
	if (k == 0 && a.pr.lo_p) { a.pr.lo_p[col] = vp; if (HEAT) a.pr.lo_u[col] = vu; }
	if (k == n3 - 1 && a.pr.hi_p) { a.pr.hi_p[col] = vp; if (HEAT) a.pr.hi_u[col] = vu; }
}

// ---- stage 5 epilogue: grid-wide max of the error estimate, acceptance decision, host notification.
// Several GPUs: the last CTA of every rank's launch hands its slab maximum to every rank (peer stores,
// one thread per rank), waits for the others' and forms the global maximum -- the role of
// MPI_Allreduce(MAX), RK_MPI_SAsolver_hybrid2.c:572 -- so that every rank takes the same decision
// without a further launch.  Peers never wait for THIS CTA (their boundary waits concern CTAs that
// are long finished), so the spin cannot deadlock.
__device__ __forceinline__ void stage5_finish(const StageArgs &a, int tid, double emax, int nonfinite)
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
	if (tid == 0) {
		Ctl *ctl = a.ctl;
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

	if (a.pr.nranks > 1) {
		const int b = (int)(a.seq & 1ull);
		const int nthr = blockDim.x * blockDim.y;
		if (tid == 0) PFT_TRACE_MAX(a, 8, gtime());
		const double mine = __longlong_as_double((long long)s_bits);
		const int mine_nf = s_nonfinite;
		__syncthreads();                     // everyone has read the slab result before it is overwritten below
		for (int r = tid; r < a.pr.nranks; r += nthr) {
			EpsSlot *e = &a.pr.ranks[r]->slot[b][a.pr.rank];
			e->eps = mine;
			e->nonfinite = mine_nf;
			__threadfence_system();
			st_release_sys(&e->seq, a.seq);
		}
		if (tid == 0) { s_bits = 0ull; s_nonfinite = 0; }
		__syncthreads();
		for (int r = tid; r < a.pr.nranks; r += nthr) {
			const EpsSlot *e = &a.pr.ranks[a.pr.rank]->slot[b][r];
			while (ld_acquire_sys(&e->seq) != a.seq) { }
			const double v = *(volatile const double *)&e->eps;
			if (v > 0.0) atomicMax(&s_bits, (unsigned long long)__double_as_longlong(v));
			if (*(volatile const int *)&e->nonfinite) atomicOr(&s_nonfinite, 1);
		}
		__syncthreads();
	}
	if (tid == 0) {
		const double eps = __longlong_as_double((long long)s_bits);
		const int nfl = s_nonfinite;
		const double ecmp = a.scale_eps ? eps * a.eps_scale : eps;        // RK_MPI_SAsolver_hybrid2.c:578
		int acc = (ecmp < a.delta) || (a.h_abs < a.h_min);                 // :599
		if (a.handle_nan && nfl) acc = 0;                                   // :626-646
		a.ctl->accept = acc;
		__threadfence();
		HostSlot *s = a.slot;
		s->eps = eps;
		s->accept = acc;
		s->nonfinite = nfl;
		__threadfence_system();
		*(volatile unsigned long long *)&s->seq = a.seq;
		PFT_TRACE_MAX(a, 9, gtime());
	}
}

// MODEL: 0 GradP (calc_mode 0/10), 1 SigmaP1-P (1/11), 2 temperature model (2)
// HEAT : the heat equation is active (false for calc_mode 10/11: du/dt = 0, u is static)
// SHARE: each cell computes its three '+' face fluxes and takes the '-' ones from its neighbours
template <int STAGE, int MODEL, bool HEAT, bool SHARE>
__global__ void __launch_bounds__(512, 1) merson_stage(const __grid_constant__ KParams P)
{
	const Consts &c = P.c;
	const Geom &g = P.g;
	const StageArgs &a = P.a;

	pdl_launch();
	pdl_wait();
	if (STAGE == 1) { if (a.cond != nullptr && *a.cond == 0) { halo_skip(a.pr); return; } }

	constexpr int NF = HEAT ? 3 : 1;        // fields kept in the shared tile: p [, u, gl]
	constexpr bool FLUX = HEAT && SHARE;

	const int TX = blockDim.x, TY = blockDim.y;
	const int tx = threadIdx.x, ty = threadIdx.y;
	const int tid = ty * TX + tx;
	const int SP = TX + 2;                  // tile pitch incl. halo ring
	const int TS = SP * (TY + 2);           // one field of one buffer

	PFT_DYN_SMEM(double, smem);
	double *tile = smem;                                    // [2][NF][TS]
	double *fxs = smem + 2 * NF * TS;                       // [TY][TX+1]  (FLUX)
	double *fys = fxs + (FLUX ? TY * (TX + 1) : 0);         // [TY+1][TX]  (FLUX)

	const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
	const int i = i0 + tx, j = j0 + ty;
	const bool active = (i < g.n1) && (j < g.n2);
	const int ic = min(i, g.n1 - 1), jc = min(j, g.n2 - 1);
	const long long col = (long long)jc * g.px + ic;        // in-plane offset of this thread's column

	const int k0 = chunk_of_block(a.pr.nzc) * g.zc;
	const int k1 = min(g.n3, k0 + g.zc);
	const bool chunk_lo = (k0 == 0), chunk_hi = (k1 == g.n3);
	halo_wait(a.pr, chunk_lo, chunk_hi, ty * blockDim.x + tx);

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

	// face duty (FLUX): the tile's west and south boundary faces
	const bool fduty = FLUX && tid < TX + TY;

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
		load_state<STAGE, HEAT, true>(a, base + col, uc, pc, rc);
		gc = a.gl[base + col];
		if (k0 == 0 && g.front) { um = uc; pm = pc; gm = gc; }          // z-front mirror (equation.c:164-174)
		else {
			load_state<STAGE, HEAT, false>(a, base - g.plane + col, um, pm, rdummy);
			gm = a.gl[base - g.plane + col];
		}
		double *t0 = tile;
		t0[cs] = pc;
		if (HEAT) { t0[TS + cs] = uc; t0[2 * TS + cs] = gc; }
		if (hduty) {
			double hu, hp;
			load_state<STAGE, HEAT, false>(a, base + hcol, hu, hp, rdummy);
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
			un = a.dir_u; pn = pc; gn = gc;                           // equation.c:175-183 / :251-261
		} else {
			load_state<STAGE, HEAT, true>(a, base_n + col, un, pn, rn);
			gn = a.gl[base_n + col];
		}
		if (more && hduty) {
			load_state<STAGE, HEAT, false>(a, base_n + hcol, hu, hp, rdummy);
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
			double fe = face_flux(c, pc, pe, gc, ge, uc, ue);
			double fn = face_flux(c, pc, pnn, gc, gnn, uc, unn);
			double fw, fs;
			if (FLUX) {
				fxs[ty * (TX + 1) + tx + 1] = fe;
				fys[(ty + 1) * TX + tx] = fn;
				if (fduty) {
					if (tid < TY) {          // west boundary face of row `tid`
						int s0 = (tid + 1) * SP;
						fxs[tid * (TX + 1)] = face_flux(c, tp[s0], tp[s0 + 1], tg[s0], tg[s0 + 1], tu[s0], tu[s0 + 1]);
					} else {                 // south boundary face of column `tid-TY`
						int s0 = (tid - TY) + 1;
						fys[tid - TY] = face_flux(c, tp[s0], tp[s0 + SP], tg[s0], tg[s0 + SP], tu[s0], tu[s0 + SP]);
					}
				}
				__syncthreads();
				fw = fxs[ty * (TX + 1) + tx];
				fs = fys[ty * TX + tx];
			} else {
				const double uw = tu[cs - 1], gw = tg[cs - 1];
				const double us = tu[cs - SP], gs = tg[cs - SP];
				fw = face_flux(c, pw, pc, gw, gc, uw, uc);
				fs = face_flux(c, ps, pc, gs, gc, us, uc);
			}
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
				store_out<HEAT>(a, o, col, k, g.n3, du, dp);
			} else if (STAGE == 2) {
				const double vp = (rc.k1p + dp) * a.hb + rc.xp;                               // :408
				const double vu = HEAT ? (rc.k1u + du) * a.hb + rc.xu : 0.0;
				store_out<HEAT>(a, o, col, k, g.n3, vu, vp);
			} else {
				double e = fabs(0.2 * rc.k1p - 0.9 * rc.k3p + 0.8 * rc.k4p - 0.1 * dp);   // :521
				if (!(e <= 1.7976931348623157e308)) nonfinite = 1;
				emax = (e > emax) ? e : emax;
				const double vp = rc.xp + a.hb * (0.5 * (rc.k1p + dp) + 2.0 * rc.k4p);        // :667
				double vu = 0.0;
				if (HEAT) {
					double eu = fabs(0.2 * rc.k1u - 0.9 * rc.k3u + 0.8 * rc.k4u - 0.1 * du);
					if (!(eu <= 1.7976931348623157e308)) nonfinite = 1;
					emax = (eu > emax) ? eu : emax;
					vu = rc.xu + a.hb * (0.5 * (rc.k1u + du) + 2.0 * rc.k4u);
				}
				store_out<HEAT>(a, o, col, k, g.n3, vu, vp);
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

	halo_signal(a.pr, chunk_lo, chunk_hi, tid);
	if (STAGE == 5) stage5_finish(a, tid, emax, nonfinite);
}

// ----------------------------------------------------------------------------------------
// layout conversion between the reference's padded host layout (staged in device memory) and the
// compact device layout

// dir 0: padded -> compact (interior only), dir 1: compact -> padded (interior only)
__global__ void repack_kernel(double *padded, double *compact, int n1, int n2, int n3, int px, long long plane,
                              int N1, long long rowsize, int bt, int dir)
{
	const long long total = (long long)n1 * n2 * n3;
	for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		int i = (int)(t % n1);
		long long r = t / n1;
		int j = (int)(r % n2);
		int k = (int)(r / n2);
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

// sphere rasteriser (equation.c:508-529): gl = max(gl, 0.5*(1 - tanh(0.5/xi_gl*(|x-c| + 1e-10 - R))))
__global__ void rasterise_kernel(double *gl, const double *bx, const double *by, const double *bz, int count,
                                 int n1, int n2, int n3, int px, long long plane, int first_row, int total_n3,
                                 double L1, double L2, double L3, double xi_gl, double radius)
{
	PFT_DYN_SMEM(double, sb);           // staged sphere centres, `chunk` at a time
	const int chunk = blockDim.x;
	const long long total = (long long)n1 * n2 * n3;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = t < total;
	int i = 0, j = 0, k = 0;
	if (valid) { i = (int)(t % n1); long long r = t / n1; j = (int)(r % n2); k = (int)(r / n2); }
	const double x = L1 * (0.5 + i) / n1;
	const double y = L2 * (0.5 + j) / n2;
	const double z = L3 * (0.5 + k + first_row) / total_n3;
	const long long ic = (long long)(k + 1) * plane + (long long)j * px + i;
	double v = valid ? gl[ic] : 0.0;
	const double s = 0.5 / xi_gl;
	// tanh(a) rounds to exactly 1.0 for a > 19.07, i.e. the profile of a sphere further away than
	// R + 19.5/s is exactly 0 and cannot raise gl >= 0: skip its sqrt/tanh (result-neutral, bit for bit)
	const double cut = radius + 19.5 / s;
	const double cut2 = cut * cut;
	for (int q0 = 0; q0 < count; q0 += chunk) {
		int q = q0 + threadIdx.x;
		__syncthreads();
		if (q < count) { sb[threadIdx.x] = bx[q]; sb[chunk + threadIdx.x] = by[q]; sb[2 * chunk + threadIdx.x] = bz[q]; }
		__syncthreads();
		int m = min(chunk, count - q0);
		for (int qq = 0; qq < m; ++qq) {
			double dx = x - sb[qq], dy = y - sb[chunk + qq], dz = z - sb[2 * chunk + qq];
			double d2 = dx * dx + dy * dy + dz * dz;
			if (d2 > cut2) continue;
			double d = sqrt(d2) + 1E-10;
			double f = 0.5 * (1.0 - tanh(s * (d - radius)));
			v = (v < f) ? f : v;
		}
	}
	if (valid) gl[ic] = v;
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
