// pft_kernels_tma.cuh -- the TMA-staged Merson stage kernel (sm_100a, float64, HEAT modes 0/1/2).
//
// Same arithmetic and the same five fused stages as `merson_stage` in pft_kernels.cuh (see the
// header comment there for the reference citations); what changes is how HBM reaches the SM:
//
//   * every array a stage reads (x, K1, K3, K4 pairs or aux3, plus gl) has a 3-D tensor map over
//     its (px, n2, n3+2) layout; ONE elected thread issues one `cp.async.bulk.tensor.3d` per array
//     and z-plane, box = (TX+4) x (TY+2) x 1 (the inner start coordinate must be 16-byte aligned,
//     measured on B200: an odd fp64 x-coordinate raises an illegal-instruction fault), so the tile arrives in shared memory together with
//     its one-cell halo ring, without per-thread address arithmetic, without halo-duty loads and
//     without holding in-flight loads in registers (SASS: UTMALDG);
//   * three raw stage buffers rotate (plane k = centre, k+1 = being formed, k+2 = in flight);
//     an mbarrier per buffer (expect_tx / complete_tx) tells the consumers when a plane landed,
//     the CTA barrier that ends an iteration doubles as the "buffer free" signal;
//   * the Merson stage input (x + h*sum a_i K_i) of plane k+1 is formed from the raw buffer right
//     before plane k is computed: the centre value goes to the register window (z-1/z/z+1) and,
//     with the ring cells, to a double-buffered state tile that plane k+1 will read its in-plane
//     neighbours from; stages 1 and 3 read the raw buffer directly (their input is stored as is);
//   * the operands of the output combination (x, K1, K3, K4 at the centre cell) are re-read from
//     the raw buffer instead of being carried in registers across the iteration.
//
// Mirror boundaries: TMA zero-fills out-of-range box elements, so neighbour slots are clamped to
// the domain instead (value-identical to the reference's mirrored ghost nodes).
#pragma once
#include "pft_kernels.cuh"

#ifdef PFT_CUDA_EMU
#include <cuda.h>           // tests/cuda_emu/cuda.h: emulated CUtensorMap
#else
#include <cuda.h>
#endif

namespace pft {

constexpr int TMA_MAX_ARRAYS = 9;
constexpr int TMA_STAGES = 3;

struct alignas(64) TmaParams {
	CUtensorMap maps[TMA_MAX_ARRAYS];   // order: p-type arrays, u-type arrays, gl (see stage_np); x = pair A_X
	CUtensorMap xalt[2];                 // the other state pair (A_XN): p, u -- used when StepCtl.cur == 1
	CUtensorMap k1alt[2];                // speculative stage 1: the K1 buffer that goes with pair A_XN: p, u
	Consts c;
	Geom g;
	StageArgs a;
	int bx, by;                          // box extents: bx = TX+4 (TX even; the box starts at the even x = i0-2), by = TY+2
	int as;                              // doubles per array buffer (bx*by rounded up to 16 = 128 B)
	int tx, ty;                          // merson_stage_v2 (1-D thread block): threads per tile row / thread rows
	unsigned smem_bytes;                 // dynamic shared memory of the launch
};

// number of p-type (and u-type) arrays stage S stages through shared memory
__host__ __device__ constexpr int stage_np(int S) { return S == 1 ? 1 : S == 2 ? 2 : S == 3 ? 1 : S == 4 ? 3 : 4; }

#ifndef PFT_CUDA_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The wait carries a suspend-time hint: a warp whose plane has not landed sleeps in hardware (NANOSLEEP.SYNCS)
// instead of re-issuing try_wait.  In the barrier-free stages 1 and 3 a warp that runs ahead of its CTA waits at
// almost every plane (SYNCS + BRA were ~12 % of the instructions in ncu's source view); measured, the hint is
// neutral (61.2-61.3 Gcell-RHS/s either way, profiles/r02_mbar_hint_ab.txt) -- those spins sat in idle issue slots.
#ifndef PFT_MBAR_SUSPEND_NS
#define PFT_MBAR_SUSPEND_NS 1000000u
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"PFT_MBAR_WAIT:\n"
#ifdef PFT_MBAR_NO_HINT
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#else
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
#endif
		"@p bra PFT_MBAR_DONE;\n"
		"bra PFT_MBAR_WAIT;\n"
		"PFT_MBAR_DONE:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(PFT_MBAR_SUSPEND_NS) : "memory");
}
// descriptor fetch ahead of the first tensor load (the maps live in the kernel's parameter block)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
	asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
		::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
		: "memory");
}
#else
// CPU emulation (tests/cuda_emu): a tensor load is a synchronous box copy with zero fill; an
// mbarrier is {tx bytes pending [63:44], expected arrivals [43:32], pending arrivals [31:20], phase [19:0]}
inline void emu_mbar_settle(uint64_t *bar)
{
	uint64_t tx = *bar >> 44, cnt = (*bar >> 32) & 0xfff, pend = (*bar >> 20) & 0xfff, ph = *bar & 0xfffff;
	if (pend == 0 && tx == 0) { ph = (ph + 1) & 0xfffff; pend = cnt; }
	*bar = (tx << 44) | (cnt << 32) | (pend << 20) | ph;
}
inline void mbar_init(uint64_t *bar, uint32_t count) { *bar = ((uint64_t)count << 32) | ((uint64_t)count << 20); }
inline void fence_mbar_init() {}
inline void mbar_arrive(uint64_t *bar) { *bar -= (1ull << 20); emu_mbar_settle(bar); }
inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { *bar += ((uint64_t)bytes << 44); *bar -= (1ull << 20); emu_mbar_settle(bar); }
inline void mbar_wait(uint64_t *bar, uint32_t parity)
{
	while ((uint32_t)(*bar & 1u) == parity) pft_emu::yield();
}
inline void tma_prefetch_desc(const CUtensorMap *) {}
inline void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
	uint32_t n = pft_emu::tensor_load_3d(dst, map, x, y, z);
	*bar -= ((uint64_t)n << 44);
	emu_mbar_settle(bar);
}
#endif

// stage input from the raw values of one cell: r[0] = x (or aux3), r[1..] = K1, K3, K4
template <int STAGE>
__device__ __forceinline__ double form_state(const double *buf, int as, int slot, int first, double ha)
{
	const double *r = buf + first * as + slot;
	if (STAGE == 1 || STAGE == 3) return r[0];
	if (STAGE == 2) return r[as] * ha + r[0];                                               // hybrid2.c:388
	if (STAGE == 4) return (r[as] + 3.0 * r[2 * as]) * ha + r[0];                           // :428
	return (0.5 * r[as] - 1.5 * r[2 * as] + 2.0 * r[3 * as]) * ha + r[0];                   // :449
}

// everything of the right-hand side of one cell that comes after the face fluxes
template <int MODEL>
__device__ __forceinline__ void cell_rhs(const Consts &c, double uc, double pc, double gc,
                                         double pw, double pe, double ps, double pnn, double pm, double pn,
                                         double sum, double &du, double &dp)
{
	const double rho = mix3(pc, gc, c.g_rho, c.i_rho, c.w_rho);
	const double cp = mix3(pc, gc, c.g_cp, c.i_cp, c.w_cp);
	const double wi = 1.0 - c.zeta * gc;
	const double ind = (wi > 0.0) ? wi : 0.0;                                                // fmax(0, .), equation.c:362
	if (MODEL == 2) {
		double ch = cosh(c.gamma * (uc - c.u_star));                                           // equation.c:835-867
		double dp_du = (-0.5 * c.gamma / (ch * ch)) * ind;
		du = div_pos(sum, rho * (cp - c.L * dp_du));
		dp = dp_du * du;
	} else {
		dp = c.h1_2 * (-(-pw + pc) + (-pc + pe))                                               // equation.c:659-669
		   + c.h2_2 * (-(-ps + pc) + (-pc + pnn))
		   + c.h3_2 * (-(-pm + pc) + (-pc + pn));
		double one_m_p = 1.0 - pc;
		double well = c.xi_2_inv_a * pc * one_m_p * (pc - 0.5);
		double du_star = uc - c.u_star;
		if (MODEL == 0) {
			double gx = c.h1_d2 * (-pw + pe), gy = c.h2_d2 * (-ps + pnn), gz = c.h3_d2 * (-pm + pn);
			double nrm = sqrt(gx * gx + gy * gy + gz * gz) + 1E-10;                           // :332-336
			dp += well - c.bam * nrm * du_star;                                                   // :370
		} else {
			double pq = pc * one_m_p;
			double pqm = (pq >= 0.0) ? pq : 0.0;                                                  // fmax(., 0), :387
			dp += well - c.sam * sshape(c, pc) * sshape(c, one_m_p) * pqm * du_star;
		}
		dp = div_pos(dp, c.alpha);
		dp *= ind;
		du = div_pos(div_pos(sum, rho) + c.L * dp, cp);                                           // :711-723
	}
}

// the fused combination that follows stage STAGE, for one cell (operands from the raw buffer)
template <int STAGE>
__device__ __forceinline__ void stage_output(const OutPtrs &a, double hb, const double *rb, int AS, int IU, int cs, long long o,
                                             long long col, int k, int n3, double du, double dp, double &emax, int &nonfinite)
{
	if (STAGE == 1 || STAGE == 3 || STAGE == 4) {
		store_out<true>(a, o, col, k, n3, du, dp);
	} else if (STAGE == 2) {
		const double vp = (rb[AS + cs] + dp) * hb + rb[cs];                                     // hybrid2.c:408
		const double vu = (rb[(IU + 1) * AS + cs] + du) * hb + rb[IU * AS + cs];
		store_out<true>(a, o, col, k, n3, vu, vp);
	} else {
		const double xp = rb[cs], k1p = rb[AS + cs], k3p = rb[2 * AS + cs], k4p = rb[3 * AS + cs];
		const double xu = rb[IU * AS + cs], k1u = rb[(IU + 1) * AS + cs], k3u = rb[(IU + 2) * AS + cs], k4u = rb[(IU + 3) * AS + cs];
		double e = fabs(0.2 * k1p - 0.9 * k3p + 0.8 * k4p - 0.1 * dp);                          // :521
		double eu = fabs(0.2 * k1u - 0.9 * k3u + 0.8 * k4u - 0.1 * du);
		if (!(e <= 1.7976931348623157e308) || !(eu <= 1.7976931348623157e308)) nonfinite = 1;
		emax = (e > emax) ? e : emax;
		emax = (eu > emax) ? eu : emax;
		const double vp = xp + hb * (0.5 * (k1p + dp) + 2.0 * k4p);                           // :667
		const double vu = xu + hb * (0.5 * (k1u + du) + 2.0 * k4u);
		store_out<true>(a, o, col, k, n3, vu, vp);
	}
}

// the values a stage stores for one cell (the fused combination that follows the stage), and the error estimate
template <int STAGE>
__device__ __forceinline__ void stage_values(double hb, const double *rb, int AS, int IU, int cs, double du, double dp,
                                             double &emax, int &nonfinite, double &vu, double &vp)
{
	if (STAGE == 1 || STAGE == 3 || STAGE == 4) {
		vu = du; vp = dp;
	} else if (STAGE == 2) {
		vp = (rb[AS + cs] + dp) * hb + rb[cs];                                                 // hybrid2.c:408
		vu = (rb[(IU + 1) * AS + cs] + du) * hb + rb[IU * AS + cs];
	} else {
		const double xp = rb[cs], k1p = rb[AS + cs], k3p = rb[2 * AS + cs], k4p = rb[3 * AS + cs];
		const double xu = rb[IU * AS + cs], k1u = rb[(IU + 1) * AS + cs], k3u = rb[(IU + 2) * AS + cs], k4u = rb[(IU + 3) * AS + cs];
		double e = fabs(0.2 * k1p - 0.9 * k3p + 0.8 * k4p - 0.1 * dp);                          // :521
		double eu = fabs(0.2 * k1u - 0.9 * k3u + 0.8 * k4u - 0.1 * du);
		if (!(e <= 1.7976931348623157e308) || !(eu <= 1.7976931348623157e308)) nonfinite = 1;
		emax = (e > emax) ? e : emax;
		emax = (eu > emax) ? eu : emax;
		vp = xp + hb * (0.5 * (k1p + dp) + 2.0 * k4p);                                          // :667
		vu = xu + hb * (0.5 * (k1u + du) + 2.0 * k4u);
	}
}

template <int STAGE, int MODEL>
__global__ void __launch_bounds__(256, (STAGE == 5) ? 2 : 3) merson_stage_tma(const __grid_constant__ TmaParams P)
{
	const Consts &c = P.c;
	const Geom &g = P.g;
	const StageArgs &a = P.a;

	pdl_launch();

	constexpr int NP = stage_np(STAGE);
	constexpr int NA = 2 * NP + 1;
	constexpr bool FORM = !(STAGE == 1 || STAGE == 3);
	constexpr int IU = NP, IG = 2 * NP;          // first u-type array, the gl array

	const int TX = blockDim.x, TY = blockDim.y;
	const int tx = threadIdx.x, ty = threadIdx.y;
	const int tid = ty * TX + tx;
	const int BX = P.bx, AS = P.as;

	PFT_DYN_SMEM_ALIGNED(double, smem_t);
	double *raw = smem_t;                                            // [3][NA][AS]
	double *tile = raw + TMA_STAGES * NA * AS;                       // [2][2][AS]  (FORM)
	__shared__ uint64_t full[TMA_STAGES];    // raw plane landed (TMA complete_tx)

	const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
	const int i = i0 + tx, j = j0 + ty;
	const bool active = (i < g.n1) && (j < g.n2);
	const int cs = (ty + 1) * BX + (tx + 2);     // the box starts at x = i0-2: TMA needs a 16-byte aligned inner offset
	// in-plane neighbour slots, clamped to the DOMAIN (zero-Neumann mirror, equation.c:142-160)
	const int sW = (i > 0) ? cs - 1 : cs;
	const int sE = (i < g.n1 - 1) ? cs + 1 : cs;
	const int sS = (j > 0) ? cs - BX : cs;
	const int sN = (j < g.n2 - 1) ? cs + BX : cs;

	const int k0 = chunk_of_block(a.pr.nzc) * g.zc;
	const int k1 = min(g.n3, k0 + g.zc);
	const bool chunk_lo = (k0 == 0), chunk_hi = (k1 == g.n3);
	// chunk-local plane index q: z = k0 - 1 + q, array plane index = k0 + q
	const int qlo = (k0 == 0 && g.front) ? 1 : 0;                    // z-front mirror: plane -1 is never fetched
	const int qhi = (k1 == g.n3 && g.rear) ? (k1 - k0) : (k1 - k0 + 1);   // z-rear ghost is Dirichlet/mirror

	// ring duty (FORM stages): the first 2*TX+2*TY threads each form one ring cell per plane
	const bool hduty = FORM && tid < 2 * TX + 2 * TY;
	int hs = 0;
	if (hduty) {
		int hx, hy;
		if (tid < TX) { hx = tid; hy = -1; }
		else if (tid < 2 * TX) { hx = tid - TX; hy = TY; }
		else if (tid < 2 * TX + TY) { hx = -1; hy = tid - 2 * TX; }
		else { hx = TX; hy = tid - 2 * TX - TY; }
		hs = (hy + 1) * BX + (hx + 2);
	}

	if (tid == 0) {
		for (int b = 0; b < TMA_STAGES; b++) mbar_init(&full[b], 1);
		fence_mbar_init();
	}
	__syncthreads();
	pdl_wait();                                   // from here on this launch reads what its predecessor wrote
	StageScalars sc;
	if (!stage_begin<STAGE>(a, sc)) return;
	if (STAGE == 1 && !a.standalone && ctl_ld(&a.ctl->skip1)) { halo_skip(a.pr, sc.hseq); return; }
	const OutPtrs w = out_ptrs<STAGE>(a, sc.cur);
	const double ha = sc.ha;
	halo_wait(a.pr, sc.hseq, chunk_lo, chunk_hi, tid);     // peer path: the neighbours' planes of the previous launch have landed

	// thread 0 requests plane q into raw buffer b: one tensor copy per array.  The state pair x is
	// array 0 (p) and array IU (u) of every stage but the third; which pair holds x is StepCtl.cur.
	const uint32_t plane_bytes = (uint32_t)(NA * BX * P.by * sizeof(double));
	const bool alt = (STAGE != 3) && sc.cur != 0;
	auto issue = [&](int q, int b) {
		mbar_arrive_expect_tx(&full[b], plane_bytes);
		#pragma unroll
		for (int m = 0; m < NA; m++) {
			const CUtensorMap *map = &P.maps[m];
			if (alt && m == 0) map = &P.xalt[0];
			if (alt && m == IU) map = &P.xalt[1];
			tma_load_3d(raw + (b * NA + m) * AS, map, i0 - 2, j0 - 1, k0 + q, &full[b]);
		}
	};
	int next_q = qlo;
	if (tid == 0) {
		#pragma unroll 1
		for (int n = 0; n < TMA_STAGES && next_q <= qhi; n++, next_q++) issue(next_q, next_q % TMA_STAGES);
	}

	uint32_t full_bits = 0;                      // parity of the next completion to wait for, per buffer
	double um, pm, gm, uc, pc, gc, un, pn, gn;
	double fzm = 0.0, emax = 0.0;
	int nonfinite = 0;

	// ---- prologue: planes k0-1 (registers only) and k0 (registers + tile 0)
	if (qlo == 0) {
		mbar_wait(&full[0], 0); full_bits ^= 1u;
		pm = form_state<STAGE>(raw, AS, cs, 0, ha);
		um = form_state<STAGE>(raw, AS, cs, IU, ha);
		gm = raw[IG * AS + cs];
	}
	{
		mbar_wait(&full[1], 0); full_bits ^= 2u;
		const double *rb = raw + NA * AS;
		pc = form_state<STAGE>(rb, AS, cs, 0, ha);
		uc = form_state<STAGE>(rb, AS, cs, IU, ha);
		gc = rb[IG * AS + cs];
		if (FORM) {
			tile[cs] = pc; tile[AS + cs] = uc;
			if (hduty) { tile[hs] = form_state<STAGE>(rb, AS, hs, 0, ha); tile[AS + hs] = form_state<STAGE>(rb, AS, hs, IU, ha); }
		}
		if (qlo == 1) { um = uc; pm = pc; gm = gc; }                  // equation.c:164-174
		fzm = face_flux(c, pm, pc, gm, gc, um, uc);
	}
	__syncthreads();

	int bq = 1, bn = 2, bf = 0;  // raw buffers of plane q (centre), q+1 (next), q+2 (to be requested)
	int tc = 0, tn = 1;          // state tiles of plane k and k+1
	for (int it = 0, k = k0; k < k1; ++k, ++it) {
		const int q = it + 1;
		const bool more = (k + 1 < k1);
		const double *rb = raw + bq * NA * AS;                        // raw planes of the centre plane k
		const double *rg = rb + IG * AS;

		// ---- thread 0: request plane q+2 into the buffer plane q-1 just vacated
		if (tid == 0 && next_q <= qhi && next_q <= q + 2) {
			issue(next_q, bf);
			next_q++;
		}

		// ---- plane k+1: form the stage input (centre -> registers, centre + ring -> its state tile)
		if (q + 1 <= qhi) {
			mbar_wait(&full[bn], (full_bits >> bn) & 1u);
			full_bits ^= (1u << bn);
			const double *rn = raw + bn * NA * AS;
			pn = form_state<STAGE>(rn, AS, cs, 0, ha);
			un = form_state<STAGE>(rn, AS, cs, IU, ha);
			gn = rn[IG * AS + cs];
			if (FORM && more) {
				double *tw = tile + tn * 2 * AS;
				tw[cs] = pn; tw[AS + cs] = un;
				if (hduty) { tw[hs] = form_state<STAGE>(rn, AS, hs, 0, ha); tw[AS + hs] = form_state<STAGE>(rn, AS, hs, IU, ha); }
			}
		} else {
			un = sc.dir_u; pn = pc; gn = gc;                            // equation.c:175-183 / :251-261
		}

		// ---- plane k
		const double *tp = FORM ? tile + tc * 2 * AS : rb;
		const double *tu = FORM ? tp + AS : rb + IU * AS;
		const double pw = tp[sW], pe = tp[sE], ps = tp[sS], pnn = tp[sN];

		double sum;
		{
			const double ue = tu[sE], ge = rg[sE];
			const double unn = tu[sN], gnn = rg[sN];
			double fe = face_flux(c, pc, pe, gc, ge, uc, ue);
			double fn = face_flux(c, pc, pnn, gc, gnn, uc, unn);
			double fw = face_flux(c, pw, pc, rg[sW], gc, tu[sW], uc);
			double fs = face_flux(c, ps, pc, rg[sS], gc, tu[sS], uc);
			double fzp = face_flux(c, pc, pn, gc, gn, uc, un);
			sum = c.h1_2 * (-fw + fe) + c.h2_2 * (-fs + fn) + c.h3_2 * (-fzm + fzp);   // equation.c:711-720
			fzm = fzp;
		}

		double du, dp;
		cell_rhs<MODEL>(c, uc, pc, gc, pw, pe, ps, pnn, pm, pn, sum, du, dp);

		// ---- fused combination that follows this stage (operands re-read from the raw buffer)
		if (active) {
			const long long col = (long long)j * g.px + i;
			stage_output<STAGE>(w, sc.hb, rb, AS, IU, cs, (long long)(k + 1) * g.plane + col, col, k, g.n3, du, dp, emax, nonfinite);
		}

		// ---- this plane is done: raw[bq] may be refilled, tile[tc] overwritten (two planes later)
		__syncthreads();
		um = uc; pm = pc; gm = gc;
		uc = un; pc = pn; gc = gn;
		{ const int t = bq; bq = bn; bn = bf; bf = t; }
		if (FORM) { tc = tn; tn ^= 1; }
	}

	halo_signal(a.pr, sc.hseq, chunk_lo, chunk_hi, tid);
	if (STAGE == 5) stage5_finish(a, tid, emax, nonfinite, smem_t, P.smem_bytes);
}

}  // namespace pft
