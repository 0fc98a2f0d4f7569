// pft_kernels_v2.cuh -- second generation of the TMA-staged Merson stage kernel (sm_100a, float64).
//
// Same arithmetic, same fused stages, same TMA box pipeline and the same control-block protocol as
// merson_stage_tma (pft_kernels_tma.cuh); what changes is the instruction stream around the fp64 work.
// Measured on B200 (profiles/): a stage's time is the issue-slot count of its inner loop -- an fp64
// instruction holds the issue port for two cycles, everything else for one, and 2*fp64 + other predicts
// the kernel time of stages 1-3 to 3 %.  Roughly half of the 350-490 instructions per cell and plane of
// the first generation were not fp64 arithmetic.  Hence:
//
//   * R vertically adjacent cells per thread (template R): the y-face between two cells of a thread is
//     evaluated once (exactly: the reference evaluates a face with the same operands from either side,
//     equation.c:712-719), the cells' y-neighbours come from registers instead of shared memory, and loop
//     control, barrier, mbarrier wait and address set-up are paid once per R cells;
//   * the z-march is unrolled by three, the depth of the raw-buffer ring: buffer and tile indices are
//     compile-time constants in every unrolled body, so the register window rotates by renaming instead of
//     ~30 MOVs per plane, and shared-memory addresses are `base register + immediate`;
//   * tile shapes the library instantiates (CTX x CTY threads) make box pitch and array stride immediates;
//   * stages 1 and 3 read only TMA-landed raw planes -- nothing a warp writes is read by another warp --
//     so they run WITHOUT any CTA barrier in the plane loop (NOBAR): a warp that leaves a plane takes a
//     ticket on the plane's raw buffer, and the warp that takes the last ticket requests the next plane
//     into it.  Nobody ever waits for a slower warp; warps drift up to two planes apart.
//
// Bit-exactness: the per-cell operation order is untouched; tests/ hold every variant against the oracle.
#pragma once
#include "pft_kernels_tma.cuh"

#include <type_traits>

namespace pft {

template <int STAGE, int MODEL, int R, int CTX, int CTY, bool NOBAR>
__global__ void __launch_bounds__(CTX > 0 ? ((CTX * CTY + 31) & ~31) : 256, (STAGE == 5 || R > 1) ? 2 : 3)
merson_stage_v2(const __grid_constant__ TmaParams P)
{
	// NOBAR: stages 1 and 3 (they read raw planes only) run without any CTA barrier in the plane loop; the forming stages
	// 2, 4, 5 replace their per-plane CTA barrier by an arrive / wait pair (SPLIT below)
	const Consts &c = P.c;
	const Geom &g = P.g;
	const StageArgs &a = P.a;

	pdl_launch();

	constexpr int NP = stage_np(STAGE);
	constexpr int NA = 2 * NP + 1;
	constexpr bool FORM = !(STAGE == 1 || STAGE == 3);
	constexpr bool SPLIT = FORM && NOBAR;        // forming stage with the split barrier
	constexpr bool FREE = !FORM && NOBAR;        // barrier-free raw-plane stage
	constexpr int IU = NP, IG = 2 * NP;          // first u-type array, the gl array

	const int TX = CTX > 0 ? CTX : P.tx, TY = CTX > 0 ? CTY : P.ty;   // threads; the tile is TX x (TY*R) cells
	const int RY = TY * R;
	const int BX = CTX > 0 ? CTX + 4 : P.bx;
	const int AS = CTX > 0 ? (((CTX + 4) * (CTY * R + 2) + 15) & ~15) : P.as;
	const int RS = NA * AS;                                          // one raw plane (all arrays)
	const int NC = TX * TY;                                          // threads
	const int tid = threadIdx.x;
	const int tx = tid % TX, ty = tid / TX;
	const int lane = tid & 31;
	const unsigned nwarps = (unsigned)((NC + 31) >> 5);

	PFT_DYN_SMEM_ALIGNED(double, smem_t);
	double *raw = smem_t;                                            // [3][NA][AS]
	double *tile = raw + TMA_STAGES * RS;                            // [3][2][AS]  (FORM)
	__shared__ uint64_t full[TMA_STAGES];     // raw plane landed (TMA complete_tx)
	__shared__ unsigned int left[TMA_STAGES]; // NOBAR: warps that have left the plane in this buffer
	__shared__ uint64_t formed;               // split barrier of the forming stages: every warp has written its part of a state tile
	__shared__ unsigned int formed_tick;      // ... and the ticket that elects the last of them (it requests the next raw plane)

	const int i0 = blockIdx.x * TX, j0 = blockIdx.y * RY;
	const int i = i0 + tx, jb = j0 + ty * R;                         // this thread's column, its lowest row
	const int cs = (ty * R + 1) * BX + (tx + 2);                     // slot of its lowest cell; the box starts at x = i0-2
	// in-plane neighbour offsets, clamped to the DOMAIN (zero-Neumann mirror, equation.c:142-160)
	const int dW = (i > 0) ? -1 : 0, dE = (i < g.n1 - 1) ? 1 : 0;
	const int sS = (jb > 0) ? cs - BX : cs;                          // below the lowest cell
	const int jt = jb + R - 1;                                       // the highest row
	// rows beyond the domain (ragged tiles): a cell whose upper neighbour row is outside mirrors itself
	int sN;                                                          // above the highest cell that exists
	{
		const int jlast = min(jt, g.n2 - 1);
		const int slast = cs + (jlast - jb) * BX;
		sN = (jlast < g.n2 - 1) ? slast + BX : slast;
	}

	const int k0 = chunk_of_block(a.pr.nzc) * g.zc;
	const int k1 = min(g.n3, k0 + g.zc);
	const bool chunk_lo = (k0 == 0), chunk_hi = (k1 == g.n3);
	// chunk-local plane index q: z = k0 - 1 + q, array plane index = k0 + q
	const int qlo = (k0 == 0 && g.front) ? 1 : 0;                    // z-front mirror: plane -1 is never fetched
	const int qhi = (k1 == g.n3 && g.rear) ? (k1 - k0) : (k1 - k0 + 1);   // z-rear ghost is Dirichlet/mirror

	// ring duty (FORM stages): the first 2*TX + 2*RY threads each form one ring cell per plane
	const bool hduty = FORM && tid < 2 * TX + 2 * RY;
	int hs = 0;
	if (hduty) {
		int hx, hy;
		if (tid < TX) { hx = tid; hy = -1; }
		else if (tid < 2 * TX) { hx = tid - TX; hy = RY; }
		else if (tid < 2 * TX + RY) { hx = -1; hy = tid - 2 * TX; }
		else { hx = TX; hy = tid - 2 * TX - RY; }
		hs = (hy + 1) * BX + (hx + 2);
	}

	if (tid == 0) {
		for (int b = 0; b < TMA_STAGES; b++) { mbar_init(&full[b], 1); left[b] = 0u; }
		mbar_init(&formed, nwarps); formed_tick = 0u;
		fence_mbar_init();
	}
	// the tensor-map descriptors are fetched while this CTA waits for its predecessors (warp 2: warps 0 and 1 poll)
	if (tid >= 64 && tid < 64 + NA) tma_prefetch_desc(&P.maps[tid - 64]);
	if (STAGE != 3 && tid >= 64 + NA && tid < 64 + NA + 2) tma_prefetch_desc(&P.xalt[tid - 64 - NA]);
	if (NP > 1 && a.spec && tid >= 64 + NA + 2 && tid < 64 + NA + 4) tma_prefetch_desc(&P.k1alt[tid - 64 - NA - 2]);
	__syncthreads();
	// from here on this launch reads what its predecessor wrote: the whole predecessor (stage 1, or data flow off),
	// or the CTAs of the stage before that surround this one
	const int my_chunk = chunk_of_block(a.pr.nzc);
	const bool s1spec = (STAGE == 1) && a.s1spec;       // S1': K1 of the next attempt from stage 5's candidate (pft_ctl.h)
	const bool signals = (STAGE != 5) || a.spec;        // stage 5 has a successor within the attempt only then
	if ((STAGE == 1 && !s1spec) || !a.dataflow || a.standalone) pdl_wait();
	else {
		if (!a.spec) dep_wait(a, my_chunk, a.pr.nzc, tid);
		else if (!ctl_dep_wait(a, my_chunk, a.pr.nzc, tid)) { if (signals) dep_signal(a, my_chunk, tid); return; }
	}
	StageScalars sc;
	if (!stage_begin<STAGE>(a, sc)) { if (signals) dep_signal(a, my_chunk, tid); return; }
	if (STAGE == 1 && !a.standalone && !s1spec && ctl_ld(&ctl_block(a)->skip1)) { halo_skip(a.pr, sc.hseq); dep_signal(a, my_chunk, tid); return; }
	if (STAGE == 1 && a.s1spec == 1 && candidate_rejected(a, sc.cur, tid)) {
		// the decision overtook this CTA and turned the candidate down: its K1 will never be read.  (Early CTAs of S1'
		// run before the decision exists and evaluate in vain: a rejected attempt costs part of an S1', not all of it.)
		halo_signal(a.pr, sc.hseq, chunk_lo, chunk_hi, tid);
		dep_signal(a, my_chunk, tid);
		return;
	}
	const OutPtrs w = out_ptrs<STAGE>(a, sc.cur);
	const double ha = sc.ha;
	halo_wait(a.pr, sc.hseq, chunk_lo, chunk_hi, tid);     // peer path: the neighbours' planes of the previous launch have landed

	// plane q goes to raw buffer q % 3: one tensor copy per array.  The state pair x is array 0 (p) and
	// array IU (u) of every stage but the third; which pair holds x is StepCtl.cur.
	const uint32_t plane_bytes = (uint32_t)(NA * BX * (RY + 2) * sizeof(double));
	// (S1' reads the candidate pair; with speculative stage 1 K1 -- array 1 of stages 2, 4, 5 -- has a buffer per pair)
	const bool alt = (STAGE != 3) && (sc.cur ^ (s1spec ? 1 : 0)) != 0;
	const bool altk = (STAGE == 2 || STAGE == 4 || STAGE == 5) && a.spec && sc.cur != 0;
	auto issue = [&](int q, int b) {
		mbar_arrive_expect_tx(&full[b], plane_bytes);
		#pragma unroll
		for (int m = 0; m < NA; m++) {
			const CUtensorMap *map = &P.maps[m];
			if (alt && m == 0) map = &P.xalt[0];
			if (alt && m == IU) map = &P.xalt[1];
			if (NP > 1 && altk && m == 1) map = &P.k1alt[0];
			if (NP > 1 && altk && m == IU + 1) map = &P.k1alt[1];
			tma_load_3d(raw + b * RS + m * AS, map, i0 - 2, j0 - 1, k0 + q, &full[b]);
		}
	};
	int next_q = qlo;
	if (tid == 0) {
		#pragma unroll 1
		for (int n = 0; n < TMA_STAGES && next_q <= qhi; n++, next_q++) issue(next_q, next_q % TMA_STAGES);
	}
	// NOBAR: this warp has read everything it needs of plane q in raw buffer b.  The last warp to say so requests
	// plane q+3 into the buffer (the fence orders the warp's reads before the asynchronous refill).
	auto leave = [&](int q, int b) {
		__syncwarp();
		if (lane == 0) {
			__threadfence_block();
			if (atomicAdd(&left[b], 1u) == nwarps - 1) {
				left[b] = 0u;
				if (q + TMA_STAGES <= qhi) issue(q + TMA_STAGES, b);
			}
		}
	};

	// Forming stages (2, 4, 5) with NOBAR: the CTA barrier at the end of every plane becomes an arrive / wait pair.
	// A warp ARRIVES when its part of the state tile of plane k+1 is written (and with that it has finished plane k-1),
	// computes plane k, and WAITS at the top of the next plane until every warp has arrived -- warps drift up to one plane
	// apart instead of meeting after each.  The last warp to arrive requests raw plane q+2 into the buffer of plane q-1
	// (every warp is past it).  Three state tiles: the tile a warp overwrites was last read two planes ago, and the wait
	// it has just passed says that every warp has finished that plane.
	// Measured on B200 (profiles/r02_splitbar_ab.txt): bit-exact, but SLOWER than the CTA barrier (56.0 against 61.8
	// Gcell-RHS/s, each forming stage +42 us) -- the warp barrier, ticket, arrive and wait per plane and warp cost more
	// issue slots than the barrier stalls they remove.  Opt-in (PFT_SPLITBAR=1), kept as the record of that experiment.
	constexpr bool split = SPLIT;
	auto arrive_formed = [&](int q_req, int b, bool refill) {
		__syncwarp();
		if (lane == 0) {
			__threadfence_block();
			if (atomicAdd(&formed_tick, 1u) == nwarps - 1) {
				formed_tick = 0u;
				if (refill) issue(q_req, b);
			}
			mbar_arrive(&formed);
		}
	};

	uint32_t full_bits = 0;                      // parity of the next completion to wait for, per buffer (bit 3: `formed`)
	// register window: slot (q % 3) holds plane q -- the three roles (k-1, k, k+1) rotate with the unrolled
	// bodies, nothing is copied.  (u and g of plane k-1 are dead once its z-face flux is kept in fz.)
	double Pw[3][R], Uw[3][R], Gw[3][R];
	double fz[R];                                // flux through the face between planes k-1 and k
	double emax = 0.0;
	int nonfinite = 0;

	// shared-memory slot of each of the thread's cells; a row beyond the domain (ragged tile) re-reads the last
	// real row: that duplicate IS the zero-Neumann mirror of its lower neighbour, and it stores nothing
	int csr[R];
	#pragma unroll
	for (int r = 0; r < R; r++) csr[r] = cs + (min(jb + r, g.n2 - 1) - jb) * BX;

	// running output pointers (plane k0 of the output pair) and the peer pushes of the boundary planes
	double *opp[R], *oup[R];
	long long colr[R];
	#pragma unroll
	for (int r = 0; r < R; r++) {
		colr[r] = (long long)min(jb + r, g.n2 - 1) * g.px + min(i, g.n1 - 1);
		opp[r] = w.op + (long long)(k0 + 1) * g.plane + colr[r];
		oup[r] = w.ou + (long long)(k0 + 1) * g.plane + colr[r];
	}
	const bool push_lo = chunk_lo && w.lo_p != nullptr, push_hi = chunk_hi && w.hi_p != nullptr;
	const bool store_ok = (i < g.n1);

	// ---- prologue: planes k0-1 (slot 0, registers only) and k0 (slot 1, registers + tile 1)
	{
		double um[R], gm[R];
		if (qlo == 0) {
			mbar_wait(&full[0], 0); full_bits ^= 1u;
			#pragma unroll
			for (int r = 0; r < R; r++) {
				Pw[0][r] = form_state<STAGE>(raw, AS, csr[r], 0, ha);
				um[r] = form_state<STAGE>(raw, AS, csr[r], IU, ha);
				gm[r] = raw[IG * AS + csr[r]];
			}
			if (FREE) leave(0, 0);
		}
		mbar_wait(&full[1], 0); full_bits ^= 2u;
		const double *rb = raw + RS;
		double *t1 = tile + 2 * AS;
		#pragma unroll
		for (int r = 0; r < R; r++) {
			Pw[1][r] = form_state<STAGE>(rb, AS, csr[r], 0, ha);
			Uw[1][r] = form_state<STAGE>(rb, AS, csr[r], IU, ha);
			Gw[1][r] = rb[IG * AS + csr[r]];
			if (FORM) { t1[csr[r]] = Pw[1][r]; t1[AS + csr[r]] = Uw[1][r]; }
			if (qlo == 1) { um[r] = Uw[1][r]; Pw[0][r] = Pw[1][r]; gm[r] = Gw[1][r]; }      // equation.c:164-174
			fz[r] = face_flux(c, Pw[0][r], Pw[1][r], gm[r], Gw[1][r], um[r], Uw[1][r]);
		}
		if (FORM && hduty) { t1[hs] = form_state<STAGE>(rb, AS, hs, 0, ha); t1[AS + hs] = form_state<STAGE>(rb, AS, hs, IU, ha); }
	}
	if (split) arrive_formed(0, 0, false);
	else if (!FREE) __syncthreads();

	// one plane; J = it % 3 fixes every buffer, tile and window index at compile time
	auto plane = [&](auto Jc, int it) {
		constexpr int J = decltype(Jc)::value;
		constexpr int bm = J, bq = (J + 1) % 3, bn = (J + 2) % 3;   // window slots / raw buffers of planes q-1, q (centre), q+1
		constexpr int bf = J;                                        // the raw buffer to refill (plane q-1 has left it)
		const int k = k0 + it, q = it + 1;
		const bool more = (k + 1 < k1);
		const double *rb = raw + bq * RS;                            // raw planes of the centre plane k
		const double *rg = rb + IG * AS;
		double (&pm)[R] = Pw[bm], (&pc)[R] = Pw[bq], (&pn)[R] = Pw[bn];
		double (&uc)[R] = Uw[bq], (&un)[R] = Uw[bn];
		double (&gc)[R] = Gw[bq], (&gn)[R] = Gw[bn];

		// ---- request plane q+2 into the buffer plane q-1 just vacated
		if (!NOBAR && tid == 0 && next_q <= qhi && next_q <= q + 2) { issue(next_q, bf); next_q++; }
		// ---- split barrier: the state tile of plane k is complete, and nobody reads the tile of plane k-2 any more
		if (split) { mbar_wait(&formed, (full_bits >> 3) & 1u); full_bits ^= 8u; }

		// ---- plane k+1: form the stage input (centres -> registers, centres + ring -> its state tile)
		if (q + 1 <= qhi) {
			mbar_wait(&full[bn], (full_bits >> bn) & 1u);
			full_bits ^= (1u << bn);
			const double *rn = raw + bn * RS;
			double *tw = tile + bn * 2 * AS;
			#pragma unroll
			for (int r = 0; r < R; r++) {
				pn[r] = form_state<STAGE>(rn, AS, csr[r], 0, ha);
				un[r] = form_state<STAGE>(rn, AS, csr[r], IU, ha);
				gn[r] = rn[IG * AS + csr[r]];
				if (FORM && more) { tw[csr[r]] = pn[r]; tw[AS + csr[r]] = un[r]; }
			}
			if (FORM && more && hduty) { tw[hs] = form_state<STAGE>(rn, AS, hs, 0, ha); tw[AS + hs] = form_state<STAGE>(rn, AS, hs, IU, ha); }
		} else {
			#pragma unroll
			for (int r = 0; r < R; r++) { un[r] = sc.dir_u; pn[r] = pc[r]; gn[r] = gc[r]; }     // equation.c:175-183 / :251-261
		}

		if (split) arrive_formed(it + 3, bf, it >= qlo && it + 3 <= qhi);

		// ---- plane k: in-plane neighbours of the column of R cells
		const double *tp = FORM ? tile + bq * 2 * AS : rb;
		const double *tu = FORM ? tp + AS : rb + IU * AS;
		// below the lowest cell / above the highest one (other threads' rows, or the mirror)
		const double pS = tp[sS], uS = tu[sS], gS = rg[sS];
		const double pN = tp[sN], uN = tu[sN], gN = rg[sN];
		// y-faces: fy[r] lies below cell r, fy[R] above the highest cell; the faces between the thread's own cells
		// are evaluated once (same operands from either side, equation.c:712-719)
		double fy[R + 1];
		fy[0] = face_flux(c, pS, pc[0], gS, gc[0], uS, uc[0]);
		#pragma unroll
		for (int r = 1; r < R; r++) fy[r] = face_flux(c, pc[r - 1], pc[r], gc[r - 1], gc[r], uc[r - 1], uc[r]);
		fy[R] = face_flux(c, pc[R - 1], pN, gc[R - 1], gN, uc[R - 1], uN);

		const bool plane_lo = push_lo && it == 0, plane_hi = push_hi && k == g.n3 - 1;
		#pragma unroll
		for (int r = 0; r < R; r++) {
			const int s0 = csr[r];
			const double pw = tp[s0 + dW], pe = tp[s0 + dE];
			const double uw = tu[s0 + dW], ue = tu[s0 + dE];
			const double gw = rg[s0 + dW], ge = rg[s0 + dE];
			const double p_dn = (r == 0) ? pS : pc[r - 1];
			const double p_up = (r + 1 < R) ? pc[r + 1] : pN;
			const double fw = face_flux(c, pw, pc[r], gw, gc[r], uw, uc[r]);
			const double fe = face_flux(c, pc[r], pe, gc[r], ge, uc[r], ue);
			const double fzp = face_flux(c, pc[r], pn[r], gc[r], gn[r], uc[r], un[r]);
			const double sum = c.h1_2 * (-fw + fe) + c.h2_2 * (-fy[r] + fy[r + 1]) + c.h3_2 * (-fz[r] + fzp);   // equation.c:711-720
			fz[r] = fzp;
			double du, dp;
			cell_rhs<MODEL>(c, uc[r], pc[r], gc[r], pw, pe, p_dn, p_up, pm[r], pn[r], sum, du, dp);
			// ---- fused combination that follows this stage (operands re-read from the raw buffer)
			if (store_ok && jb + r < g.n2) {
				double vu, vp;
				stage_values<STAGE>(sc.hb, rb, AS, IU, s0, du, dp, emax, nonfinite, vu, vp);
				*opp[r] = vp; *oup[r] = vu;
				if (plane_lo) { w.lo_p[colr[r]] = vp; w.lo_u[colr[r]] = vu; }     // the neighbours' ghost planes (peer path)
				if (plane_hi) { w.hi_p[colr[r]] = vp; w.hi_u[colr[r]] = vu; }
			}
			opp[r] += g.plane; oup[r] += g.plane;
		}

		// ---- this plane is done: raw[bq] may be refilled, its state tile overwritten (three planes later)
		if (FREE) leave(q, bq);
		else if (!split) __syncthreads();
	};

	const int nplanes = k1 - k0;
	int it = 0;
	#pragma unroll 1
	for (; it + 3 <= nplanes; it += 3) {
		plane(std::integral_constant<int, 0>(), it);
		plane(std::integral_constant<int, 1>(), it + 1);
		plane(std::integral_constant<int, 2>(), it + 2);
	}
	if (it < nplanes) { plane(std::integral_constant<int, 0>(), it); it++; }
	if (it < nplanes) { plane(std::integral_constant<int, 1>(), it); it++; }

	halo_signal(a.pr, sc.hseq, chunk_lo, chunk_hi, tid);
	if (signals) dep_signal(a, my_chunk, tid);
	if (STAGE == 5) stage5_finish(a, tid, emax, nonfinite, smem_t, P.smem_bytes);
}

}  // namespace pft
