// pft_pow.h -- x^y for the step-size controller, bit-identical to the HOST libm's pow().
//
// The reference computes the next step as pow(delta/max_eps, 0.2)*0.8*h with the C library's pow()
// (modules/RK_MPI_SAsolver_hybrid2/RK_MPI_SAsolver_hybrid2.c:580).  The accepted-step sequence is
// chaotic with respect to one-ulp differences of that value, CUDA's pow() is a different algorithm,
// and glibc's is not correctly rounded either (~0.52 ulp), so "a good pow" is not enough: the
// device has to repeat the host library's own operation sequence on the host library's own tables.
//
//  * The operation sequence below is glibc's table-driven pow (log with a 128-entry table and a
//    degree-7 polynomial in double-double, exp with a 128-entry table), in the exact order and with
//    the exact fused multiply-adds of the x86-64 FMA build of the library (the one the dynamic
//    loader selects on every FMA-capable host); fma() has the same meaning on the host and on sm_100a.
//  * The tables and polynomial coefficients are NOT in this repository: pft_pow_load_tables() reads
//    them out of the libm image the process has loaded (located through dladdr(pow)), finding the
//    two constant blocks by their leading values (ln2 split / 128/ln2 and the 1.5*2^52 shift).
//  * pft_pow_selftest() then compares the host instance of pft_pow_twin() with the library's pow()
//    on the operand range the controller can produce.  Only if every value agrees bit for bit is the
//    device controller enabled; otherwise the integrator asks the host for pow() once per attempt
//    (same results, one PCIe round trip per attempt).  Nothing is approximated in either mode.
//
// Covered domain: y finite with 2^-65 <= |y| < 2^63 (the controller uses 0.2), x >= +0 incl. subnormals and +inf.
#pragma once
#include <stdint.h>
#include <string.h>

#ifndef PFT_HD
#if defined(__CUDACC__) && !defined(PFT_CUDA_EMU)
#define PFT_HD __host__ __device__ __forceinline__
#else
#define PFT_HD inline
#endif
#endif

namespace pft {

struct PowTables {
	double ln2hi, ln2lo;
	double A[7];                 // log polynomial
	double log_tab[128][4];      // invc, (pad), logc, logctail
	double inv_ln2N, shift, neg_ln2hiN, neg_ln2loN;
	double C[4];                 // exp polynomial C2..C5
	uint64_t exp_tab[256];       // tail bits, scale bits
};

PFT_HD uint64_t pow_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
PFT_HD double pow_dbl(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

// fused multiply-add with one rounding: fma() on both sides (the host build needs -mfma or libm's fma)
// (the non-fused operations go through round-to-nearest intrinsics on the device so that no build
// flag -- the opt-in FMA variant of the library uses -fmad=true -- can contract them; the host side
// is compiled with -ffp-contract=off)
#if defined(__CUDA_ARCH__)
#define PFT_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PFT_MUL(a, b) __dmul_rn((a), (b))
#define PFT_ADD(a, b) __dadd_rn((a), (b))
#define PFT_SUB(a, b) __dsub_rn((a), (b))
#else
#define PFT_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define PFT_MUL(a, b) ((a) * (b))
#define PFT_ADD(a, b) ((a) + (b))
#define PFT_SUB(a, b) ((a) - (b))
#endif

PFT_HD double pft_pow_twin(const PowTables &T, double x, double y)
{
	uint64_t ix = pow_bits(x);
	if (ix == 0) return 0.0;                                   // pow(+0, y > 0)
	if (ix == 0x7ff0000000000000ull) return x;                 // pow(+inf, y > 0)
	if ((ix >> 52) == 0) {                                     // subnormal x: normalise, exponent becomes negative
		ix = pow_bits(PFT_MUL(x, 0x1p52));
		ix &= 0x7fffffffffffffffull;
		ix -= 52ull << 52;
	}
	// ---- log(x) = hi + lo
	const uint64_t tmp = ix - 0x3fe6955500000000ull;
	const int i = (int)((tmp >> 45) & 127);
	const int k = (int)((int64_t)tmp >> 52);
	const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
	const double z = pow_dbl(iz);
	const double kd = (double)k;
	const double invc = T.log_tab[i][0], logc = T.log_tab[i][2], logctail = T.log_tab[i][3];
	const double t1 = PFT_FMA(kd, T.ln2hi, logc);
	const double lo1 = PFT_FMA(kd, T.ln2lo, logctail);
	const double r = PFT_FMA(z, invc, -1.0);
	const double ar = PFT_MUL(r, T.A[0]);
	const double q12 = PFT_FMA(r, T.A[2], T.A[1]);
	const double q34 = PFT_FMA(r, T.A[4], T.A[3]);
	const double t2 = PFT_ADD(r, t1);
	const double lo2 = PFT_ADD(PFT_SUB(t1, t2), r);
	const double ar2 = PFT_MUL(r, ar);
	const double ar3 = PFT_MUL(r, ar2);
	const double lo3 = PFT_FMA(ar, r, -ar2);
	const double hi = PFT_ADD(t2, ar2);
	const double q56 = PFT_FMA(r, T.A[6], T.A[5]);
	const double lo4 = PFT_ADD(PFT_SUB(t2, hi), ar2);
	const double q = PFT_FMA(ar2, PFT_FMA(q56, ar2, q34), q12);
	const double lo = PFT_FMA(ar3, q, PFT_ADD(PFT_ADD(PFT_ADD(lo1, lo2), lo3), lo4));
	const double lh = PFT_ADD(hi, lo);
	const double ll = PFT_ADD(PFT_SUB(hi, lh), lo);
	// ---- y*log(x) = ehi + elo
	const double ehi = PFT_MUL(y, lh);
	const double elo = PFT_FMA(y, ll, PFT_FMA(lh, y, -ehi));
	// ---- exp(ehi + elo)
	const uint32_t abstop = (uint32_t)(pow_bits(ehi) >> 52) & 0x7ff;
	if (abstop - 0x3c9u > 0x3eu) {
		if ((int32_t)(abstop - 0x3c9u) < 0) return PFT_ADD(1.0, ehi);  // |y log x| < 2^-54
		return (ehi > 0) ? pow_dbl(0x7ff0000000000000ull) : 0.0;   // not reachable for x^0.2; overflow / underflow otherwise
	}
	const double kz = PFT_FMA(ehi, T.inv_ln2N, T.shift);
	const uint64_t ki = pow_bits(kz);
	const double kk = PFT_SUB(kz, T.shift);
	double rr = PFT_FMA(kk, T.neg_ln2loN, PFT_FMA(kk, T.neg_ln2hiN, ehi));
	const unsigned idx = 2u * (unsigned)(ki & 127);
	const uint64_t sbits = T.exp_tab[idx + 1] + (ki << 45);
	rr = PFT_ADD(elo, rr);
	const double c23 = PFT_FMA(rr, T.C[1], T.C[0]);
	const double tr = PFT_ADD(rr, pow_dbl(T.exp_tab[idx]));
	const double r2 = PFT_MUL(rr, rr);
	const double c45 = PFT_FMA(rr, T.C[3], T.C[2]);
	const double p2 = PFT_FMA(c23, r2, tr);
	const double r4 = PFT_MUL(r2, r2);
	const double tmpv = PFT_FMA(c45, r4, p2);
	const double scale = pow_dbl(sbits);
	return PFT_FMA(tmpv, scale, scale);
}

}  // namespace pft

// ---- host only: tables out of the loaded libm, and the gate that decides whether they may be used
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace pft {

// returns 0 and fills T, or a negative code with a message in err
inline int pft_pow_load_tables(PowTables *T, char *err, size_t errlen)
{
	Dl_info info;
	double (*fn)(double, double) = pow;
	if (!dladdr((void *)fn, &info) || !info.dli_fname) { snprintf(err, errlen, "dladdr(pow) failed"); return -1; }
	FILE *f = fopen(info.dli_fname, "rb");
	if (!f) { snprintf(err, errlen, "cannot open %s", info.dli_fname); return -1; }
	fseek(f, 0, SEEK_END);
	const long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	unsigned char *b = (n > 0) ? (unsigned char *)malloc((size_t)n) : nullptr;
	if (!b || fread(b, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(b); snprintf(err, errlen, "cannot read %s", info.dli_fname); return -1; }
	fclose(f);
	auto dbl_at = [&](long off) { double v; memcpy(&v, b + off, 8); return v; };
	auto u64_at = [&](long off) { uint64_t v; memcpy(&v, b + off, 8); return v; };
	const double ln2hi = 0x1.62e42fefa3800p-1, ln2lo = 0x1.ef35793c76730p-45;
	bool have_log = false, have_exp = false;
	for (long o = 0; o + 8 * (9 + 512) <= n && !have_log; o += 8) {
		// {ln2hi, ln2lo, A[0] = -0.5, A[1..6], tab[128] x {invc, pad, logc, logctail}}
		if (dbl_at(o) != ln2hi || dbl_at(o + 8) != ln2lo || dbl_at(o + 16) != -0.5) continue;
		bool ok = true;
		for (int i = 0; i < 128 && ok; i++) {
			const double invc = dbl_at(o + 72 + 32 * i), pad = dbl_at(o + 72 + 32 * i + 8), logc = dbl_at(o + 72 + 32 * i + 16);
			ok = invc > 0.5 && invc < 1.5 && pad == 0.0 && fabs(logc + log(invc)) < 1e-9;
		}
		if (!ok) continue;
		T->ln2hi = ln2hi; T->ln2lo = ln2lo;
		for (int i = 0; i < 7; i++) T->A[i] = dbl_at(o + 16 + 8 * i);
		for (int i = 0; i < 128; i++) for (int c = 0; c < 4; c++) T->log_tab[i][c] = dbl_at(o + 72 + 32 * i + 8 * c);
		have_log = true;
	}
	const double inv_ln2N = 0x1.71547652b82fep0 * 128.0;
	for (long o = 0; o + 8 * (64 + 256) <= n && !have_exp; o += 8) {
		// {InvLn2N, Shift, NegLn2hiN, NegLn2loN, C2..C5, (exp2 / exp10 constants, count differs between
		//  library versions), tab[256]}; the table starts with the pair {0, bits(1.0)}
		if (dbl_at(o) != inv_ln2N || dbl_at(o + 8) != 0x1.8p52) continue;
		long t = -1;
		for (int q = 8; q < 64; q++)
			if (u64_at(o + 8 * q) == 0 && u64_at(o + 8 * q + 8) == 0x3ff0000000000000ull && u64_at(o + 8 * q + 24) > 0x3fe0000000000000ull &&
			    u64_at(o + 8 * q + 24) < 0x3ff0000000000000ull) { t = o + 8 * q; break; }
		if (t < 0) continue;
		T->inv_ln2N = inv_ln2N; T->shift = 0x1.8p52;
		T->neg_ln2hiN = dbl_at(o + 16); T->neg_ln2loN = dbl_at(o + 24);
		for (int i = 0; i < 4; i++) T->C[i] = dbl_at(o + 32 + 8 * i);
		for (int i = 0; i < 256; i++) T->exp_tab[i] = u64_at(t + 8 * i);
		// the scale entries must be 2^(i/128) with the exponent contribution i<<45 taken out
		bool ok = true;
		for (int i = 0; i < 128 && ok; i++) {
			const double s = pow_dbl(T->exp_tab[2 * i + 1] + ((uint64_t)i << 45));
			ok = fabs(s - exp2(i / 128.0)) < 1e-15 * 4;
		}
		have_exp = ok;
	}
	free(b);
	if (!have_log || !have_exp) {
		snprintf(err, errlen, "pow tables not found in %s (log %d, exp %d)", info.dli_fname, (int)have_log, (int)have_exp);
		return -2;
	}
	return 0;
}

// compares the twin with the library on `n` operands x^y: x log-uniform over the whole positive
// range (incl. subnormals), plus operands packed around 1 and around the controller's usual
// delta/eps range; returns the number of mismatches
inline long pft_pow_selftest(const PowTables &T, double y, long n, uint64_t seed)
{
	long bad = 0;
	uint64_t s = seed ? seed : 0x9e3779b97f4a7c15ull;
	auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
	for (long i = 0; i < n; i++) {
		const uint64_t r = next();
		double x;
		switch (i & 3) {
		case 0: x = pow_dbl(r % 0x7ff0000000000000ull); break;                                   // any finite positive bit pattern
		case 1: x = pow_dbl(0x3ff0000000000000ull + (int64_t)(r % 0x0010000000000000ull) - 0x0008000000000000ll); break;   // [0.75, 1.5)
		case 2: x = pow_dbl(0x3f50000000000000ull + r % 0x0160000000000000ull); break;             // ~[1e-3, 4e3): delta/eps of working steps
		default: x = pow_dbl(0x3ff0000000000000ull + (int64_t)(r % 4096) - 2048); break;             // the 4096 doubles around 1.0
		}
		const double a = pft_pow_twin(T, x, y), b = pow(x, y);
		if (pow_bits(a) != pow_bits(b)) bad++;
	}
	return bad;
}

}  // namespace pft
