// pft_ctl.h -- the step-size controller of the Runge-Kutta-Merson integrator as ONE piece of code
// that runs on the device (single thread: the last CTA of stage 5, or a one-thread kernel) and, for
// the unit tests and the CPU emulation, on the host.
//
// It restates the scalar part of modules/RK_MPI_SAsolver_hybrid2/RK_MPI_SAsolver_hybrid2.c:351-766
// (error test :578-611, NaN branch :626-646, update :651-705, NEXTFINISH :743-761) on a control
// block that lives in device memory.  Because every scalar an attempt needs (h/3, h/6, h/8, the
// Dirichlet temperature at the five stage times, which array pair holds x, whether K1 is still
// valid) is READ FROM THAT BLOCK by the stage kernels, the launches of an attempt do not depend on
// the outcome of the previous one: the host enqueues attempts ahead of the device and only watches
// a progress slot (pft_b200.cu).  No host round trip, no host spin, in the time loop.
//
// Speculative stage 1 (StepCtl.spec, pft_b200.cu): K1 of the NEXT attempt, f(candidate state), does not depend
// on the decision either -- only on stage 5's output -- so the launches of an attempt become
// [S2 S3 S4 S5 S1'] with S1' writing the K1 buffer that goes with the candidate pair.  The decision then merely
// selects which (x, K1) pair is live (StepCtl.cur), and the grid-wide wait for it moves from "before stage 1"
// to "before stage 2", where it hides behind S1'.  Two control blocks alternate by attempt parity: the decision
// of attempt L reads block L&1 and WRITES block (L+1)&1, so S1' of attempt L still reads a stable block.
//
// All arithmetic is IEEE double in the reference's order; x^0.2 comes from pft_pow.h.
#pragma once
#include "pft_pow.h"
#include <stddef.h>

namespace pft {

// RK_MPI_SAsolver.h:30-35
enum { CMD_h_TOO_SMALL = 1, CMD_NAN = 2, CMD_UPDATE = 4, CMD_FINISHED = 8, CMD_NEXTFINISH = 16, CMD_BREAK = 32 };

// state of the integration on the device
enum { ST_RUNNING = 0, ST_DONE = 1, ST_BUDGET = 2, ST_NAN = 3 };

struct StepCtl {
	// ---- integration state (RK_MPI_S_SOLUTION's scalars + the solver's locals)
	double t, h;                     // time level of x; step length of the attempt in flight
	double final_time, h_min, delta;
	double new_h;                    // last proposal of the controller (:580)
	double sys_h;                    // what system->h has to read at return (:754, budget exit) ...
	int sys_h_set;                   // ... if it was written at all during this solve call
	int delta_local, handle_nan, recompute_k1;
	int command;                     // CMD_FINISHED carried into the attempt in flight (:324, :758)
	int status;                      // ST_*
	int last_nan;
	long long steps, steps_total, attempts, max_attempts, rhs_evals;
	// Dirichlet switch (equation.c:96-111)
	double switch_time, top_temp1, top_temp2;
	// ---- what the stage kernels of the attempt in flight read
	int cur;                         // 0: pair A_X holds x and A_XN receives the candidate; 1: the other way round
	int skip1;                       // K1 == f(t, x) already holds (rejected attempt, :373 recomputes it needlessly)
	double h3, h6, h8;               // :355
	double dir[5];                   // Dirichlet temperature at t, t+h/3, t+h/3, t+h/2, t+h
	unsigned long long aseq;         // attempts started on this solver so far: sequence base of halo flags and eps slots
	// ---- error reduction of stage 5 (grid-wide max through atomics, last-CTA ticket)
	unsigned long long eps_bits;
	unsigned int ticket;
	int nonfinite;
	double red[2];                   // {eps, nonfinite} of this slab when the decision is taken elsewhere (NCCL / host pow)
	// ---- speculative stage 1 (two alternating blocks; see the header comment)
	unsigned long long hbase;        // halo-flag sequence base: S1 (at solve entry, if K1 is not valid) +0, S2..S5 +1..+4, S1' +5
	int spec, pad_;
	// (not copied from block to block: the words the stage kernels wait on)
	unsigned long long ready;        // host attempt number this block is valid for
	unsigned long long closed_from;  // attempts numbered >= this lie beyond the end of the integration (~0: open)
};
// 8-byte words of a block that the decision carries over to the next block
#define PFT_CTL_COPY_WORDS (offsetof(pft::StepCtl, ready) / 8)

// progress slot in pinned, mapped host memory; written by the controller after every attempt
struct HostSlot {
	unsigned long long seq;          // == StepCtl.aseq after the attempt (written last)
	unsigned long long seq_eps;      // host-pow mode: eps of attempt seq_eps is ready, the device waits for pow(delta/eps, 0.2)
	double eps;                      // global error maximum of the attempt
	int accept, nonfinite;
	int status, command;             // ST_*; CMD_* bits of the attempt (UPDATE, NEXTFINISH, NAN ...)
	double t, h_attempt, new_h;      // time level after the attempt, the step it used, the proposal
	double h_next; int command_next; // step and carried command (CMD_FINISHED) of the attempt that follows
	double sys_h; int sys_h_set, cur;
	long long steps, steps_total, attempts, rhs_evals;
	int last_nan, delta_local;
	double h3_attempt;               // |h/3| of the attempt (DELTA_LOCAL scaling of eps, :578)
	unsigned long long hbase;        // StepCtl.hbase after the attempt
};

PFT_HD double ctl_dirichlet(const StepCtl &c, double t)
{
	return t < c.switch_time ? c.top_temp1 : c.top_temp2;          // equation.c:96-111
}

// derived scalars of the attempt that is about to run
PFT_HD void ctl_prepare(StepCtl &c)
{
	const double h = c.h;
	const double h2 = h / 2.0;
	c.h3 = h / 3.0; c.h6 = h / 6.0; c.h8 = h / 8.0;                // :355
	c.dir[0] = ctl_dirichlet(c, c.t);
	c.dir[1] = c.dir[2] = ctl_dirichlet(c, c.t + c.h3);            // :392,412
	c.dir[3] = ctl_dirichlet(c, c.t + h2);                         // :432
	c.dir[4] = ctl_dirichlet(c, c.t + h);                          // :453
	if (!c.skip1) c.rhs_evals++;                                   // stage 1 will run
}

// solve entry (:319-325): direction, truncation of the first step
PFT_HD void ctl_begin(StepCtl &c)
{
	c.command = 0; c.status = ST_RUNNING; c.last_nan = 0; c.attempts = 0; c.sys_h_set = 0;
	if (!c.skip1) c.hbase += 1;      // the stage-1 launch at solve entry gets a halo sequence number of its own
	if ((c.final_time > c.t && c.h < 0) || (c.final_time < c.t && c.h > 0)) c.h *= -1;
	const double d = c.final_time - c.t;
	if (c.h == 0 || (d < 0 ? -d : d) <= (c.h < 0 ? -c.h : c.h)) { c.h = d; c.command |= CMD_FINISHED; }
	ctl_prepare(c);
}

PFT_HD double ctl_abs(double x) { return x < 0 ? -x : (x == 0 ? 0.0 : x); }

// the decision after stage 5.  eps = global error maximum, pw = pow(delta/max_eps, 0.2) or, when
// tab != nullptr, computed here.  Returns the CMD_* bits of this attempt (for the progress slot).
PFT_HD int ctl_decide(StepCtl &c, double eps, int nonfinite, const PowTables *tab, double pw_host)
{
	const double h = c.h, t = c.t;
	c.steps_total++; c.attempts++;                                 // :460
	c.rhs_evals += 4;
	double max_eps = eps;
	if (c.delta_local) max_eps *= ctl_abs(c.h3);                   // :578
	double new_h;
	if (max_eps > 0.0) {
		const double pw = tab ? pft_pow_twin(*tab, c.delta / max_eps, 0.2) : pw_host;
		new_h = (pw * 0.8) * h;                                    // :580
	} else new_h = 2.0 * h;
	c.new_h = new_h;

	int command = c.command;
	if (c.handle_nan && nonfinite) {                               // :493-503
		command |= CMD_NAN;
		if (h / (c.final_time - t) < 1e-11) command |= CMD_h_TOO_SMALL;
	}
	if (max_eps < c.delta || ctl_abs(h) < c.h_min) {               // :599-611
		command |= CMD_UPDATE;
		if (ctl_abs(c.final_time - (t + h)) <= ctl_abs(new_h)) command |= CMD_NEXTFINISH;
	}
	const int result = command;

	if (command & CMD_NAN) {                                       // :626-646
		c.last_nan = 1;
		c.skip1 = c.recompute_k1 ? 0 : 1;
		if (command & CMD_h_TOO_SMALL) { c.status = ST_NAN; c.skip1 = 1; c.aseq++; c.hbase += 5; return result; }
		c.h = h / 10;
		c.command = 0;
	} else {
		if (command & CMD_UPDATE) {                                // :651-734
			c.t = t + h;
			c.cur ^= 1;
			c.steps++;
			c.skip1 = 0;
			if (command & CMD_FINISHED) c.status = ST_DONE;
		} else c.skip1 = c.recompute_k1 ? 0 : 1;
		if (c.status == ST_RUNNING) {
			if (command & CMD_NEXTFINISH) {                        // :745-760
				c.sys_h = new_h; c.sys_h_set = 1;
				c.h = c.final_time - c.t;
				c.command = CMD_FINISHED;
			} else {
				c.command = 0;
				c.h = new_h;
			}
		}
	}
	if (c.status == ST_RUNNING && c.max_attempts > 0 && c.attempts >= c.max_attempts) {
		if (!(c.command & CMD_FINISHED)) { c.sys_h = c.h; c.sys_h_set = 1; }
		c.status = ST_BUDGET;
	}
	c.aseq++;
	c.hbase += 5;
	if (c.status == ST_RUNNING) ctl_prepare(c);
	return result;
}

PFT_HD void ctl_publish(const StepCtl &c, HostSlot *s, double eps, int nonfinite, int command, double h_attempt)
{
	s->eps = eps;
	s->accept = ((command & CMD_UPDATE) && !(command & CMD_NAN)) ? 1 : 0;
	s->nonfinite = nonfinite;
	s->status = c.status; s->command = command;
	s->t = c.t; s->h_attempt = h_attempt; s->new_h = c.new_h;
	s->h_next = c.h; s->command_next = c.command;
	s->sys_h = c.sys_h; s->sys_h_set = c.sys_h_set; s->cur = c.cur;
	s->steps = c.steps; s->steps_total = c.steps_total; s->attempts = c.attempts; s->rhs_evals = c.rhs_evals;
	s->last_nan = c.last_nan;
	s->hbase = c.hbase;
}

}  // namespace pft
