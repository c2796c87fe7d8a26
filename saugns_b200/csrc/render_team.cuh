/* render_team.cuh -- part of kernels.cu (one translation unit; included inside namespace saugen):
 * parallelism ALONG TIME inside one voice.
 *
 * With fewer voices than resident warps, one warp per voice leaves the machine idle and a
 * launch takes as long as one voice's serial render.  In a steady stretch nothing but the
 * oscillator accumulators and the differentiator's one-sample look-back carries over from sample
 * to sample (wosc.h:129,135-169,247-262; the line trajectories are closed forms in the position,
 * line.c:27-37).  A TEAM of T warps of one CTA therefore splits a stretch of C chunks into T
 * ranges [a_w, a_w+1), w = 0..T-1; member w > 0 needs, at its start,
 *   - every operator's phase accumulator: phase0 + n * inc in closed form where the frequency is
 *     uniform (acc level 0); for a frequency-modulated operator the sum of its rounded increments
 *     over everything before -- found by COUNTING passes, level by level of the FM nesting: pass p
 *     renders only what the level-p operators' frequencies depend on (everything of output level
 *     < p), accumulates the level-p increments per range, and a prefix over the team's members
 *     gives every member its start value (integer sums: bit-identical to the serial accumulation);
 *   - prev_phase / prev_Is / prev_s: each member starts L = P + 1 chunks early (P = deepest acc
 *     level) and renders those lead-in chunks without output; an operator of level k gets its exact
 *     accumulator at lead-in chunk k, its inputs (levels < k) are exact from the chunk before, so
 *     its own output is exact from chunk k + 1 at the latest (a differentiated sample needs the
 *     phases of two consecutive samples; the nesting depth is far below a chunk's 128 samples).
 * The leader (member 0) is the voice's own warp: it applies events, renders everything that is not
 * a steady stretch, builds and lowers the plan, and hands eligible stretches to the team; the last
 * member's operator state becomes the voice's state.  Members synchronise on one named barrier per
 * team.  Eligible: lowered plans (render_fast.cuh) made of wave operators, lines, range / mix
 * records and the voice output -- no noise / rumble / self-PM records (their state is not a closed
 * form or a prefix sum of independent terms), no operator standing still (a zero phase increment
 * without PM repeats one output forever: the look-back is unbounded).
 */
#pragma once

constexpr uint32_t TEAM_INELIGIBLE = 0xffu;
constexpr uint32_t TEAM_MAX_P = 6;
constexpr uint32_t OS_PAD0 = 184, OS_PAD1 = 188;      /* OpState::_pad: a member's count / its start value */
static_assert(offsetof(OpState, _pad) == OS_PAD0, "OpState::_pad offset");
constexpr uint32_t TC_OP = 0, TC_NREC = 4, TC_NOPS = 8, TC_CHUNKS = 12, TC_P = 16, TC_TEFF = 20,
	TC_FUSED = 24;       /* the command block */

__device__ __forceinline__ void team_bar(uint32_t id, uint32_t nthreads) {
	asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

/* lane 0: acc level A (w1 bits 24..27; 15 = no accumulator) and output level O (bits 28..31) of
 * every record of a lowered plan; returns P = the deepest acc level, or TEAM_INELIGIBLE. */
__device__ __noinline__ uint32_t team_analyse(uint32_t plan, uint32_t nrec) {
	uint8_t blev[32];
	for (int i = 0; i < 32; ++i) blev[i] = 0;
	uint32_t vlev = 0, P = 0;
	auto lev = [&](uint32_t b) -> uint32_t { return b < 32u ? blev[b] : 0u; };
	auto mx = [](uint32_t a, uint32_t b) { return a > b ? a : b; };
	for (uint32_t r = 0; r < nrec; ++r) {
		const uint32_t a = plan + r * PLAN_REC;
		const uint32_t w0 = lds32(a), w1 = lds32(a + 4);
		const uint32_t kind = w0 & 0xffu, fl = (w0 >> 8) & 0xffu, xf = (w1 >> 16) & 0xffu;
		const uint32_t bufa = (w0 >> 16) & 0xffu, bufb = w0 >> 24;
		uint32_t A = 15, O = 0;
		if (kind == P_EXT) continue;
		if (kind >= X_OSC0 && kind < X_RANGE) {
			const uint32_t v = kind - X_OSC0, fs = v / 3u, pm = v % 3u;
			if (bufa >= 32u) return TEAM_INELIGIBLE;
			if (fs == 0) {
				A = 0;
				if (pm == 0 && lds32(a + 20) == 0u) return TEAM_INELIGIBLE;     /* stands still */
			} else {
				const uint32_t src = fs == 1 ? bufb : (w1 >> 8) & 0xffu;
				if (!(xf & XF_SRC_VAL) && src >= 32u) return TEAM_INELIGIBLE;
				A = 1u + ((xf & XF_SRC_VAL) ? vlev : lev(src));
			}
			O = A;
			if (pm == 1) { if ((w1 & 0xffu) >= 32u) return TEAM_INELIGIBLE; O = mx(O, lev(w1 & 0xffu)); }
			if (pm == 2) O = mx(O, vlev);
			if (fl & PF_LAYER) O = mx(O, lev(bufa));
			if (A > TEAM_MAX_P || O > TEAM_MAX_P) return TEAM_INELIGIBLE;
			P = mx(P, A);
			blev[bufa] = (uint8_t) O; vlev = O;
		} else if (kind == X_RANGE) {
			if (bufa >= 32u) return TEAM_INELIGIBLE;
			const uint32_t m = w1 & 0xffu;
			if (!(xf & XF_SRC_VAL) && m >= 32u) return TEAM_INELIGIBLE;
			O = (xf & XF_SRC_VAL) ? vlev : lev(m);
			blev[bufa] = (uint8_t) O; vlev = O;
		} else if (kind == X_VOUT) {
			O = 15;                        /* never part of a counting pass */
		} else if (kind == P_LINE) {       /* a line value, times a multiplier buffer when it is a ratio */
			if (bufa >= 32u || (bufb != NO_BUF && bufb >= 32u)) return TEAM_INELIGIBLE;
			O = bufb != NO_BUF ? lev(bufb) : 0u;
			blev[bufa] = (uint8_t) O; vlev = 0;
		} else if (kind == P_WHEAD) {      /* a frequency line into buffer b */
			const uint32_t e = (w1 >> 8) & 0xffu;
			if (bufb >= 32u || (e != NO_BUF && e >= 32u)) return TEAM_INELIGIBLE;
			O = e != NO_BUF ? lev(e) : 0u;
			blev[bufb] = (uint8_t) O; vlev = 0;
		} else if (kind == P_RANGE) {
			const uint32_t m = w1 & 0xffu;
			if (bufa >= 32u || bufb >= 32u || m >= 32u) return TEAM_INELIGIBLE;
			O = mx(mx(lev(bufa), lev(bufb)), lev(m));
			blev[bufa] = (uint8_t) O; vlev = 0;
		} else if (kind == P_MIX) {
			const uint32_t c = w1 & 0xffu;
			if (bufa >= 32u || (bufb != NO_BUF && bufb >= 32u) || (!(fl & PF_ACONST) && c >= 32u)) return TEAM_INELIGIBLE;
			O = bufb != NO_BUF ? lev(bufb) : 0u;
			if (!(fl & PF_ACONST)) O = mx(O, lev(c));
			if (fl & PF_LAYER) O = mx(O, lev(bufa));
			blev[bufa] = (uint8_t) O; vlev = 0;
		} else {
			return TEAM_INELIGIBLE;        /* unlowered wave operators, noise, rumble, self-PM */
		}
		sts32(a + 4, (w1 & 0x00ffffffu) | A << 24 | O << 28);
	}
	return P;
}

/* Which records carry an operator's shared address in w2 (relocated for a member's copy). */
__device__ __forceinline__ bool rec_has_op(uint32_t kind) {
	return (kind >= X_OSC0 && kind < X_RANGE) || kind == X_VOUT || kind == X_COUNT1 || kind == X_COUNT2 ||
		kind == P_LINE || kind == P_WHEAD;
}

/* lane 0: the member's executable plan for counting pass `pass` (0 = the full plan): header and
 * records copied from the master, operator addresses moved by `delta`.  Returns the shared
 * address of the voice-output record in the copy (0 when the pass has none). */
__device__ __noinline__ uint32_t team_build_plan(uint32_t master, uint32_t exec, uint32_t nrec, uint32_t delta,
		uint32_t pass) {
	for (uint32_t i = 0; i < PLAN_HDR; i += 16) {
		const uint4 h = lds128u(master + i);
		asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(exec + i), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w) : "memory");
	}
	uint32_t out = exec + PLAN_HDR, vout = 0;
	for (uint32_t r = 0; r < nrec; ++r) {
		const uint32_t a = master + PLAN_HDR + r * PLAN_REC;
		uint4 x = lds128u(a);
		const uint4 y = lds128u(a + 16);
		uint32_t kind = x.x & 0xffu;
		const uint32_t fl = (x.x >> 8) & 0xffu;
		const bool ext = (kind == P_WLEAF || kind == P_WTAIL || (kind >= X_OSC0 && kind < X_RANGE)) && (fl & PF_AEXT);
		const uint32_t A = (x.y >> 24) & 0xfu, O = x.y >> 28;
		bool keep = true, second = ext;
		if (pass) {
			if (O < pass) keep = true;
			else if (kind >= X_OSC0 && kind < X_RANGE && A == pass) {
				const uint32_t fs = (kind - X_OSC0) / 3u;            /* fs >= 1: A >= 1 */
				kind = fs == 1 ? X_COUNT1 : X_COUNT2;
				x.x = (x.x & ~0xffffu) | kind;                        /* no flags: no second slot to skip */
				second = false;
			} else {
				keep = false;
			}
		}
		if (keep) {
			if (rec_has_op(kind)) x.z += delta;
			if (kind == X_VOUT) vout = out;
			asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(out), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
			asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(out + 16), "r"(y.x), "r"(y.y), "r"(y.z), "r"(y.w) : "memory");
			out += PLAN_REC;
			if (second) {
				const uint4 e0 = lds128u(a + PLAN_REC), e1 = lds128u(a + PLAN_REC + 16);
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(out), "r"(e0.x), "r"(e0.y), "r"(e0.z), "r"(e0.w) : "memory");
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(out + 16), "r"(e1.x), "r"(e1.y), "r"(e1.z), "r"(e1.w) : "memory");
				out += PLAN_REC;
			}
		}
		if (ext) ++r;
	}
	sts32(out, P_STOP);
	return vout;
}

/* lane 0: give the accumulators of the exec plan's level-`k` operators their start values:
 * k == 0: the closed form at chunk `at` (uniform increment in w6); k > 0: the member's prefix
 * (OS_PAD1), or zero where a counting pass starts counting that level. */
__device__ __noinline__ void team_patch(uint32_t exec, uint32_t k, uint32_t at, bool zero) {
	for (uint32_t a = exec + PLAN_HDR; ; a += PLAN_REC) {
		const uint4 x = lds128u(a);
		const uint32_t kind = x.x & 0xffu;
		if (kind == P_STOP || a - exec > PLAN_WALK_MAX) break;
		const bool osc = kind >= X_OSC0 && kind < X_RANGE, cnt = kind == X_COUNT1 || kind == X_COUNT2;
		if (!(osc || cnt) || ((x.y >> 24) & 0xfu) != k) continue;
		if (k == 0) sts32(x.z + OS_I0, lds32(x.z + OS_I0) + lds32(a + 20) * (at * (uint32_t) CHUNK));
		else sts32(x.z + OS_I0, zero ? 0u : lds32(x.z + OS_PAD1));
	}
}

struct TeamCtx {
	uint32_t T, rank, bar;     /* members, this member, its named barrier */
	uint32_t so_a, so_b;       /* shared addr of this member's operator states: the voice's own (leader) / work copy */
	uint32_t plan_x;           /* ... of its executable plan */
	uint32_t cmd;              /* ... of the LEADER's command block */
	uint32_t per_warp;         /* bytes between consecutive members' areas */
	uint32_t lead_so, lead_plan;   /* the leader's operator states and master plan */
};

/* One team stretch, run by every member (those beyond t_eff only keep the barriers). */
__device__ __noinline__ void team_run(const TeamCtx &tc, uint32_t sb, int lane) {
	const uint32_t nrec = lds32(tc.cmd + TC_NREC), nops = lds32(tc.cmd + TC_NOPS), C = lds32(tc.cmd + TC_CHUNKS);
	const uint32_t P = lds32(tc.cmd + TC_P), t_eff = lds32(tc.cmd + TC_TEFF);
	const uint32_t fused = lds32(tc.cmd + TC_FUSED);       /* the plan spells a listed shape (render_fast.cuh) */
	const uint32_t w = tc.rank, nthreads = tc.T * 32u;
	const bool active = w < t_eff;
	const uint32_t L = P + 1u;
	const uint32_t a_w = active ? (uint32_t) ((uint64_t) w * C / t_eff) : 0u;
	const uint32_t a_next = active ? (uint32_t) ((uint64_t) (w + 1u) * C / t_eff) : 0u;
	const uint32_t s_w = w ? a_w - L : 0u;
	const uint32_t delta = tc.so_b - tc.lead_so;
	auto copy_ops = [&]() {            /* the voice's operator states -> this member's work copy (not the pads) */
		__syncwarp();
		for (uint32_t i = lane; i < nops * 46u; i += 32) {
			const uint32_t slot = i / 46u, wd = i % 46u;
			sts32(tc.so_b + slot * 192u + wd * 4u, lds32(tc.lead_so + slot * 192u + wd * 4u));
		}
		__syncwarp();
	};
	auto run = [&](uint32_t c0, uint32_t c1) {
		__syncwarp();
		if (c1 > c0) run_block_lowered<false>(sb, tc.plan_x, lane, c0 * (uint32_t) CHUNK, (c1 - c0) * (uint32_t) CHUNK);
		__syncwarp();
	};
	for (uint32_t p = 1; p <= P; ++p) {
		if (active && w + 1u < t_eff) {        /* a later member needs this one's count */
			copy_ops();
			if (lane == 0) {
				team_build_plan(tc.lead_plan, tc.plan_x, nrec, delta, p);
				if (w) team_patch(tc.plan_x, 0, s_w, false);
				if (!w) team_patch(tc.plan_x, p, 0, true);
			}
			uint32_t cur = s_w;
			if (w) {
				for (uint32_t k = 1; k <= p; ++k) {
					run(cur, s_w + k);
					cur = s_w + k;
					if (lane == 0) team_patch(tc.plan_x, k, 0, k == p);
				}
			}
			run(cur, a_next - L + p);              /* = the next member's level-p start */
			__syncwarp();
			if (lane == 0) {                       /* publish the counts */
				for (uint32_t a = tc.plan_x + PLAN_HDR; ; a += PLAN_REC) {
					const uint4 x = lds128u(a);
					const uint32_t kind = x.x & 0xffu;
					if (kind == P_STOP || a - tc.plan_x > PLAN_WALK_MAX) break;
					if (kind == X_COUNT1 || kind == X_COUNT2) sts32(x.z + OS_PAD0, lds32(x.z + OS_I0));
				}
			}
		}
		__threadfence_block();
		team_bar(tc.bar, nthreads);
		if (active && w && lane == 0) {
			/* this member's start values of the level-p accumulators: the voice's own + every
			 * earlier member's count (the master plan names the level-p operators) */
			for (uint32_t r = 0; r < nrec; ++r) {
				const uint32_t a = tc.lead_plan + PLAN_HDR + r * PLAN_REC;
				const uint4 x = lds128u(a);
				const uint32_t kind = x.x & 0xffu;
				if (kind == P_EXT || !(kind >= X_OSC0 && kind < X_RANGE) || ((x.y >> 24) & 0xfu) != p) continue;
				const uint32_t off = x.z - tc.lead_so;         /* the operator's offset in a member's area */
				uint32_t acc = lds32(x.z + OS_I0);
				for (uint32_t j = 0; j < w; ++j)
					acc += lds32(tc.so_b + off - (w - j) * tc.per_warp + OS_PAD0);
				sts32(tc.so_b + off + OS_PAD1, acc);
			}
		}
		__syncwarp();
	}
	if (active) {
		copy_ops();
		uint32_t vout = 0;
		if (lane == 0) {
			vout = team_build_plan(tc.lead_plan, tc.plan_x, nrec, delta, 0);
			if (w) {
				team_patch(tc.plan_x, 0, s_w, false);
				sts32(vout, (lds32(vout) & ~0xffu) | P_STOP);       /* lead-in: no output */
			}
		}
		vout = __shfl_sync(FULL, vout, 0);
		uint32_t cur = s_w;
		if (w) {
			for (uint32_t k = 1; k <= P; ++k) {
				run(cur, s_w + k);
				cur = s_w + k;
				if (lane == 0) team_patch(tc.plan_x, k, 0, false);
			}
			run(cur, a_w);
			cur = a_w;
			__syncwarp();
			if (lane == 0) sts32(vout, (lds32(vout) & ~0xffu) | X_VOUT);
		}
		__syncwarp();
		if (fused && a_next > cur)         /* the member's own range: the full plan, as one straight-line function */
			fused_run(fused, sb, tc.plan_x, lane, cur * (uint32_t) CHUNK, (a_next - cur) * (uint32_t) CHUNK);
		else run(cur, a_next);
		__syncwarp();
	}
	__threadfence_block();
	team_bar(tc.bar, nthreads);
}

/* The leader: offer a lowered steady stretch of `span` samples to the team.  Returns false when
 * it is not eligible (the caller renders it alone); else the stretch is rendered and the voice's
 * operator accumulators / look-back values (tc.so_a) are those after it. */
__device__ __noinline__ bool team_stretch(const TeamCtx &tc, uint32_t sb, int lane, uint32_t plan, uint32_t nrec,
		uint32_t nops, uint32_t span, uint32_t fused) {
	const uint32_t C = span / (uint32_t) CHUNK;
	uint32_t P = 0;
	if (lane == 0) P = team_analyse(plan + PLAN_HDR, nrec);
	P = __shfl_sync(FULL, P, 0);
	if (P == TEAM_INELIGIBLE) return false;
	/* every member but the first renders P + 1 lead-in chunks (and P counting passes): worth it
	 * from a few times that per member */
	uint32_t t_eff = C / (4u * (P + 2u));
	if (t_eff > tc.T) t_eff = tc.T;
	if (t_eff < 2u) return false;
	if (lane == 0) {
		sts32(tc.cmd + TC_OP, 1u); sts32(tc.cmd + TC_NREC, nrec); sts32(tc.cmd + TC_NOPS, nops);
		sts32(tc.cmd + TC_CHUNKS, C); sts32(tc.cmd + TC_P, P); sts32(tc.cmd + TC_TEFF, t_eff);
		sts32(tc.cmd + TC_FUSED, fused);
	}
	__syncwarp();
	__threadfence_block();
	team_bar(tc.bar, tc.T * 32u);
	team_run(tc, sb, lane);
	/* the last active member's accumulators and look-back values are the voice's */
	const uint32_t last = tc.so_b + (t_eff - 1u) * tc.per_warp;
	for (uint32_t i = lane; i < nops * 5u; i += 32) {
		const uint32_t slot = i / 5u, k = i % 5u;
		const uint32_t off = k == 4u ? OS_PREVS : OS_I0 + k * 4u;          /* i0, i1, prev_Is; prev_s */
		sts32(tc.so_a + slot * 192u + off, lds32(last + slot * 192u + off));
	}
	__syncwarp();
	return true;
}

/* members 1 .. T-1 */
__device__ __noinline__ void team_helper(const TeamCtx &tc, uint32_t sb, int lane) {
	for (;;) {
		team_bar(tc.bar, tc.T * 32u);
		if (lds32(tc.cmd + TC_OP) == 0u) return;
		team_run(tc, sb, lane);
	}
}

/* the leader, when its voice is done */
__device__ __forceinline__ void team_dismiss(const TeamCtx &tc, int lane) {
	if (lane == 0) sts32(tc.cmd + TC_OP, 0u);
	__syncwarp();
	__threadfence_block();
	team_bar(tc.bar, tc.T * 32u);
}
