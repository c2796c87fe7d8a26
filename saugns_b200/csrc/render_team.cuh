/* render_team.cuh -- part of kernels.cu (one translation unit; included inside namespace saugen):
 * parallelism ALONG TIME inside one voice.
 *
 * With fewer voices than resident warps, one warp per voice leaves the machine idle and a
 * launch takes as long as one voice's serial render.  In a steady stretch nothing but the
 * oscillator accumulators and the differentiator's one-sample look-back carries over from sample
 * to sample (wosc.h:129,135-169,247-262; the line trajectories are closed forms in the position,
 * line.c:27-37).  A TEAM of T warps of one CTA therefore splits a stretch of C chunks into T
 * ranges [a_w, a_w+1), w = 0..T-1.  Member w > 0 needs, at its start, every operator's phase
 * accumulator: phase0 + n * inc in closed form where the frequency is uniform (acc level 0); for
 * a frequency-modulated operator the sum of its rounded increments over everything before --
 * integer sums, so a prefix over the members' partial sums is bit-identical to the serial
 * accumulation.
 *
 * The plan's records are ordered in LEVELS: a record's output level O is the deepest accumulator
 * level it depends on; an oscillator's accumulator level A is 1 + the level of its frequency.
 * The stretch runs in PHASES q = 0..P (P = the deepest A), every member over its own range:
 *   phase q renders the records of level q -- inputs of lower levels come back from the
 *   stretch's CACHE in global memory (X_LOAD in the place of the record that produced them),
 *   outputs a later phase needs go there (X_SAVE) -- and COUNTS the increments of the
 *   oscillators with A = q + 1, whose frequency now exists; a barrier, a prefix over the
 *   members' counts, and phase q + 1 starts with exact accumulators.
 * Every record is rendered once per member (plus the lead-in below), whatever the FM depth.
 * prev_phase / prev_Is / prev_s: member w > 0 starts its level-q records L - q chunks early
 * (L = P + 1), without output: a level-q record gets its exact accumulator at chunk
 * a_w - L + q; its inputs (levels < q) are exact from that chunk on (they started a chunk earlier),
 * so all but its first sample there is exact, and everything from chunk a_w - L + q + 1 on.  An
 * oscillator r is therefore counted over [a_w - L + O_r, a_w+1 - L + O_r) by member w.
 * A TEAM CAN SPAN SEVERAL CTAs (one voice per CTA, K CTAs per voice, member = CTA part x T + warp): with one
 * voice alive, one SM's issue slots and shared memory bound the team.  The other CTAs mirror the leader's command
 * block, master plan and operator states in their own shared memory (same layout, so the plan's shared addresses
 * hold), fetched from the voice's mailbox in global memory when the leader's epoch word moves; barriers become
 * "every CTA's team barrier, then an arrival count in global memory", the counts of a phase and the last member's
 * oscillator state travel through the mailbox.  The launch is cooperative: all CTAs are resident.
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
constexpr uint32_t TEAM_MAX_P = TEAM_LEAD_CHUNKS - 1;
constexpr uint32_t TEAM_NONE = 0xffffu;
constexpr uint32_t OS_PAD0 = 184, OS_PAD1 = 188;      /* OpState::_pad: a member's count / its start value */
static_assert(offsetof(OpState, _pad) == OS_PAD0, "OpState::_pad offset");
/* the command block: the stretch, the voice's cache, and per record of the master plan
 * TC_INFO + 4r: phases that need the record's output (bits 0..7) | its cache slot << 8 */
constexpr uint32_t TC_OP = 0, TC_NREC = 4, TC_NOPS = 8, TC_CHUNKS = 12, TC_P = 16, TC_TEFF = 20,
	TC_FUSED = 24, TC_STRIDE = 28, TC_CACHE = 32, TC_INFO = 64;
constexpr uint32_t TC_SHARED_BYTES = TC_INFO + 4 * 68;      /* what the other CTAs of a team mirror */
constexpr uint32_t TC_SYNCS = TC_SHARED_BYTES;              /* this CTA's own: team-wide barriers passed in this launch */
static_assert(TC_SYNCS + 4 <= TEAM_CMD_BYTES, "command block");

/* developer aid (saugen_debug_team): the last stretch the first team of CTA 0 was offered --
 * [0] stretches offered, [1] P (0xff = not eligible), [2] t_eff, [3] records, [4] chunks, [5] cache slots,
 * [6] cycles in team_analyse, [7] cycles of the stretch, [8..15] cycles of the leader's phases,
 * [16] stretches split, [17] cycles between stretches (the leader alone) */
__device__ uint32_t g_team_dump[32];
__device__ __forceinline__ bool team_traced(const uint32_t bar) { return blockIdx.x == 0 && bar == 1u; }

__device__ __forceinline__ void team_bar(uint32_t id, uint32_t nthreads) {
	asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
/* what a record reads (buffers < 32, or TEAM_VAL = the record before it) and writes */
constexpr uint32_t TEAM_VAL = 0x100u;
struct RecIO { uint32_t out, freq, in[3], nin; };
__device__ __forceinline__ bool team_rec_io(uint32_t w0, uint32_t w1, RecIO &io) {
	const uint32_t kind = w0 & 0xffu, fl = (w0 >> 8) & 0xffu, xf = (w1 >> 16) & 0xffu;
	const uint32_t bufa = (w0 >> 16) & 0xffu, bufb = w0 >> 24;
	io.out = TEAM_NONE; io.freq = TEAM_NONE; io.nin = 0;
	auto in = [&](uint32_t b) { io.in[io.nin++] = b; };
	if (kind >= X_OSC0 && kind < X_RANGE) {
		const uint32_t v = kind - X_OSC0, fs = v / 3u, pm = v % 3u;
		if (fs == 1) io.freq = (xf & XF_SRC_VAL) ? TEAM_VAL : bufb;
		if (fs == 2) io.freq = (xf & XF_SRC_VAL) ? TEAM_VAL : (w1 >> 8) & 0xffu;
		if (pm == 1) in(w1 & 0xffu);
		if (pm == 2) in(TEAM_VAL);
		if (fl & PF_LAYER) in(bufa);
		io.out = bufa;
	} else if (kind == X_RANGE) {
		in((xf & XF_SRC_VAL) ? TEAM_VAL : (w1 & 0xffu));
		io.out = bufa;
	} else if (kind == X_VOUT) {
		in((xf & XF_SRC_VAL) ? TEAM_VAL : bufa);
	} else if (kind == P_LINE) {
		if (bufb != NO_BUF) in(bufb);
		io.out = bufa;
	} else if (kind == P_WHEAD) {
		const uint32_t e = (w1 >> 8) & 0xffu;
		if (e != NO_BUF) in(e);
		io.out = bufb;
	} else if (kind == P_RANGE) {
		in(bufa); in(bufb); in(w1 & 0xffu);
		io.out = bufa;
	} else if (kind == P_MIX) {
		if (bufb != NO_BUF) in(bufb);
		if (!(fl & PF_ACONST)) in(w1 & 0xffu);
		if (fl & PF_LAYER) in(bufa);
		io.out = bufa;
	} else {
		return false;              /* unlowered wave operators, noise, rumble, self-PM */
	}
	return true;
}

/* lane 0: acc level A (w1 bits 24..27; 15 = no accumulator) and output level O (bits 28..31; the
 * voice output: 15) of every record of a lowered plan, and TC_INFO.  Returns P = the deepest acc
 * level, or TEAM_INELIGIBLE. */
__device__ __noinline__ uint32_t team_analyse(uint32_t plan, uint32_t nrec, uint32_t cmd) {
	uint8_t blev[32], bprod[32];
	for (int i = 0; i < 32; ++i) { blev[i] = 0; bprod[i] = 0xff; }
	if (nrec > 68u) return TEAM_INELIGIBLE;
	uint32_t vlev = 0, P = 0;
	auto mx = [](uint32_t a, uint32_t b) { return a > b ? a : b; };
	for (uint32_t r = 0; r < nrec; ++r) {
		const uint32_t a = plan + r * PLAN_REC;
		const uint32_t w0 = lds32(a), w1 = lds32(a + 4);
		const uint32_t kind = w0 & 0xffu;
		sts32(cmd + TC_INFO + 4u * r, 0u);
		if (kind == P_EXT) continue;
		RecIO io;
		if (!team_rec_io(w0, w1, io)) return TEAM_INELIGIBLE;
		auto lev = [&](uint32_t b) -> uint32_t { return b == TEAM_VAL ? vlev : (b < 32u ? blev[b] : 0xffu); };
		uint32_t A = 15, O = 0;
		if (io.freq != TEAM_NONE) {
			if (lev(io.freq) == 0xffu) return TEAM_INELIGIBLE;
			A = 1u + lev(io.freq);
			O = A;
		} else if (kind >= X_OSC0 && kind < X_RANGE) {
			A = 0;
			if ((kind - X_OSC0) % 3u == 0u && lds32(a + 20) == 0u) return TEAM_INELIGIBLE;     /* stands still */
		}
		for (uint32_t i = 0; i < io.nin; ++i) {
			if (lev(io.in[i]) == 0xffu) return TEAM_INELIGIBLE;
			O = mx(O, lev(io.in[i]));
		}
		if (O > TEAM_MAX_P) return TEAM_INELIGIBLE;
		if (A != 15) P = mx(P, A);
		if (kind == X_VOUT) O = 15;
		if (io.out != TEAM_NONE) {
			if (io.out >= 32u) return TEAM_INELIGIBLE;
			blev[io.out] = (uint8_t) O;
		}
		vlev = (kind >= X_OSC0 && kind <= X_RANGE) ? O : 0u;
		sts32(a + 4, (w1 & 0x00ffffffu) | A << 24 | O << 28);
	}
	/* who needs whose output, in which phase: a record runs in phase O (the voice output: P); an
	 * oscillator with A >= 1 is also counted in phase A - 1, from its frequency */
	uint32_t prev = TEAM_NONE;
	for (uint32_t r = 0; r < nrec; ++r) {
		const uint32_t a = plan + r * PLAN_REC;
		const uint32_t w0 = lds32(a), w1 = lds32(a + 4);
		if ((w0 & 0xffu) == P_EXT) continue;
		RecIO io;
		team_rec_io(w0, w1, io);
		const uint32_t A = (w1 >> 24) & 0xfu;
		uint32_t O = w1 >> 28;
		if (O == 15) O = P;
		auto want = [&](uint32_t b, uint32_t phases) -> bool {
			const uint32_t p = b == TEAM_VAL ? prev : (uint32_t) bprod[b];
			if (p >= nrec) return false;                   /* read before anything wrote it */
			sts32(cmd + TC_INFO + 4u * p, lds32(cmd + TC_INFO + 4u * p) | phases);
			return true;
		};
		if (io.freq != TEAM_NONE && !want(io.freq, 1u << O | 1u << (A - 1u))) return TEAM_INELIGIBLE;
		for (uint32_t i = 0; i < io.nin; ++i)
			if (!want(io.in[i], 1u << O)) return TEAM_INELIGIBLE;
		if (io.out != TEAM_NONE) bprod[io.out] = (uint8_t) r;
		prev = r;
	}
	/* a cache slot for every output a LATER phase needs */
	uint32_t nslot = 0;
	for (uint32_t r = 0; r < nrec; ++r) {
		const uint32_t a = plan + r * PLAN_REC;
		if ((lds32(a) & 0xffu) == P_EXT) continue;
		uint32_t O = lds32(a + 4) >> 28;
		if (O == 15) O = P;
		const uint32_t need = lds32(cmd + TC_INFO + 4u * r) & 0xffu;
		if (need >> (O + 1u)) {
			if (nslot >= TEAM_SLOTS) return TEAM_INELIGIBLE;
			sts32(cmd + TC_INFO + 4u * r, need | nslot << 8);
			++nslot;
		}
	}
	return P;
}

/* Which records carry an operator's shared address in w2 (relocated for a member's copy). */
__device__ __forceinline__ bool rec_has_op(uint32_t kind) {
	return (kind >= X_OSC0 && kind < X_RANGE) || kind == X_VOUT || kind == X_COUNT1 || kind == X_COUNT2 ||
		kind == P_LINE || kind == P_WHEAD;
}

/* lane 0: the member's executable plan of phase q: header and the level-q records copied from the
 * master (operator addresses moved by `delta`), loads in the place of earlier levels' records this
 * phase reads, saves behind the records a later phase reads, counting records for the oscillators
 * with A = q + 1.  `cache`: this member's window of the voice's cache (slot stride `stride` floats).
 * Returns the shared address of the voice-output record in the copy (0 when the phase has none). */
__device__ __noinline__ uint32_t team_build_phase(uint32_t master, uint32_t exec, uint32_t nrec, uint32_t delta,
		uint32_t q, uint32_t P, uint32_t cmd, float *cache, uint32_t stride, uint32_t &nout) {
	for (uint32_t i = 0; i < PLAN_HDR; i += 16) {
		const uint4 h = lds128u(master + i);
		asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(exec + i), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w) : "memory");
	}
	uint32_t out = exec + PLAN_HDR, vout = 0;
	auto put = [&](const uint4 &x, const uint4 &y) {
		asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(out), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
		asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(out + 16), "r"(y.x), "r"(y.y), "r"(y.z), "r"(y.w) : "memory");
		out += PLAN_REC;
	};
	for (uint32_t r = 0; r < nrec; ++r) {
		const uint32_t a = master + PLAN_HDR + r * PLAN_REC;
		uint4 x = lds128u(a);
		const uint4 y = lds128u(a + 16);
		const uint32_t kind = x.x & 0xffu, fl = (x.x >> 8) & 0xffu, xf = (x.y >> 16) & 0xffu;
		if (kind == P_EXT) continue;
		const bool osc = kind >= X_OSC0 && kind < X_RANGE;
		const bool ext = (kind == P_WLEAF || kind == P_WTAIL || osc) && (fl & PF_AEXT);
		const uint32_t A = (x.y >> 24) & 0xfu;
		uint32_t O = x.y >> 28;
		if (O == 15) O = P;
		const uint32_t info = lds32(cmd + TC_INFO + 4u * r), need = info & 0xffu, slot = info >> 8;
		const uint32_t obuf = kind == P_WHEAD ? x.x >> 24 : (x.x >> 16) & 0xffu;
		const bool lowered = osc || kind == X_RANGE;
		const uint64_t cp = reinterpret_cast<uint64_t>(cache + (size_t) slot * stride);
		if (O == q) {
			uint4 xr = x;
			if (rec_has_op(kind)) xr.z += delta;
			if (kind == X_VOUT) vout = out;
			put(xr, y);
			if (ext) put(lds128u(a + PLAN_REC), lds128u(a + PLAN_REC + 16));
			if (need >> (q + 1u)) {
				const uint4 sv = make_uint4(X_SAVE | (lowered ? 0u : 1u) << 8 | obuf << 16, 0u, (uint32_t) cp, (uint32_t) (cp >> 32));
				put(sv, make_uint4(0u, 0u, 0u, 0u));
			}
		} else if (O < q && ((need >> q) & 1u)) {
			const bool st = !lowered || (xf & XF_ST);
			const uint4 ld = make_uint4(X_LOAD | (st ? 1u : 0u) << 8 | obuf << 16, 0u, (uint32_t) cp, (uint32_t) (cp >> 32));
			put(ld, make_uint4(0u, 0u, 0u, 0u));
		}
		if (osc && A == q + 1u) {
			/* counted now: its frequency exists (flags: which counting kind, for the window's on / off) */
			const uint32_t fs = (kind - X_OSC0) / 3u;
			uint4 xc = x;
			xc.x = (x.x & ~0xffffu) | X_NOP | (fs == 1 ? 1u : 2u) << 8;
			xc.z += delta;
			put(xc, y);
		}
	}
	sts32(out, P_STOP);
	nout = (out - exec - PLAN_HDR) / PLAN_REC;
	return vout;
}

/* lane 0, at the start of phase q: the accumulators of the phase's oscillators get their start
 * values -- level 0: the closed form at chunk `at` (uniform increment in w5); else the member's
 * prefix (OS_PAD1; member 0: the voice's own) -- and the counting records start from zero. */
__device__ __noinline__ void team_phase_init(uint32_t exec, uint32_t at, uint32_t w, uint32_t delta) {
	for (uint32_t a = exec + PLAN_HDR; ; a += PLAN_REC) {
		const uint4 x = lds128u(a);
		const uint32_t kind = x.x & 0xffu;
		if (kind == P_STOP || a - exec > PLAN_WALK_MAX + TEAM_SLOTS * PLAN_REC) break;
		if (kind == X_NOP) { sts32(x.z + OS_I0, 0u); continue; }
		if (!(kind >= X_OSC0 && kind < X_RANGE)) continue;
		const uint32_t A = (x.y >> 24) & 0xfu;
		if (A == 0) sts32(x.z + OS_I0, lds32(x.z - delta + OS_I0) + lds32(a + 20) * (at * (uint32_t) CHUNK));
		else sts32(x.z + OS_I0, w ? lds32(x.z + OS_PAD1) : lds32(x.z - delta + OS_I0));
	}
}

struct TeamCtx {
	uint32_t T, rank, bar;     /* members of this CTA, this member among them, their named barrier */
	uint32_t K, part;          /* CTAs of the team, this CTA among them (member = part * T + rank) */
	uint32_t *hdr;             /* K > 1: the voice's {arrivals, epoch} words ... */
	unsigned char *mail;       /* ... and its mailbox (device_types.h:team_mail_*) */
	uint32_t plan_bytes, max_ops;  /* (the mailbox's layout) */
	uint32_t so_a, so_b;       /* shared addr of this member's operator states: the voice's own (leader) / work copy */
	uint32_t plan_x;           /* ... of its executable plan */
	uint32_t cmd;              /* ... of the LEADER's command block */
	uint32_t per_warp;         /* bytes between consecutive members' areas */
	uint32_t lead_so, lead_plan;   /* the leader's operator states and master plan */
};

/* Every member of the team, on every CTA it spans: the CTA's named barrier; with K > 1 then one arrival per CTA
 * in global memory (the n-th barrier of the launch is passed when n * K have arrived) and the named barrier again. */
__device__ __forceinline__ void team_sync(const TeamCtx &tc, int lane) {
	__threadfence_block();
	team_bar(tc.bar, tc.T * 32u);
	if (tc.K > 1u) {
		if (tc.rank == 0u && lane == 0) {
			const uint32_t n = lds32(tc.cmd + TC_SYNCS) + 1u;
			sts32(tc.cmd + TC_SYNCS, n);
			__threadfence();
			atomicAdd(tc.hdr, 1u);
			while (ld_acquire_gpu(tc.hdr) < n * tc.K) __nanosleep(64);
		}
		team_bar(tc.bar, tc.T * 32u);
	}
}

/* One team stretch, run by every member (those beyond t_eff only keep the barriers). */
__device__ __noinline__ void team_run(const TeamCtx &tc, uint32_t sb, int lane) {
	const uint32_t nrec = lds32(tc.cmd + TC_NREC), nops = lds32(tc.cmd + TC_NOPS), C = lds32(tc.cmd + TC_CHUNKS);
	const uint32_t P = lds32(tc.cmd + TC_P), t_eff = lds32(tc.cmd + TC_TEFF);
	const uint32_t fused = lds32(tc.cmd + TC_FUSED);       /* the plan spells a listed shape (render_fast.cuh) */
	const uint32_t stride = lds32(tc.cmd + TC_STRIDE);
	float *cache = reinterpret_cast<float*>((uint64_t) lds32(tc.cmd + TC_CACHE) | ((uint64_t) lds32(tc.cmd + TC_CACHE + 4) << 32));
	const uint32_t w = tc.part * tc.T + tc.rank;
	const bool active = w < t_eff;
	uint32_t *mcounts = reinterpret_cast<uint32_t*>(tc.mail + team_mail_counts_off(tc.plan_bytes, tc.max_ops));
	uint32_t *final_st = reinterpret_cast<uint32_t*>(tc.mail + team_mail_final_off(tc.plan_bytes, tc.max_ops));
	const uint32_t L = P + 1u;
	const uint32_t a_w = active ? (uint32_t) ((uint64_t) w * C / t_eff) : 0u;
	const uint32_t a_next = active ? (uint32_t) ((uint64_t) (w + 1u) * C / t_eff) : 0u;
	const uint32_t delta = tc.so_b - tc.lead_so;
	/* members' ranges overlap by their lead-in chunks: every member has its own window of the cache */
	if (cache) cache += (size_t) w * L * (uint32_t) CHUNK;
	if (active) {          /* the voice's operator states -> this member's work copy (not the pads) */
		__syncwarp();
		for (uint32_t i = lane; i < nops * 46u; i += 32) {
			const uint32_t slot = i / 46u, wd = i % 46u;
			sts32(tc.so_b + slot * 192u + wd * 4u, lds32(tc.lead_so + slot * 192u + wd * 4u));
		}
		__syncwarp();
	}
	for (uint32_t q = 0; q <= P; ++q) {
		const long long tq = clock64();
		if (active) {
			const uint32_t start = w ? a_w - L + q : 0u;
			const bool counts = q < P && w + 1u < t_eff;           /* a later member needs this one's counts */
			uint32_t vout = 0, fq = 0;
			if (lane == 0) {
				uint32_t nx = 0;
				vout = team_build_phase(tc.lead_plan, tc.plan_x, nrec, delta, q, P, tc.cmd, cache, stride, nx);
				team_phase_init(tc.plan_x, start, w, delta);
				/* P = 0: the full plan, whose shape the leader looked up; else this phase's own */
				fq = P == 0u ? fused : (fused ? fused_match(tc.plan_x, nx, false) : 0u);
			}
			vout = __shfl_sync(FULL, vout, 0);
			fq = __shfl_sync(FULL, fq, 0);
			uint32_t cur = start;
			while (cur < a_next) {
				uint32_t nxt = a_next;
				if (lane == 0) {
					/* counting windows: oscillator r over [a_w - L + O_r, a_w+1 - L + O_r) (member 0: from 0) */
					for (uint32_t a = tc.plan_x + PLAN_HDR; ; a += PLAN_REC) {
						const uint4 x = lds128u(a);
						const uint32_t kind = x.x & 0xffu;
						if (kind == P_STOP || a - tc.plan_x > PLAN_WALK_MAX + TEAM_SLOTS * PLAN_REC) break;
						if (kind != X_NOP && kind != X_COUNT1 && kind != X_COUNT2) continue;
						uint32_t O = x.y >> 28;
						if (O == 15) O = P;
						const uint32_t lo = w ? a_w - L + O : 0u, hi = a_next - L + O;
						const bool on = counts && cur >= lo && cur < hi;
						const uint32_t k = on ? (((x.x >> 8) & 0xffu) == 1u ? X_COUNT1 : X_COUNT2) : X_NOP;
						sts32(a, (x.x & ~0xffu) | k);
						if (counts && lo > cur && lo < nxt) nxt = lo;
						if (counts && hi > cur && hi < nxt) nxt = hi;
					}
					if (vout) {            /* lead-in: no output */
						const bool on = !w || cur >= a_w;
						sts32(vout, (lds32(vout) & ~0xffu) | (on ? X_VOUT : P_STOP));
						if (!on && a_w < nxt) nxt = a_w;
					}
				}
				nxt = __shfl_sync(FULL, nxt, 0);
				__syncwarp();
				if (fq && (!vout || !w || cur >= a_w))          /* the phase's plan as one straight-line function */
					fused_run(fq, sb, tc.plan_x, lane, cur * (uint32_t) CHUNK, (nxt - cur) * (uint32_t) CHUNK);
				else
					run_block_lowered<false, true>(sb, tc.plan_x, lane, cur * (uint32_t) CHUNK, (nxt - cur) * (uint32_t) CHUNK);
				__syncwarp();
				cur = nxt;
			}
			if (counts && lane == 0) {                     /* publish the counts */
				for (uint32_t a = tc.plan_x + PLAN_HDR; ; a += PLAN_REC) {
					const uint4 x = lds128u(a);
					const uint32_t kind = x.x & 0xffu;
					if (kind == P_STOP || a - tc.plan_x > PLAN_WALK_MAX + TEAM_SLOTS * PLAN_REC) break;
					if (kind == X_NOP || kind == X_COUNT1 || kind == X_COUNT2) {
						const uint32_t cnt = lds32(x.z + OS_I0);
						sts32(x.z + OS_PAD0, cnt);
						if (tc.K > 1u) __stcg(mcounts + w * TEAM_MAIL_OPS + (x.z - tc.so_b) / 192u, cnt);
					}
				}
			}
		}
		team_sync(tc, lane);
		if (!w && lane == 0 && q < 8u && team_traced(tc.bar)) g_team_dump[8 + q] = (uint32_t) (clock64() - tq);
		if (q < P && active && w) {
			/* this member's start values of the accumulators counted in this phase: the voice's own +
			 * every earlier member's count (the master plan names them; the lanes share the members) */
			for (uint32_t r = 0; r < nrec; ++r) {
				const uint32_t a = tc.lead_plan + PLAN_HDR + r * PLAN_REC;
				const uint4 x = lds128u(a);
				const uint32_t kind = x.x & 0xffu;
				if (kind == P_EXT || !(kind >= X_OSC0 && kind < X_RANGE) || ((x.y >> 24) & 0xfu) != q + 1u) continue;
				const uint32_t off = x.z - tc.lead_so;         /* the operator's offset in a member's area */
				uint32_t sum = 0;
				if (tc.K > 1u) {
					for (uint32_t j = lane; j < w; j += 32) sum += __ldcg(mcounts + j * TEAM_MAIL_OPS + off / 192u);
				} else {
					for (uint32_t j = lane; j < w; j += 32)
						sum += lds32(tc.so_b + off - (w - j) * tc.per_warp + OS_PAD0);
				}
				sum = __reduce_add_sync(FULL, sum);
				if (lane == 0) sts32(tc.so_b + off + OS_PAD1, lds32(x.z + OS_I0) + sum);
			}
		}
		__syncwarp();
	}
	if (tc.K > 1u) {
		/* the last active member's accumulators and look-back values, on their way to the voice */
		if (w + 1u == t_eff)
			for (uint32_t i = lane; i < nops * 5u; i += 32) {
				const uint32_t slot = i / 5u, k = i % 5u;
				__stcg(final_st + i, lds32(tc.so_b + slot * 192u + (k == 4u ? OS_PREVS : OS_I0 + k * 4u)));
			}
		team_sync(tc, lane);
	}
}

/* The leader: offer a lowered steady stretch of `span` samples to the team.  Returns false when
 * it is not eligible (the caller renders it alone); else the stretch is rendered and the voice's
 * operator accumulators / look-back values (tc.so_a) are those after it.  cache / stride: the
 * voice's cache in global memory (TEAM_SLOTS slots of `stride` floats; none: plans with P > 0 are
 * not eligible). */
__device__ __noinline__ bool team_stretch(const TeamCtx &tc, uint32_t sb, int lane, uint32_t plan, uint32_t nrec,
		uint32_t nops, uint32_t span, uint32_t fused, float *cache, uint32_t stride, const uint4 *kinfo, uint32_t &kP) {
	const uint32_t C = span / (uint32_t) CHUNK;
	uint32_t P = 0;
	const long long t0 = clock64();
	if (kP != 0u && kinfo) {
		/* a kept plan whose analysis was kept with it (render_kernel.cuh): levels are in the records */
		P = kP == 0x100u ? TEAM_INELIGIBLE : kP - 1u;
		if (P != TEAM_INELIGIBLE)
			for (uint32_t i = lane; i < 17u; i += 32) {
				const uint4 x = __ldcg(kinfo + i);
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(tc.cmd + TC_INFO + 16u * i), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
			}
		__syncwarp();
	} else {
		if (lane == 0) P = team_analyse(plan + PLAN_HDR, nrec, tc.cmd);
		__syncwarp();              /* (it wrote the records' levels) */
		P = __shfl_sync(FULL, P, 0);
		kP = P == TEAM_INELIGIBLE ? 0x100u : P + 1u;
	}
	if (lane == 0 && team_traced(tc.bar)) {
		uint32_t ns = 0;
		if (P != TEAM_INELIGIBLE)
			for (uint32_t r = 0; r < nrec; ++r) {
				uint32_t O = lds32(plan + PLAN_HDR + r * PLAN_REC + 4) >> 28;
				if (O == 15) O = P;
				if ((lds32(tc.cmd + TC_INFO + 4u * r) & 0xffu) >> (O + 1u)) ++ns;
			}
		g_team_dump[0]++; g_team_dump[1] = P; g_team_dump[3] = nrec; g_team_dump[4] = C; g_team_dump[5] = ns;
		g_team_dump[6] = (uint32_t) (clock64() - t0);
	}
	if (P == TEAM_INELIGIBLE) return false;
	if (P > 0 && !cache) return false;
	/* every member but the first renders up to P + 1 lead-in chunks: worth it from twice that per member */
	uint32_t t_eff = C / (tc.K > 1u ? (P + 2u) : 2u * (P + 2u));
	if (t_eff > tc.T * tc.K) t_eff = tc.T * tc.K;
	if (t_eff > TEAM_MAX_MEMBERS) t_eff = TEAM_MAX_MEMBERS;
	if (t_eff < 2u) return false;
	if (tc.K > 1u && (nops > TEAM_MAIL_OPS || nops > tc.max_ops || PLAN_HDR + (nrec + 1u) * PLAN_REC > tc.plan_bytes)) return false;
	if (P > 0 && (uint64_t) span + (uint64_t) (t_eff - 1u) * (P + 1u) * (uint32_t) CHUNK > stride) return false;
	if (lane == 0) {
		sts32(tc.cmd + TC_OP, 1u); sts32(tc.cmd + TC_NREC, nrec); sts32(tc.cmd + TC_NOPS, nops);
		sts32(tc.cmd + TC_CHUNKS, C); sts32(tc.cmd + TC_P, P); sts32(tc.cmd + TC_TEFF, t_eff);
		sts32(tc.cmd + TC_FUSED, fused); sts32(tc.cmd + TC_STRIDE, stride);
		const uint64_t cp = reinterpret_cast<uint64_t>(cache);
		sts32(tc.cmd + TC_CACHE, (uint32_t) cp); sts32(tc.cmd + TC_CACHE + 4, (uint32_t) (cp >> 32));
	}
	__syncwarp();
	if (tc.K > 1u) {
		/* the other CTAs' copy: command block, master plan with its end mark, operator states; then the epoch */
		const uint32_t np = (PLAN_HDR + (nrec + 1u) * PLAN_REC) / 16u, no = nops * 12u, nc = TC_SHARED_BYTES / 16u;
		uint4 *m = reinterpret_cast<uint4*>(tc.mail);
		for (uint32_t i = lane; i < nc; i += 32) __stcg(m + i, lds128u(tc.cmd + 16u * i));
		for (uint32_t i = lane; i < np; i += 32) __stcg(m + team_mail_plan_off() / 16u + i, lds128u(plan + 16u * i));
		for (uint32_t i = lane; i < no; i += 32) __stcg(m + team_mail_ops_off(tc.plan_bytes) / 16u + i, lds128u(tc.lead_so + 16u * i));
		__threadfence();
		__syncwarp();
		if (lane == 0) atomicAdd(tc.hdr + 1, 1u);
	}
	__threadfence_block();
	team_bar(tc.bar, tc.T * 32u);
	team_run(tc, sb, lane);
	if (lane == 0 && team_traced(tc.bar)) {
		g_team_dump[2] = t_eff; g_team_dump[7] = (uint32_t) (clock64() - t0); g_team_dump[16]++;
	}
	/* the last active member's accumulators and look-back values are the voice's */
	const uint32_t last = tc.so_b + (t_eff - 1u) * tc.per_warp;
	const uint32_t *final_st = reinterpret_cast<const uint32_t*>(tc.mail + team_mail_final_off(tc.plan_bytes, tc.max_ops));
	for (uint32_t i = lane; i < nops * 5u; i += 32) {
		const uint32_t slot = i / 5u, k = i % 5u;
		const uint32_t off = k == 4u ? OS_PREVS : OS_I0 + k * 4u;          /* i0, i1, prev_Is; prev_s */
		sts32(tc.so_a + slot * 192u + off, tc.K > 1u ? __ldcg(final_st + i) : lds32(last + slot * 192u + off));
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

/* every warp of a CTA other than the leader's (K > 1): warp 0 waits for the leader's next epoch, mirrors the
 * command block, master plan and operator states from the mailbox, and releases the CTA's members */
__device__ __noinline__ void team_remote(const TeamCtx &tc, uint32_t sb, int lane) {
	uint32_t seen = 0;
	for (;;) {
		if (tc.rank == 0u) {
			if (lane == 0) while (ld_acquire_gpu(tc.hdr + 1) == seen) __nanosleep(128);
			__syncwarp();
			++seen;
			const uint4 *m = reinterpret_cast<const uint4*>(tc.mail);
			for (uint32_t i = lane; i < TC_SHARED_BYTES / 16u; i += 32) {
				const uint4 x = __ldcg(m + i);
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(tc.cmd + 16u * i), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
			}
			__syncwarp();
			if (lds32(tc.cmd + TC_OP) != 0u) {
				const uint32_t np = (PLAN_HDR + (lds32(tc.cmd + TC_NREC) + 1u) * PLAN_REC) / 16u, no = lds32(tc.cmd + TC_NOPS) * 12u;
				for (uint32_t i = lane; i < np; i += 32) {
					const uint4 x = __ldcg(m + team_mail_plan_off() / 16u + i);
					asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(tc.lead_plan + 16u * i), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
				}
				for (uint32_t i = lane; i < no; i += 32) {
					const uint4 x = __ldcg(m + team_mail_ops_off(tc.plan_bytes) / 16u + i);
					asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(tc.lead_so + 16u * i), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
				}
			}
			__syncwarp();
			__threadfence_block();
		}
		team_bar(tc.bar, tc.T * 32u);
		if (lds32(tc.cmd + TC_OP) == 0u) return;
		team_run(tc, sb, lane);
	}
}

/* the leader, when its voice is done */
__device__ __forceinline__ void team_dismiss(const TeamCtx &tc, int lane) {
	if (lane == 0) sts32(tc.cmd + TC_OP, 0u);
	__syncwarp();
	if (tc.K > 1u && lane == 0) {
		__stcg(reinterpret_cast<uint32_t*>(tc.mail) + TC_OP / 4u, 0u);
		__threadfence();
		atomicAdd(tc.hdr + 1, 1u);
	}
	__threadfence_block();
	team_bar(tc.bar, tc.T * 32u);
}
