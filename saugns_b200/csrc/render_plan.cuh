/* render_plan.cuh -- part of kernels.cu (one translation unit; included inside namespace saugen):
 * the steady-stretch fast path: plan building (steady_plan), the plan-driven chunk loop (run_block_fast), per-stretch bookkeeping (steady_update). */
#pragma once

/* ---- steady-block fast path --------------------------------------------- *
 * Most of a render is spent in blocks where nothing changes shape: a whole
 * 1024-sample reference block (BUF_LEN, generator.c:28) lies inside one
 * inter-event segment, every operator of the voice outlasts it, every line
 * either holds its value or is on a trajectory that does not end inside the
 * block, no differentiator restart or self-PM is pending.  For such a block
 * the state machines of sauLine_run / run_block need no per-chunk decisions:
 * the reference itself advances them once per block.  steady_check() proves
 * the block is of that kind (else the general interpreter above renders it),
 * run_chunk_fast() renders its chunks with read-only line state and only the
 * oscillator accumulators written back, steady_update() then advances lines
 * and operator times by one block exactly as sauLine_run / sauLine_skip /
 * run_block do for len = 1024 (line.c:417-473, generator.c:716-728).
 * Supported bytecode: the wave-operator forms (HEAD/TAIL/LEAF, ENTER + LINE +
 * RANGE for FM carriers), static pan; anything else makes steady_check fail. */

/* For how many whole 1024-sample blocks, at most `k`, a run line stays steady: it
 * holds its value, or is on a trajectory that ends after them with no ratio
 * reconciliation due (line.c:358-369).  0 = not even one. */
__device__ __forceinline__ uint32_t line_span(const OpState *o, int li, uint32_t k) {
	const uint32_t flags = LM_FLAGS(o->lmeta[li]);
	if (!(flags & SAUABI_LINEP_GOAL)) return k;
	const bool gr = (flags & SAUABI_LINEP_GOAL_RATIO) != 0, sr = (flags & SAUABI_LINEP_STATE_RATIO) != 0;
	const uint32_t pos = o->line[li].pos, end = o->line[li].end;
	if (gr != sr || pos >= end) return 0;
	const uint32_t a = (end - pos - 1u) / (uint32_t) REF_BLOCK;      /* end - pos > a * 1024 */
	return a < k ? a : k;
}
/* ... and an operator keeps running (run_block, generator.c:694-698) */
__device__ __forceinline__ uint32_t op_span(const OpState *o, uint32_t k) {
	if (o->flags & ON_TIME_INF) return k;
	const uint32_t a = o->time / (uint32_t) REF_BLOCK;
	return a < k ? a : k;
}

/* ---- block plan ---------------------------------------------------------- *
 * steady_plan() proves the block steady and, while it walks the bytecode, writes
 * the block's PLAN into the warp's shared memory: one 32-byte record per
 * instruction that does something per chunk (ENTER / VPAN / END and skipped
 * lines drop out), with everything that is fixed for the block resolved: the
 * operator's shared address, its table, its differentiator constants, whether
 * its amplitude holds one value, and whether its FREQUENCY is one value over
 * the block (a line without a goal, times a parent frequency that is itself
 * uniform).  A uniform frequency f makes sauPhasor_fill (wosc.h:135-169) a
 * closed form: every sample adds the same inc = lrintf(coeff * f), so sample i
 * of the chunk is at phase0 + (i + 1) * inc in wrap-around uint32 arithmetic --
 * bit-identical to the serial accumulation, without conversions or a scan. */
enum : uint32_t { P_STOP = 0,          /* end mark after the last record (steady_plan's finish) */
	P_LINE = 1, P_WHEAD, P_WTAIL, P_WLEAF, P_PHASE, P_WOSC, P_RANGE, P_VOUT,
	P_NOISE, P_CYCLE, P_RASG, P_MIX, P_WSELF,
	P_EXT };           /* second slot of the record before it (never dispatched on its own) */
enum : uint32_t {
	PF_LAYER = 1, PF_WAVEENV = 2,
	PF_FUNI = 4,       /* frequency (or the LINE's value) is uniform over the block: w6 holds the
	                    * value (LINE, WHEAD) or the phase increment (WTAIL, WLEAF, PHASE) */
	PF_FMUL = 8,       /* WHEAD / WLEAF, not uniform: the frequency is the constant w6 times the
	                    * (varying) multiplier buffer: a ratio to a modulated parent frequency */
	PF_ACONST = 16,    /* amplitude line holds av */
	PF_ABUF = 32,      /* amplitude comes from work buffer c (the operator has amplitude modulators) */
	PF_AEXT = 64,      /* amplitude line on a lin / xpe / lge trajectory: its constants for the stretch
	                    * are in the P_EXT slot after the record (w0 type << 8, w1 position or centred
	                    * position at the stretch start, w2 1/time, w3 slope / span, w4 offset) */
};
constexpr uint32_t PLAN_FBUF = 32 * FAST_NS * 4;    /* FastCfg<FAST_NS>::FBUF_BYTES */
constexpr uint32_t PLAN_WALK_MAX = 67 * 32;   /* bytes: more than any plan holds (runtime.cpp: np <= 64) */
constexpr uint32_t PLAN_HDR = 64;     /* bytes of the plan header (two slots) before the first record */
constexpr uint32_t PLAN_REC = 32;     /* bytes: w0 kind|flags<<8|a<<16|b<<24, w1 c|e<<8|line<<16,
                                       * w2 operator state (shared address), w3 table (shared address),
                                       * w4 diff_scale, w5 diff_offset, w6 uniform value / phase increment, w7 av */

/* every lane walks the bytecode (each needs the result); lane 0 alone writes the plan */
__device__ __forceinline__ void plan_put(uint32_t plan, uint32_t n, uint32_t w0, uint32_t w1, uint32_t w2,
		uint32_t w3, float w4, float w5, float w6, float w7) {
	if ((threadIdx.x & 31u) != 0u) return;
	const uint32_t a = plan + n * PLAN_REC;
	asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
	sts128(a + 16, make_float4(w4, w5, w6, w7));
}

/* kb = the whole blocks ahead in this unit.  Returns blocks << 16 | records: how
 * many of those blocks are steady as ONE stretch (the plan holds for all of them:
 * nothing in it depends on the block), and the number of plan records; 0 = the
 * next block is not steady (or there is no room for its plan). */
__device__ __noinline__ uint32_t steady_plan(OpState *sops, uint32_t so, uint32_t st, uint32_t wave_mask,
		const WaveCoeffs *wc, const Instr *code, uint32_t code_len, uint32_t plan, uint32_t cap, uint32_t kb,
		uint32_t sb, float coeff) {
	uint32_t seen = 0;         /* operator slots already visited (< 32 of them) */
	uint32_t uni = 0;          /* work buffers (< 32) holding one value over the block */
	/* A uniform value is known NOW: it is kept in the buffer's own first word (every
	 * lane in its own slot) while the plan is built, so that a child's ratio
	 * frequency and the operator's phase increment are worked out here, once.  A
	 * frequency buffer that nothing reads as a vector before its operator's phase
	 * fill (need) then has no per-chunk use at all: its HEAD record is dropped. */
	uint32_t need = 0;
	uint32_t line_uni = 0;     /* buffers filled by a uniform LINE record that nothing has touched since */
	uint8_t head_rec[32];
	uint32_t killed = 0;
	const bool lane0 = (threadIdx.x & 31u) == 0u;
	uint32_t lstack = 0, depth = 0;    /* layer flags of the unfused operators being walked */
	uint32_t entered = 0;              /* operator slots that came in through an ENTER */
	uint32_t selfmask = 0;             /* operator slots whose pm_a line runs: self-PM (generator.c:485-490) */
	uint32_t other = 0;                /* the plan has serial self-PM records (bit 31 of the result) */
	uint32_t n = 0;
	plan += PLAN_HDR;          /* the first two slots are the header (render_units) */
	cap = cap > 3u ? cap - 3u : 0u;   /* ... and one slot stays free for the end mark */
	auto is_uni = [&](uint32_t b) { return b < 32 && ((uni >> b) & 1u); };
	auto touch = [&](uint32_t b) { if (b < 32) { need |= 1u << b; line_uni &= ~(1u << b); } };   /* read as a vector */
	auto dirty = [&](uint32_t b) {                                             /* rewritten */
		if (b < 32) { uni &= ~(1u << b); need |= 1u << b; line_uni &= ~(1u << b); }
	};
	auto uval = [&](uint32_t b) { return lds32f(sb + b * PLAN_FBUF); };
	auto set_uni = [&](uint32_t b, float f) {
		if (b < 32) { uni |= 1u << b; need &= ~(1u << b); sts32(sb + b * PLAN_FBUF, __float_as_uint(f)); }
	};
	auto finish = [&](uint32_t nrec) -> uint32_t {
		if (!nrec) return 0u;
		__syncwarp();                              /* lane 0's records are in place */
		if (killed) {                              /* close the gaps the dropped records left */
			uint32_t w = 0;
			for (uint32_t r = 0; r < nrec; ++r) {
				const uint4 x = lds128u(plan + r * PLAN_REC), y = lds128u(plan + r * PLAN_REC + 16);
				__syncwarp();                      /* every lane has read slot r before slot w <= r is rewritten */
				if ((x.x & 0xffu) == 0u) continue;
				if (w != r && lane0) {
					asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(plan + w * PLAN_REC),
							"r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
					asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(plan + w * PLAN_REC + 16),
							"r"(y.x), "r"(y.y), "r"(y.z), "r"(y.w) : "memory");
				}
				++w;
			}
			__syncwarp();
			nrec = w;
		}
		/* the end mark: the chunk loop runs until it meets it (no record count to keep) */
		if (lane0) sts32(plan + nrec * PLAN_REC, P_STOP);
		__syncwarp();
		return other | kb << 16 | nrec;
	};
	/* the operator's phase fill takes its frequency from uniform buffer b: the
	 * increment is known, and the HEAD that filled b may have nothing left to do */
	auto uni_inc = [&](uint32_t b) -> uint32_t {
		const uint32_t inc = ftoi_lo32(coeff * uval(b));
		if (!((need >> b) & 1u)) { if (lane0) sts32(plan + head_rec[b] * PLAN_REC, 0u); ++killed; }
		return inc;
	};
	uint4 raw_next = __ldg(reinterpret_cast<const uint4*>(code));
	for (uint32_t pc = 0; pc < code_len; ++pc) {
		const uint4 raw = raw_next;
		if (pc + 1 < code_len) raw_next = __ldg(reinterpret_cast<const uint4*>(code + pc + 1));
		Instr in;
		memcpy(&in, &raw, sizeof(in));
		const OpState *o = sops + in.op;
		const uint32_t opa = so + in.op * (uint32_t) sizeof(OpState);
		bool head = false, tail = false;
		if (n >= cap) return 0;
		switch (in.opcode) {
		case I_WLEAF: head = tail = true; break;
		case I_WHEAD: head = true; break;
		case I_WTAIL: tail = true; break;
		case I_ENTER:
			if (!(kb = op_span(o, kb))) return 0;
			if (in.op >= 32 || (seen & (1u << in.op))) return 0;
			seen |= 1u << in.op;
			if ((in.flags & F_LAYER_PMA) || depth >= 31) return 0;   /* self-PM modulators: general path */
			lstack = (lstack << 1) | ((in.flags & F_LAYER) ? 1u : 0u);
			++depth;
			entered |= 1u << in.op;
			break;
		case I_LEAVE:                  /* full chunks: nothing to zero-fill (generator.c:716-725) */
			if (!depth) return 0;
			lstack >>= 1;
			--depth;
			break;
		case I_NOISE:                                                /* run_block_noiseg, generator.c:527-541 */
			plan_put(plan, n++, P_NOISE | (uint32_t) in.a << 16, 0u, opa, 0u, 0.f, 0.f, 0.f, 0.f);
			dirty(in.a);
			break;
		case I_CYCLOR:                                               /* run_block_rasg, generator.c:609-664 */
			if (in.e != NO_BUF) return 0;                            /* fPM: general path */
			touch(in.c);
			if (in.d != NO_BUF) touch(in.d);
			plan_put(plan, n++, P_CYCLE | (uint32_t) in.a << 16 | (uint32_t) in.b << 24,
					(uint32_t) in.c | (uint32_t) in.d << 8, opa, 0u, 0.f, 0.f, 0.f, 0.f);
			dirty(in.a); dirty(in.b);
			break;
		case I_RASG:
			if (in.flags & F_HAS_APMODS) return 0;
			touch(in.b);
			{
				/* self-PM (sauRasG_run_selfmod, rasg.h:242-294): PF_FUNI marks it, c = amount buffer */
				const bool self = in.op < 32 && ((selfmask >> in.op) & 1u);
				if (self) touch(in.c);
				if (self) other = 0x80000000u;
				plan_put(plan, n++, P_RASG | (self ? PF_FUNI : 0u) << 8 | (uint32_t) in.a << 16 | (uint32_t) in.b << 24,
						(uint32_t) in.c, opa, 0u, 0.f, 0.f, 0.f, 0.f);
			}
			dirty(in.a);
			break;
		case I_MIX:                                                  /* generator.c:384-440 */
			if (!depth) return 0;
			if (in.b != NO_BUF) touch(in.b);
			{
				/* a constant amplitude (uniform LINE nothing else has read) goes into the
				 * record as a scalar (in the operator word) and its LINE record is dropped */
				const bool ac = in.c < 32 && ((line_uni >> in.c) & 1u);
				uint32_t av = 0;
				if (ac) {
					av = __float_as_uint(uval(in.c));
					if (lane0) sts32(plan + head_rec[in.c] * PLAN_REC, 0u);
					++killed;
					dirty(in.c);
				} else {
					touch(in.c);
				}
				plan_put(plan, n++, P_MIX | (((lstack & 1u) ? PF_LAYER : 0u) | ((in.flags & F_WAVEENV) ? PF_WAVEENV : 0u) |
						(ac ? PF_ACONST : 0u)) << 8 | (uint32_t) in.a << 16 | (uint32_t) in.b << 24,
						(uint32_t) in.c, av, 0u, 0.f, 0.f, 0.f, 0.f);
			}
			dirty(in.a);
			break;
		case I_LINE:
			if (in.d) {
				if (!(kb = line_span(o, in.c, kb))) return 0;
				const uint32_t lf = LM_FLAGS(o->lmeta[in.c]);
				const bool ratio = in.b != NO_BUF && (lf & SAUABI_LINEP_STATE_RATIO);
				const bool u = !(lf & SAUABI_LINEP_GOAL) && (!ratio || is_uni(in.b));
				float f = o->line[in.c].v0;
				if (u && ratio) f = f * uval(in.b);
				if (!u && in.b != NO_BUF) touch(in.b);
				plan_put(plan, n++, P_LINE | (u ? PF_FUNI : 0u) << 8 |
						(uint32_t) in.a << 16 | (uint32_t) in.b << 24, (uint32_t) in.c << 16, opa, 0u,
						0.f, 0.f, f, 0.f);
				dirty(in.a);
				if (u) {           /* dropped only by a RANGE that takes the value as a scalar */
					set_uni(in.a, f); touch(in.a);
					if (in.a < 32) { line_uni |= 1u << in.a; head_rec[in.a] = (uint8_t) (n - 1); }
				}
			}
			break;
		case I_RANGE:
			if (in.a < 32 && in.b < 32 && ((line_uni >> in.a) & 1u) && ((line_uni >> in.b) & 1u)) {
				/* both ends of the range are uniform lines nothing else has read: they go
				 * into the record as scalars and their LINE records have no use left */
				const float pv = uval(in.a), rv = uval(in.b);
				if (lane0) sts32(plan + head_rec[in.a] * PLAN_REC, 0u);
				if (lane0) sts32(plan + head_rec[in.b] * PLAN_REC, 0u);
				killed += 2;
				touch(in.c);
				plan_put(plan, n++, P_RANGE | PF_FUNI << 8 | (uint32_t) in.a << 16 | (uint32_t) in.b << 24, in.c,
						0u, 0u, 0.f, 0.f, pv, rv);
				dirty(in.a); dirty(in.b);
				break;
			}
			plan_put(plan, n++, P_RANGE | (uint32_t) in.a << 16 | (uint32_t) in.b << 24, in.c, 0u, 0u,
					0.f, 0.f, 0.f, 0.f);
			dirty(in.a); touch(in.b); touch(in.c);
			break;
		case I_VOUT:
			plan_put(plan, n++, P_VOUT | (uint32_t) in.a << 16 | (uint32_t) in.b << 24, 0u, opa, 0u,
					0.f, 0.f, 0.f, 0.f);
			return finish(n);
		case I_END:
			return finish(n);
		case I_VPAN:
			if (in.d || (LM_FLAGS(o->lmeta[LINE_PAN]) & SAUABI_LINEP_GOAL)) return 0;
			break;
		/* a wave operator whose amplitude has modulators (run_block_wosc, generator.c:
		 * 548-602, unfused): ENTER [frequency] [PM] PHASOR [amplitude + its modulators]
		 * PMA WOSC MIX LEAVE */
		case I_PHASOR:
			if (in.d != NO_BUF) return 0;                            /* fPM: general path */
			if (o->oscflags & OSC_RESET_DIFF) return 0;
			{
				const bool u = is_uni(in.b);
				const uint32_t inc = u ? uni_inc(in.b) : 0u;
				if (!u) touch(in.b);
				if (in.c != NO_BUF) touch(in.c);
				plan_put(plan, n++, P_PHASE | (u ? PF_FUNI : 0u) << 8 | (uint32_t) in.a << 16 | (uint32_t) in.b << 24,
						(uint32_t) in.c, opa, 0u, 0.f, 0.f, __uint_as_float(inc), 0.f);
				dirty(in.a);
			}
			break;
		case I_PMA: {                                                /* generator.c:485-490 */
			const uint32_t lf = LM_FLAGS(o->lmeta[LINE_PMA]);
			if (o->line[LINE_PMA].v0 != 0.f || (lf & SAUABI_LINEP_GOAL)) {
				/* self-PM: the amount line fills its buffer, the operator's WOSC / RASG
				 * record then runs the serial loop (no self-PM modulator lists here) */
				if (in.op >= 32 || !(kb = line_span(o, LINE_PMA, kb))) return 0;
				plan_put(plan, n++, P_LINE | ((lf & SAUABI_LINEP_GOAL) ? 0u : PF_FUNI) << 8 |
						(uint32_t) in.a << 16 | (uint32_t) NO_BUF << 24, (uint32_t) LINE_PMA << 16, opa, 0u,
						0.f, 0.f, o->line[LINE_PMA].v0, 0.f);
				dirty(in.a);
				selfmask |= 1u << in.op;
			}
			break; }
		case I_WOSC: {
			if ((in.flags & F_HAS_APMODS) || pc + 2 >= code_len) return 0;
			Instr mix, leave;
			memcpy(&mix, &raw_next, sizeof(mix));
			const uint4 raw_leave = __ldg(reinterpret_cast<const uint4*>(code + pc + 2));
			memcpy(&leave, &raw_leave, sizeof(leave));
			if (mix.opcode != I_MIX || mix.b != in.a || leave.opcode != I_LEAVE || leave.op != in.op) return 0;
			const uint32_t wave = o->mode;
			const uint32_t slot = __popc(wave_mask & 0xfffu & ((1u << wave) - 1u));
			const uint32_t ct = (wave_mask & CTAB_FLAG) ? st + slot * CTAB_WAVE_BYTES :
				st + slot * (TAB_STRIDE * 4) + 12;
			if (!depth) return 0;
			const uint32_t fl = ((lstack & 1u) ? PF_LAYER : 0u) |
				((mix.flags & F_WAVEENV) ? PF_WAVEENV : 0u) | PF_ABUF;
			lstack >>= 1;
			--depth;
			if (in.op < 32 && ((selfmask >> in.op) & 1u)) {
				/* sauWOsc_run_selfmod (wosc.h:273-310) on lane 0, then block_mix */
				touch(in.c);
				other = 0x80000000u; plan_put(plan, n++, P_WSELF | fl << 8 | (uint32_t) mix.a << 16 | (uint32_t) in.b << 24,
						(uint32_t) mix.c | (uint32_t) in.c << 8 | (uint32_t) in.a << 16, opa, 0u,
						0.f, 0.f, 0.f, 0.f);
			} else {
				plan_put(plan, n++, P_WOSC | fl << 8 | (uint32_t) mix.a << 16 | (uint32_t) in.b << 24,
						(uint32_t) mix.c, opa, ct, wc->diff_scale[wave], wc->diff_offset[wave], 0.f, 0.f);
			}
			dirty(mix.a); dirty(in.a); touch(in.b); touch(mix.c);
			/* MIX and LEAVE are part of the record */
			pc += 2;
			if (pc + 1 < code_len) raw_next = __ldg(reinterpret_cast<const uint4*>(code + pc + 1));
			break; }
		default:
			return 0;
		}
		bool funi = false, rmul = false;
		float fval = 0.f;          /* the uniform frequency of a HEAD / LEAF */
		if (head) {
			if (in.op >= 32 || (seen & (1u << in.op))) return 0;
			seen |= 1u << in.op;
			if (!(kb = op_span(o, kb)) || !(kb = line_span(o, LINE_FREQ, kb))) return 0;
			const uint32_t lf = LM_FLAGS(o->lmeta[LINE_FREQ]);
			const bool fmul = in.e != NO_BUF && (lf & SAUABI_LINEP_STATE_RATIO);
			funi = !(lf & SAUABI_LINEP_GOAL) && (!fmul || is_uni(in.e));
			fval = o->line[LINE_FREQ].v0;
			if (funi && fmul) fval = fval * uval(in.e);
			if (!funi && in.e != NO_BUF) touch(in.e);
			rmul = !funi && fmul && !(lf & SAUABI_LINEP_GOAL);      /* v0 * parent[k] */
			if (!tail) {
				if (depth >= 31) return 0;
				lstack = (lstack << 1) | ((in.flags & F_LAYER) ? 1u : 0u);   /* popped by its WTAIL / WOSC */
				++depth;
				entered |= 1u << in.op;
				plan_put(plan, n++, P_WHEAD | ((funi ? PF_FUNI : 0u) | (rmul ? PF_FMUL : 0u)) << 8 |
						(uint32_t) in.b << 24, (uint32_t) in.e << 8, opa, 0u,
						0.f, 0.f, fval, 0.f);
				dirty(in.b);
				if (funi && in.b < 32) { set_uni(in.b, fval); head_rec[in.b] = (uint8_t) (n - 1); }
			}
		}
		if (tail) {
			if (!head && in.op < 32 && ((entered >> in.op) & 1u)) {  /* ENTER ... WTAIL: the TAIL is its LEAVE */
				if (!depth) return 0;
				lstack >>= 1;
				--depth;
			}
			if (in.d != NO_BUF) return 0;                           /* fPM: general path */
			if (!(kb = op_span(o, kb)) || !(kb = line_span(o, LINE_AMP, kb))) return 0;
			if (o->oscflags & OSC_RESET_DIFF) return 0;
			if (in.flags & F_MAY_SELFMOD) {                          /* generator.c:485-490 */
				if (o->line[LINE_PMA].v0 != 0.f ||
						(LM_FLAGS(o->lmeta[LINE_PMA]) & SAUABI_LINEP_GOAL)) return 0;
			}
			uint32_t inc = 0;
			if (!head) {
				funi = is_uni(in.b);
				if (funi) inc = uni_inc(in.b);
				else touch(in.b);
			} else if (funi) {
				inc = ftoi_lo32(coeff * fval);
			}
			if (in.c != NO_BUF) touch(in.c);
			const uint32_t wave = o->mode;
			const uint32_t slot = __popc(wave_mask & 0xfffu & ((1u << wave) - 1u));
			const uint32_t ct = (wave_mask & CTAB_FLAG) ? st + slot * CTAB_WAVE_BYTES :
				st + slot * (TAB_STRIDE * 4) + 12;                   /* planes, or &lut[-1] */
			const bool aconst = !(LM_FLAGS(o->lmeta[LINE_AMP]) & SAUABI_LINEP_GOAL);
			const uint32_t fl = ((in.flags & F_LAYER) ? PF_LAYER : 0u) | ((in.flags & F_WAVEENV) ? PF_WAVEENV : 0u) |
				(funi ? PF_FUNI : 0u) | (rmul ? PF_FMUL : 0u) | (aconst ? PF_ACONST : 0u);
			plan_put(plan, n++, (head ? P_WLEAF : P_WTAIL) | fl << 8 | (uint32_t) in.a << 16 | (uint32_t) in.b << 24,
					(uint32_t) in.c | (uint32_t) in.e << 8, opa, ct,
					wc->diff_scale[wave], wc->diff_offset[wave], rmul ? fval : __uint_as_float(inc),
					o->line[LINE_AMP].v0);
			if (!aconst && n < cap) {
				/* the amplitude trajectory's constants, worked out once (sau::line_fill_setup) */
				const LineState &al = o->line[LINE_AMP];
				int t = (int) LM_TYPE(o->lmeta[LINE_AMP]);
				if (t == sau::L_exp) t = (al.v0 > al.vt) ? sau::L_xpe : sau::L_lge;
				else if (t == sau::L_log) t = (al.v0 < al.vt) ? sau::L_xpe : sau::L_lge;
				if (t == sau::L_lin || t == sau::L_xpe || t == sau::L_lge) {
					const float inv = o->linv[LINE_AMP], vd = al.vt - al.v0;
					uint32_t w1 = al.pos;
					float w3, w4;
					if (t == sau::L_lin) {
						w1 = al.pos - (al.end / 2);
						w3 = vd * inv; w4 = (al.v0 + al.vt) * 0.5f;
					} else if (t == sau::L_xpe) {
						w3 = al.v0 - al.vt; w4 = al.vt;
					} else {
						w3 = vd; w4 = al.v0;
					}
					if (lane0) sts32(plan + (n - 1) * PLAN_REC, lds32(plan + (n - 1) * PLAN_REC) | PF_AEXT << 8);
					plan_put(plan, n++, P_EXT | (uint32_t) t << 8, w1, __float_as_uint(inv), __float_as_uint(w3),
							w4, 0.f, 0.f, 0.f);
				}
			}
			dirty(in.a);
		}
	}
	return finish(n);
}

/* nb > 0 steps of line_advance(pos, end, flags, REF_BLOCK) (line.c:385-398) in closed form: the
 * position runs up to `end` in whole blocks, falls back to 0 where it gets there, and goes round
 * again.  Returns bit 0: some step expired, bit 1: the last one did. */
__device__ __forceinline__ uint32_t line_cycle(uint32_t &pos, uint32_t end, uint32_t nb) {
	const uint32_t B = (uint32_t) REF_BLOCK;
	uint32_t k = nb, any = 0;
	if (pos >= end) {                  /* nothing left to advance: expires at once */
		pos = 0;
		if (end == 0 || --k == 0) return 3u;
		any = 1;
	}
	const uint32_t s = (end - pos + B - 1u) / B;       /* steps up to and with the next expiry */
	if (k < s) { pos += k * B; return any; }
	k -= s;
	const uint32_t r = k % ((end + B - 1u) / B);       /* steps into the round the stretch ends in */
	pos = r * B;
	return 1u | (r == 0u ? 2u : 0u);
}
/* sauLine_run's bookkeeping for nb whole blocks of a steady run line */
__device__ __forceinline__ void line_block_update(OpState *o, int li, uint32_t nb) {
	if (!nb) return;
	const uint32_t meta = o->lmeta[li];
	uint32_t flags = LM_FLAGS(meta), pos = o->line[li].pos;
	if (flags & SAUABI_LINEP_GOAL) {
		pos += nb * (uint32_t) REF_BLOCK;
	} else if (line_cycle(pos, o->line[li].end, nb) & 1u) {
		flags &= ~SAUABI_LINEP_TIME;
	}
	o->line[li].pos = pos;
	o->lmeta[li] = LM_PACK(LM_TYPE(meta), flags, 0u);
}
/* nb x line_skip(.., REF_BLOCK) from a block start (sauLine_skip, line.c:449-473): the goal is
 * taken at the first expiry, the position keeps going round */
__device__ __forceinline__ void line_skip_blocks(OpState *o, int li, uint32_t nb) {
	if (!nb) return;
	LineState *ls = &o->line[li];
	const uint32_t meta = o->lmeta[li];
	uint32_t pos = ls->pos, flags = LM_FLAGS(meta);
	const uint32_t r = line_cycle(pos, ls->end, nb);
	if (r & 1u) {
		flags &= ~SAUABI_LINEP_TIME;
		if (flags & SAUABI_LINEP_GOAL) {
			ls->v0 = ls->vt;
			if (flags & SAUABI_LINEP_GOAL_RATIO) flags |= SAUABI_LINEP_STATE_RATIO;
			else flags &= ~SAUABI_LINEP_STATE_RATIO;
			flags &= ~(SAUABI_LINEP_GOAL | SAUABI_LINEP_GOAL_RATIO);
		}
	}
	ls->pos = pos;
	o->lmeta[li] = LM_PACK(LM_TYPE(meta), flags, r >> 1);
}

/* every lane: the instructions are shared out (each touches its own lines / time word of its operator: an
 * operator's frequency, amplitude and other lines, its time and its self-PM flag each belong to ONE instruction
 * of a voice program, runtime.cpp:Compiler) */
__device__ __noinline__ void steady_update(OpState *sops, const Instr *code, uint32_t code_len, uint32_t nb, int lane) {
	for (uint32_t pc = lane; pc < code_len; pc += 32) {
		const uint4 raw = __ldg(reinterpret_cast<const uint4*>(code + pc));
		Instr in;
		memcpy(&in, &raw, sizeof(in));
		OpState *o = sops + in.op;
		bool head = false, tail = false;
		switch (in.opcode) {
		case I_WLEAF: head = tail = true; break;
		case I_WHEAD: head = true; break;
		case I_WTAIL: tail = true; break;
		case I_LINE:
			if (in.d) line_block_update(o, in.c, nb);
			else line_skip_blocks(o, in.c, nb);
			break;
		case I_VPAN:
			line_skip_blocks(o, LINE_PAN, nb);
			break;
		case I_PMA:                /* as pma_decide / run_osc_selfmod_param, generator.c:485-490 */
			if (o->line[LINE_PMA].v0 != 0.f || (LM_FLAGS(o->lmeta[LINE_PMA]) & SAUABI_LINEP_GOAL)) {
				line_block_update(o, LINE_PMA, nb);
				o->flags |= ON_PMA_RUN;
			} else {
				line_skip_blocks(o, LINE_PMA, nb);
				o->flags &= ~ON_PMA_RUN;
			}
			break;
		case I_LEAVE:              /* unfused wave operator, generator.c:726-727 */
			if (!(o->flags & ON_TIME_INF)) o->time -= nb * (uint32_t) REF_BLOCK;
			break;
		default: break;
		}
		if (head) {
			line_block_update(o, LINE_FREQ, nb);
			if (in.flags & F_SKIP_FREQ2) line_skip_blocks(o, LINE_FREQ2, nb);
		}
		if (tail) {
			line_block_update(o, LINE_AMP, nb);
			if (in.flags & F_SKIP_AMP2) line_skip_blocks(o, LINE_AMP2, nb);
			if (in.flags & F_MAY_SELFMOD) {
				line_skip_blocks(o, LINE_PMA, nb);
				o->flags &= ~ON_PMA_RUN;
			}
			if (!(o->flags & ON_TIME_INF)) o->time -= nb * (uint32_t) REF_BLOCK;   /* generator.c:726-727 */
		}
	}
}

/* the trajectory of a steady goal line at positions pos .. pos+NS-1 (line.c:27-281);
 * out of line: one copy of the 11 shapes for all call sites */
template <int TYPE, int NS>
__device__ __forceinline__ void line_fillN(const sau::LineFill &f, float out[NS]) {
	sau::LineFill g = f;
	g.type = TYPE;
#pragma unroll
	for (int k = 0; k < NS; ++k) out[k] = sau::line_fill_at(g, (uint32_t) k, false);
}
template <int NS> struct LineVec { float v[NS]; };
template <int NS>
__device__ __noinline__ LineVec<NS> line_goal_fill(float v0, float vt, float inv, uint32_t pos,
		uint32_t end, uint32_t type) {
	sau::LineFill f;
	int t = (int) type;
	if (t == sau::L_exp) t = (v0 > vt) ? sau::L_xpe : sau::L_lge;
	else if (t == sau::L_log) t = (v0 < vt) ? sau::L_xpe : sau::L_lge;
	f.type = t;
	f.v0 = v0; f.vt = vt;
	f.pos = pos;
	f.adj_pos = (int32_t) (pos - (end / 2));
	f.inv = inv;
	f.vm = (v0 + vt) * 0.5f;
	f.vd = vt - v0;
	f.c = 0.f;
	LineVec<NS> r;
	float *out = r.v;
	switch (t) {
	default:
	case sau::L_sah: line_fillN<sau::L_sah, NS>(f, out); break;
	case sau::L_lin: f.c = f.vd * f.inv; line_fillN<sau::L_lin, NS>(f, out); break;
	case sau::L_cos: line_fillN<sau::L_cos, NS>(f, out); break;
	case sau::L_xpe: f.c = v0 - vt; line_fillN<sau::L_xpe, NS>(f, out); break;
	case sau::L_lge: line_fillN<sau::L_lge, NS>(f, out); break;
	case sau::L_sqe: f.c = v0 - vt; line_fillN<sau::L_sqe, NS>(f, out); break;
	case sau::L_cub: f.inv = -2.f * f.inv; f.c = (v0 - vt) * 0.5f; line_fillN<sau::L_cub, NS>(f, out); break;
	case sau::L_smo: line_fillN<sau::L_smo, NS>(f, out); break;
	case sau::L_uwh: f.c = f.vd * (0.5f / 2147483648.f); line_fillN<sau::L_uwh, NS>(f, out); break;
	case sau::L_ncl: line_fillN<sau::L_ncl, NS>(f, out); break;
	case sau::L_nhl: line_fillN<sau::L_nhl, NS>(f, out); break;
	}
	return r;
}

/* The per-chunk code addresses shared memory by 32-bit shared-window addresses
 * through ld.shared / st.shared: one register per base, no generic loads, no
 * re-derivation of the bases.  NS = samples per lane (chunk = 32 * NS).
 * Work buffer i of the fast path: NS/4 planes of 32 float4 (lane-major, so
 * 128-bit accesses are conflict-free), FBUF_BYTES apart. */
template <int NS> struct FastCfg {
	static constexpr uint32_t CHUNKF = 32 * NS;
	static constexpr uint32_t FBUF_BYTES = CHUNKF * 4;
};
struct FastCtx {               /* all registers */
	uint32_t sb;               // shared addr of this lane's float4 in plane 0 of buffer 0
	uint32_t so;               // shared addr of the operator states
	uint32_t st;               // shared addr of the staged tables
	uint32_t wave_mask;
	uint32_t oc;               // chunk offset inside the block
	int lane;
	float coeff, amp_scale;
	uint32_t write_r;          // as Ctx::write_r
	uint32_t plan, plan_cap;   // shared addr of the block plan, records it can hold
	const WaveCoeffs *wc;
	const float *tab;          // generic pointer to the staged tables (rare paths)
	const struct TeamCtx *team;   // the voice's team of warps (render_team.cuh), or null
	bool keep_plans;           // one warp renders all of a voice's units: its plan is kept (GenDesc::plan_cache)
};
/* What the chunk loop of a steady block keeps in registers; everything else it
 * needs is in the block plan (shared memory): header at c.plan, records after it. */
struct HotCtx {
	uint32_t sb;               // as FastCtx::sb
	uint32_t plan;             // shared addr of the plan header
	uint32_t oc;               // chunk offset inside the block
	int lane;
	float coeff;
};
/* plan header (PLAN_HDR bytes): the cold paths' context, VOUT's constants, the voice's rows and
 * the frame of the stretch's first sample */
/* slot 1 is the voice output's: what every chunk needs in ONE 128-bit load (the s row, the tile stride, the frame
 * of the stretch's first sample with "the r row is written" in bit 31) + one 64-bit load (amp_scale, the static
 * pan), then the r row */
constexpr uint32_t PH_TAB = 0, PH_WC = 8, PH_WAVE_MASK = 16, PH_COEFF = 20,
	PH_ROW_S = 32, PH_TSTRIDE = 40, PH_FRAME0 = 44, PH_AMP_SCALE = 48, PH_PAN = 52, PH_ROW_R = 56;
template <int NS>
__device__ __forceinline__ void fld(const HotCtx &c, uint32_t buf, float v[NS]) {
	const uint32_t a = c.sb + buf * FastCfg<NS>::FBUF_BYTES;
#pragma unroll
	for (int h = 0; h < NS / 4; ++h) {
		const float4 t = lds128(a + h * 512);
		v[4 * h] = t.x; v[4 * h + 1] = t.y; v[4 * h + 2] = t.z; v[4 * h + 3] = t.w;
	}
}
template <int NS>
__device__ __forceinline__ void fst(const HotCtx &c, uint32_t buf, const float v[NS]) {
	const uint32_t a = c.sb + buf * FastCfg<NS>::FBUF_BYTES;
#pragma unroll
	for (int h = 0; h < NS / 4; ++h)
		sts128(a + h * 512, make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]));
}

/* byte offsets inside OpState (device_types.h) */
constexpr uint32_t OS_LINE = 0, OS_LMETA = 96, OS_LINV = 120, OS_TIME = 144, OS_PREVS = 152,
	OS_I0 = 160, OS_I1 = 164, OS_PREV = 168;
static_assert(offsetof(OpState, lmeta) == OS_LMETA && offsetof(OpState, linv) == OS_LINV &&
		offsetof(OpState, time) == OS_TIME && offsetof(OpState, i0) == OS_I0 &&
		offsetof(OpState, i1) == OS_I1 && offsetof(OpState, prev_Is) == OS_PREV &&
		offsetof(OpState, prev_s) == OS_PREVS, "OpState offsets");

/* value of a steady run line for this lane's samples of the chunk at c.oc */
template <int NS>
__device__ __forceinline__ void line_value_steady(const HotCtx &c, uint32_t op, int li,
		const float *m /* NS multipliers or nullptr */, float out[NS]) {
	const uint4 core = lds128u(op + OS_LINE + 16 * li);          /* v0, vt, pos, end */
	const uint32_t meta = lds32(op + OS_LMETA + 4 * li);
	const float v0 = __uint_as_float(core.x);
	const uint32_t flags = LM_FLAGS(meta);
	if (!(flags & SAUABI_LINEP_GOAL)) {
		if (m && (flags & SAUABI_LINEP_STATE_RATIO)) {
#pragma unroll
			for (int k = 0; k < NS; ++k) out[k] = v0 * m[k];
		} else {
#pragma unroll
			for (int k = 0; k < NS; ++k) out[k] = v0;
		}
		return;
	}
	const float inv = lds32f(op + OS_LINV + 4 * li);
	{
		const float vt = __uint_as_float(core.y);
		const uint32_t pos = core.z + c.oc + c.lane * NS;
		int t = (int) LM_TYPE(meta);
		if (t == sau::L_exp) t = (v0 > vt) ? sau::L_xpe : sau::L_lge;
		else if (t == sau::L_log) t = (v0 < vt) ? sau::L_xpe : sau::L_lge;
		if (t == sau::L_lin || t == sau::L_xpe || t == sau::L_lge) {
			/* the usual envelope shapes stay in line (no call, no stack traffic);
			 * same set-up as line_goal_fill / sau::line_fill_setup */
			sau::LineFill f;
			f.type = t;
			f.v0 = v0; f.vt = vt; f.pos = pos;
			f.adj_pos = (int32_t) (pos - (core.w / 2));
			f.inv = inv;
			f.vm = (v0 + vt) * 0.5f;
			f.vd = vt - v0;
			f.c = 0.f;
			if (t == sau::L_lin) { f.c = f.vd * f.inv; line_fillN<sau::L_lin, NS>(f, out); }
			else if (t == sau::L_xpe) { f.c = v0 - vt; line_fillN<sau::L_xpe, NS>(f, out); }
			else line_fillN<sau::L_lge, NS>(f, out);
		} else {
			const LineVec<NS> r = line_goal_fill<NS>(v0, vt, inv, pos, core.w, LM_TYPE(meta));
#pragma unroll
			for (int k = 0; k < NS; ++k) out[k] = r.v[k];
		}
	}
	if (m && (flags & SAUABI_LINEP_GOAL_RATIO)) {
#pragma unroll
		for (int k = 0; k < NS; ++k) out[k] = out[k] * m[k];
	}
}

/* sauWOsc_run over a full chunk when some phase difference is zero (the output
 * then repeats, wosc.h:251-252): same scheme as wosc_eval_any, NS samples per
 * lane, by value. */
template <int NS> struct PhaseVec { uint32_t v[NS]; };
template <int NS> struct SampVec { float v[NS]; };
template <int NS>
__device__ __noinline__ SampVec<NS> wosc_zero_diff(const ColdCtx c, OpState *o, const PhaseVec<NS> phv) {
	const uint32_t *ph = phv.v;
	SampVec<NS> sv;
	float *s = sv.v;
	const uint32_t wave = o->mode;
	const WaveRef lut = wave_ref(c, wave);
	const float ds = c.wc->diff_scale[wave], doff = c.wc->diff_offset[wave];
	const uint32_t prev_phase = o->i1;
	const double prev_Is = o->prev_Is;
	const float prev_s = o->prev_s;
	double Is[NS];
#pragma unroll
	for (int k = 0; k < NS; ++k) Is[k] = herp_ref(lut, ph[k], (double*) 0, (double*) 0);
	uint32_t pph = __shfl_up_sync(FULL, ph[NS - 1], 1);
	double pIs = __shfl_up_sync(FULL, Is[NS - 1], 1);
	if (c.lane == 0) { pph = prev_phase; pIs = prev_Is; }
	bool zd[NS];
	bool lead_zero = false, has_nz = false;
	float s_run = 0.f;
#pragma unroll
	for (int k = 0; k < NS; ++k) {
		const int32_t d = (int32_t) (ph[k] - pph);
		zd[k] = d == 0;
		if (d != 0) {
			s_run = sau::wosc_diff(Is[k], pIs, d, ds, doff);
			has_nz = true;
		}
		if (zd[k] && !has_nz) lead_zero = true;
		s[k] = s_run;
		pph = ph[k]; pIs = Is[k];
	}
	const uint32_t any_lead = __ballot_sync(FULL, lead_zero);
	if (any_lead) {
		const uint32_t nzmask = __ballot_sync(FULL, has_nz);
		const uint32_t lower = nzmask & ((1u << c.lane) - 1u);
		const int src = lower ? (31 - __clz(lower)) : 0;
		float inc = __shfl_sync(FULL, s_run, src);
		if (!lower) inc = prev_s;
		bool seen = false;
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			if (!zd[k]) seen = true;
			if (!seen) s[k] = inc;
		}
	}
	__syncwarp();
	if (c.lane == 31) { o->i1 = ph[NS - 1]; o->prev_Is = Is[NS - 1]; o->prev_s = s[NS - 1]; }
	return sv;
}

/* the phase fraction as a double, (double) ((float) frac * 2^-21) of sauWave_get_herp
 * (wave.h:131-133; both steps are exact, frac < 2^21): frac dropped into the low
 * mantissa bits of 2^31, whose unit in the last place is 2^-21, minus 2^31 --
 * one FP64 add instead of I2F + FMUL + F2F on the quarter-rate conversion pipe */
__device__ __forceinline__ double phase_frac(uint32_t phase) {
	return __hiloint2double(0x41E00000, (int) (phase & sau::WAVE_SLENMASK)) - 2147483648.0;
}
__device__ __forceinline__ double horner_frac(double c3, double c2, double c1, uint32_t phase) {
	const double x = phase_frac(phase);
	return ((c3 * x + c2) * x + c1) * x;
}

/* Phase fill of a wave operator on a steady full chunk (sauPhasor_fill,
 * wosc.h:135-169).  funi: every sample adds `inc` to the phase (see steady_plan);
 * else fr = its frequency values.  bufc: PM input or NO_BUF. */
template <int NS>
__device__ __forceinline__ void phase_plan(const HotCtx &c, const uint32_t op, const uint32_t bufc,
		const bool funi, const uint32_t inc, const float fr[NS], uint32_t ph[NS]) {
	uint2 og;                                    /* i0, i1 (phase, prev_phase) */
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(og.x), "=r"(og.y) : "r"(op + OS_I0));
	__syncwarp();              /* every lane holds the accumulator before lane 31 rewrites it */
	if (funi) {
		const uint32_t base = og.x + inc * (uint32_t) (c.lane * NS);
#pragma unroll
		for (int k = 0; k < NS; ++k) ph[k] = base + inc * (uint32_t) (k + 1);
		if (c.lane == 31) sts32(op + OS_I0, ph[NS - 1]);
	} else {
		uint32_t run = 0;
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			run += ftoi_lo32(c.coeff * fr[k]);
			ph[k] = run;
		}
		const uint32_t incl = scan_incl_u32(run, c.lane);
		const uint32_t base = og.x + (incl - run);
#pragma unroll
		for (int k = 0; k < NS; ++k) ph[k] += base;
		if (c.lane == 31) sts32(op + OS_I0, og.x + incl);
	}
	if (bufc != NO_BUF) {      /* PM; fPM operators take the general path (steady_plan) */
		float pm[NS];
		fld<NS>(c, bufc, pm);
#pragma unroll
		for (int k = 0; k < NS; ++k) ph[k] += ftoi_lo32(pm[k] * 2147483648.f);
	}
}

/* Oscillator, amplitude and block_mix of a wave operator on a steady full chunk
 * (sauWOsc_run, wosc.h:238-266; generator.c:584-601) at the phases ph.  (pure:
 * every phase difference is `inc`; one division instead of four was measured
 * SLOWER than four in one block with the table evaluation.)  The amplitude is
 * the operator's own line, or (PF_ABUF) a buffer its modulators wrote. */
template <int NS, bool CTAB>
__device__ __forceinline__ void osc_plan(const HotCtx &c, const uint4 p0, const uint32_t rec,
		const bool pure, const uint32_t inc, const uint32_t ph[NS]) {
	const uint32_t flags = (p0.x >> 8) & 0xffu, bufa = (p0.x >> 16) & 0xffu;
	const uint32_t op = p0.z;
	uint2 pg;                                    /* prev_Is lo / hi */
	uint32_t pph0;                               /* prev_phase */
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(pg.x), "=r"(pg.y) : "r"(op + OS_PREV));
	pph0 = lds32(op + OS_I1);
	__syncwarp();              /* every lane holds the carried values before lane 31 rewrites them */
	float s[NS];
	{
		/* the phase differences and their quotients diff_scale / d first (FP32, from the
		 * phases alone; a zero difference gives a quotient nobody uses): in one block
		 * with the table evaluation below, they fill the FP64 pipe's latency */
		uint32_t pph = __shfl_up_sync(FULL, ph[NS - 1], 1);
		if (c.lane == 0) pph = pph0;
		int32_t d[NS];
		d[0] = (int32_t) (ph[0] - pph);
#pragma unroll
		for (int k = 1; k < NS; ++k) d[k] = (int32_t) (ph[k] - ph[k - 1]);
		const float2 dd = lds64f(rec + 16);
		const float ds = dd.x;
		float xq[NS];
#pragma unroll
		for (int k = 0; k < NS; ++k) xq[k] = div_scale_by_int(ds, d[k]);   /* wosc.h:254-256 */
		double Is[NS];
		if (CTAB) {
			/* per-index coefficients from shared memory: two loads, Horner */
#pragma unroll
			for (int k = 0; k < NS; ++k) {
				const uint32_t ind = ph[k] >> sau::WAVE_SLENBITS;
				const double2 hi = lds128d(p0.w + (ind << 4));
				const float2 lo = lds64f(p0.w + CTAB_PLANE_BYTES + (ind << 3));
				Is[k] = horner_frac(hi.x, hi.y, (double) lo.x, ph[k]) + (double) lo.y;
			}
		} else {
#pragma unroll
			for (int k = 0; k < NS; ++k) {
				const uint32_t a = p0.w + ((ph[k] >> sau::WAVE_SLENBITS) << 2);
				const float s0 = lds32f(a), s1 = lds32f(a + 4), s2 = lds32f(a + 8), s3 = lds32f(a + 12);
				double c1, c2, c3;
				sau::herp_coefs(s0, s1, s2, s3, &c1, &c2, &c3);
				Is[k] = horner_frac(c3, c2, c1, ph[k]) + (double) s1;
			}
		}
		double pIs = __shfl_up_sync(FULL, Is[NS - 1], 1);
		if (c.lane == 0) pIs = __hiloint2double((int) pg.y, (int) pg.x);
		bool z = false;
#pragma unroll
		for (int k = 0; k < NS; ++k) z |= (d[k] == 0);
		if (__any_sync(FULL, z)) {
			const uint4 h = lds128u(c.plan);
			ColdCtx cc;
			cc.tab = reinterpret_cast<const float*>((uint64_t) h.x | ((uint64_t) h.y << 32));
			cc.wc = reinterpret_cast<const WaveCoeffs*>((uint64_t) h.z | ((uint64_t) h.w << 32));
			cc.wave_mask = lds32(c.plan + PH_WAVE_MASK); cc.lane = c.lane;
			PhaseVec<NS> pv;
#pragma unroll
			for (int k = 0; k < NS; ++k) pv.v[k] = ph[k];
			OpState *o = reinterpret_cast<OpState*>(__cvta_shared_to_generic(op));
			const SampVec<NS> sv = wosc_zero_diff<NS>(cc, o, pv);
#pragma unroll
			for (int k = 0; k < NS; ++k) s[k] = sv.v[k];
		} else {
			const double doff = (double) dd.y;
#pragma unroll
			for (int k = 0; k < NS; ++k) {
				const double dI = Is[k] - (k ? Is[k - 1] : pIs);
				s[k] = (float) (dI * (double) xq[k] + doff);
			}
			if (c.lane == 31) {
				sts32(op + OS_I1, ph[NS - 1]);
				asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(op + OS_PREV),
						"r"((uint32_t) __double2loint(Is[NS - 1])), "r"((uint32_t) __double2hiint(Is[NS - 1])) : "memory");
				sts32(op + OS_PREVS, __float_as_uint(s[NS - 1]));
			}
		}
	}
	float am[NS];
	if (flags & PF_ACONST) {
		const float av = lds32f(rec + 28);
#pragma unroll
		for (int k = 0; k < NS; ++k) am[k] = av;
	} else if (flags & PF_ABUF) {
		fld<NS>(c, p0.y & 0xffu, am);
	} else if (flags & PF_AEXT) {
		/* sauLine_fill_lin / _xpe / _lge (line.c:65-140, as sau::line_fill_at) from the
		 * stretch constants in the record's second slot */
		const uint4 e = lds128u(rec + PLAN_REC);
		const float w4 = lds32f(rec + PLAN_REC + 16);
		const float inv = __uint_as_float(e.z), w3 = __uint_as_float(e.w);
		const uint32_t i0 = e.y + c.oc + (uint32_t) (c.lane * NS);
		const uint32_t t = e.x >> 8;
		if (t == (uint32_t) sau::L_lin) {
#pragma unroll
			for (int k = 0; k < NS; ++k) am[k] = (sau::i2f((int32_t) (i0 + k)) * w3) + w4;
		} else if (t == (uint32_t) sau::L_xpe) {
#pragma unroll
			for (int k = 0; k < NS; ++k) am[k] = sau::expramp6(1.f - sau::u2f(i0 + k) * inv) * w3 + w4;
		} else {
#pragma unroll
			for (int k = 0; k < NS; ++k) am[k] = sau::expramp6(sau::u2f(i0 + k) * inv) * w3 + w4;
		}
	} else {
		line_value_steady<NS>(c, op, LINE_AMP, nullptr, am);
	}
	const bool layer = (flags & PF_LAYER) != 0;        /* F_LAYER_PMA: no self-PM here */
	float ov[NS];
	if (layer) fld<NS>(c, bufa, ov);
	if (flags & PF_WAVEENV) {                                     /* generator.c:407-426 */
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			const float s_amp = am[k] * 0.5f;
			const float v = (s[k] * s_amp) + fabsf(s_amp);
			ov[k] = layer ? ov[k] * v : v;
		}
	} else {                                                      /* generator.c:384-397 */
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			const float v = s[k] * am[k];
			ov[k] = layer ? ov[k] + v : v;
		}
	}
	fst<NS>(c, bufa, ov);
	__syncwarp();
}

/* rows not 16-byte aligned: scalar stores from the (plane-major) fast buffers */
__device__ __noinline__ void vout_unaligned(uint32_t sbuf_s, uint32_t sbuf_r, float *row_s, float *row_r,
		int lane, int ns, uint32_t write_r, uint32_t frame, uint32_t tstride) {
	for (int k = 0; k < ns; ++k) {
		/* sample lane*ns + k sits in plane k/4, float4 slot `lane`, component k%4 */
		const uint32_t off = (uint32_t) (k >> 2) * 512u + (uint32_t) lane * 16u + (uint32_t) (k & 3) * 4u;
		const size_t at = row_index(frame + (uint32_t) (lane * ns + k), tstride);
		row_s[at] = lds32f(sbuf_s + off);
		if (write_r) row_r[at] = lds32f(sbuf_r + off);
	}
}

/* The feed-forward operator types other than wave oscillators on a steady full chunk
 * (noise, rumble without self-PM, DC / mix): the general interpreter's own routines
 * (same buffer layout, FAST_NS == SPL), out of line, on a minimal context. */
static_assert(FAST_NS == SPL, "plan_ff / plan_other run the general routines on the fast buffers");
__device__ __noinline__ void plan_ff(uint32_t kind, uint32_t sb0, int lane, float coeff, uint32_t oc,
		uint32_t op, uint32_t w0, uint32_t w1) {
	Ctx c;
	c.bufs = reinterpret_cast<float*>(__cvta_shared_to_generic(sb0));
	c.sops = reinterpret_cast<OpState*>(__cvta_shared_to_generic(op));
	c.lane = lane; c.coeff = coeff; c.oc = oc % (uint32_t) REF_BLOCK; c.pma_flag = false; c.sp = 0;
	Instr in;
	in.opcode = 0; in.op = 0; in.flags = 0; in.aux = 0;
	in.a = (uint8_t) (w0 >> 16); in.b = (uint8_t) (w0 >> 24);
	in.c = (uint8_t) w1; in.d = (uint8_t) (w1 >> 8); in.e = (uint8_t) NO_BUF;
	if (kind == P_MIX) {                                           /* block_mix_*, generator.c:384-440 */
		float x[SPL] = {1.f, 1.f, 1.f, 1.f}, a[SPL];
		if (in.b != NO_BUF) ld4(c, in.b, x);
		if ((w0 >> 8) & PF_ACONST) {               /* `op` carries the constant amplitude */
#pragma unroll
			for (int k = 0; k < SPL; ++k) a[k] = __uint_as_float(op);
		} else {
			ld4(c, in.c, a);
		}
		mix_eval<true>(c, in.a, x, a, CHUNK, (w0 >> 8) & PF_LAYER, ((w0 >> 8) & PF_WAVEENV) != 0);
	}
	else if (kind == P_NOISE) noise_run(c, in, CHUNK);             /* sauNoiseG_run_*, noise.h:41-185 */
	else if (kind == P_CYCLE) cyclor_fill(c, in, CHUNK);           /* sauCyclor_fill, rasg.h:165-222 */
	else rasg_run(c, in, CHUNK, REF_BLOCK);                        /* sauRasG_run, rasg.h:692-743 */
	__syncwarp();
}

/* The same for plans with serial self-PM records (P_WSELF, self-PM P_RASG): they also
 * need the plan header; such plans run in their own instance of the chunk loop. */
__device__ __noinline__ void plan_other(uint32_t sb, float coeff, uint32_t oc, uint32_t rec, uint32_t plan) {
	/* few arguments: the call sits in the hot loop's register allocation */
	const int lane = (int) (threadIdx.x & 31u);
	const uint32_t sb0 = sb - (uint32_t) lane * 16u;
	const uint4 p0 = lds128u(rec);
	const uint32_t w0 = p0.x, w1 = p0.y, op = p0.z, kind = w0 & 0xffu;
	Ctx c;
	c.bufs = reinterpret_cast<float*>(__cvta_shared_to_generic(sb0));
	c.sops = reinterpret_cast<OpState*>(__cvta_shared_to_generic(op));
	c.lane = lane; c.coeff = coeff; c.oc = oc % (uint32_t) REF_BLOCK; c.sp = 0;
	c.pma_flag = kind == P_RASG && ((w0 >> 8) & PF_FUNI);          /* self-PM rumble */
	Instr in;
	in.opcode = 0; in.op = 0; in.flags = 0; in.aux = 0;
	in.a = (uint8_t) (w0 >> 16); in.b = (uint8_t) (w0 >> 24);
	in.c = (uint8_t) w1; in.d = (uint8_t) (w1 >> 8); in.e = (uint8_t) NO_BUF;
	if (kind == P_WSELF) {                                         /* sauWOsc_run_selfmod + block_mix */
		const uint4 h = lds128u(plan);
		ColdCtx cc;
		cc.tab = reinterpret_cast<const float*>((uint64_t) h.x | ((uint64_t) h.y << 32));
		cc.wc = reinterpret_cast<const WaveCoeffs*>((uint64_t) h.z | ((uint64_t) h.w << 32));
		cc.wave_mask = lds32(plan + PH_WAVE_MASK); cc.lane = lane;
		const uint32_t amp_buf = w1 & 0xffu, pma_buf = (w1 >> 8) & 0xffu, dst = (w1 >> 16) & 0xffu;
		wosc_selfmod(cc, c.sops, reinterpret_cast<const uint32_t*>(c.bufs + in.b * CHUNK),
				c.bufs + pma_buf * CHUNK, c.bufs + dst * CHUNK, CHUNK);
		float x[SPL], a[SPL];
		ld4(c, dst, x);
		ld4(c, amp_buf, a);
		mix_eval<true>(c, in.a, x, a, CHUNK, (w0 >> 8) & PF_LAYER, ((w0 >> 8) & PF_WAVEENV) != 0);
	}
	else if (kind == P_MIX) {                                      /* block_mix_*, generator.c:384-440 */
		float x[SPL] = {1.f, 1.f, 1.f, 1.f}, a[SPL];
		if (in.b != NO_BUF) ld4(c, in.b, x);
		if ((w0 >> 8) & PF_ACONST) {               /* `op` carries the constant amplitude */
#pragma unroll
			for (int k = 0; k < SPL; ++k) a[k] = __uint_as_float(op);
		} else {
			ld4(c, in.c, a);
		}
		mix_eval<true>(c, in.a, x, a, CHUNK, (w0 >> 8) & PF_LAYER, ((w0 >> 8) & PF_WAVEENV) != 0);
	}
	else if (kind == P_NOISE) noise_run(c, in, CHUNK);             /* sauNoiseG_run_*, noise.h:41-185 */
	else if (kind == P_CYCLE) cyclor_fill(c, in, CHUNK);           /* sauCyclor_fill, rasg.h:165-222 */
	else rasg_run(c, in, CHUNK, REF_BLOCK);                        /* sauRasG_run, rasg.h:692-743 */
	__syncwarp();
}

/* One record of the plan in its general form (anything steady_plan emits).  May consume the
 * record's second slot (rec advances).  Returns true after the voice output (the chunk is done). */
template <int NS, bool CTAB, bool OTHER>
__device__ __forceinline__ bool plan_record_generic(const HotCtx &c, uint32_t &rec, const uint4 p0,
		float *row_s, float *row_r, const uint32_t frame) {
	const uint32_t kind = p0.x & 0xffu, flags = (p0.x >> 8) & 0xffu;
	const uint32_t bufa = (p0.x >> 16) & 0xffu, bufb = p0.x >> 24;
	const uint32_t op = p0.z;
	if (kind <= P_WOSC) {
		/* LINE / WHEAD: one line evaluation into a buffer; WLEAF: the same, kept in
		 * registers, then phase fill and oscillator; WTAIL: the frequency comes from
		 * its buffer; PHASE / WOSC: the two halves of an operator whose amplitude has
		 * modulators, with the phases parked in a buffer in between */
		uint32_t ph[NS];
		uint32_t inc = 0;
		bool pure = false;
		if (kind != P_WOSC) {
			const bool is_line = kind == P_LINE;
			const bool funi = (flags & PF_FUNI) != 0;
			float fr[NS];
			if (kind == P_WTAIL || kind == P_PHASE) {
				if (funi) inc = lds32(rec + 24);
				else fld<NS>(c, bufb, fr);
			} else {
				const uint32_t mb = is_line ? bufb : (p0.y >> 8) & 0xffu;
				if (funi) {
					const uint32_t w6 = lds32(rec + 24);
					inc = w6;
#pragma unroll
					for (int k = 0; k < NS; ++k) fr[k] = __uint_as_float(w6);
				} else if (!is_line && (flags & PF_FMUL)) {
					/* a constant ratio to a modulated parent frequency (line.c:417-445, no goal) */
					const float v0 = lds32f(rec + 24);
					fld<NS>(c, mb, fr);
#pragma unroll
					for (int k = 0; k < NS; ++k) fr[k] = v0 * fr[k];
				} else {
					float m[NS];
					const bool has_mul = mb != NO_BUF;
					if (has_mul) fld<NS>(c, mb, m);
					line_value_steady<NS>(c, op, is_line ? (int) ((p0.y >> 16) & 0xffu) : (int) LINE_FREQ,
							has_mul ? m : nullptr, fr);
				}
				if (kind != P_WLEAF) { fst<NS>(c, is_line ? bufa : bufb, fr); return false; }
			}
			const uint32_t bufc = p0.y & 0xffu;
			phase_plan<NS>(c, op, bufc, funi, inc, fr, ph);
			if (kind == P_PHASE) {
				float pf[NS];
#pragma unroll
				for (int k = 0; k < NS; ++k) pf[k] = __uint_as_float(ph[k]);
				fst<NS>(c, bufa, pf);
				return false;
			}
			pure = funi && bufc == NO_BUF;
		} else {
			float pf[NS];
			fld<NS>(c, bufb, pf);
#pragma unroll
			for (int k = 0; k < NS; ++k) ph[k] = __float_as_uint(pf[k]);
		}
		osc_plan<NS, CTAB>(c, p0, rec, pure, inc, ph);
		if (flags & PF_AEXT) rec += PLAN_REC;                  /* the record's second slot */
	} else if (kind == P_RANGE) {                              /* generator.c:465-467 */
		float p[NS], rr[NS], m[NS];
		fld<NS>(c, p0.y & 0xffu, m);
		if (flags & PF_FUNI) {             /* both ends uniform: scalars from the record */
			const float2 pr = lds64f(rec + 24);
#pragma unroll
			for (int k = 0; k < NS; ++k) { p[k] = pr.x; rr[k] = pr.y; }
		} else {
			fld<NS>(c, bufa, p); fld<NS>(c, bufb, rr);
		}
#pragma unroll
		for (int k = 0; k < NS; ++k) p[k] += (rr[k] - p[k]) * m[k];
		fst<NS>(c, bufa, p);
	} else if (kind != P_VOUT) {                               /* P_NOISE, P_CYCLE, P_RASG, P_MIX, P_WSELF */
		if (OTHER) plan_other(c.sb, c.coeff, c.oc, rec, c.plan);
		else plan_ff(kind, c.sb - c.lane * 16, c.lane, c.coeff, c.oc, op, p0.x, p0.y);
	} else {                                                   /* P_VOUT, generator.c:772-786 */
		float sv[NS];
		fld<NS>(c, bufa, sv);
		const float pan = lds32f(op + OS_LINE + 16 * LINE_PAN);
		float s[NS], rv[NS];
		const float amp_scale = lds32f(c.plan + PH_AMP_SCALE);
		const uint32_t write_r = lds32(c.plan + PH_FRAME0) >> 31;
		const uint32_t tstride = lds32(c.plan + PH_TSTRIDE);
#pragma unroll
		for (int k = 0; k < NS; ++k) { s[k] = sv[k] * amp_scale; rv[k] = s[k] * pan; }
		/* row_s / row_r: this voice's piece of frame tile 0 (device_types.h:ROW_TILE) */
		const uint32_t fl = frame + c.lane * NS;
		if ((frame & 3u) == 0) {
#pragma unroll
			for (int h = 0; h < NS / 4; ++h) {                   /* 128-bit streaming stores */
				const size_t at = row_index(fl + 4 * h, tstride);
				__stcs(reinterpret_cast<float4*>(row_s + at),
						make_float4(s[4 * h], s[4 * h + 1], s[4 * h + 2], s[4 * h + 3]));
				if (write_r)
					__stcs(reinterpret_cast<float4*>(row_r + at),
							make_float4(rv[4 * h], rv[4 * h + 1], rv[4 * h + 2], rv[4 * h + 3]));
			}
		} else {
			/* segment starting at an odd frame: rare, out of line through the buffers */
			const uint32_t rb = bufb != NO_BUF ? bufb : bufa + 1u;
			fst<NS>(c, bufa, s);
			fst<NS>(c, rb, rv);
			__syncwarp();
			vout_unaligned(c.sb - c.lane * 16 + bufa * FastCfg<NS>::FBUF_BYTES,
					c.sb - c.lane * 16 + rb * FastCfg<NS>::FBUF_BYTES,
					row_s, row_r, c.lane, NS, write_r, frame, tstride);
		}
		return true;
	}
	return false;
}

template <int NS, bool CTAB, bool OTHER>
__device__ __forceinline__ void run_chunk_plan(const HotCtx &c,
		float *row_s, float *row_r, const uint32_t frame) {
	uint32_t rec = c.plan + PLAN_HDR - PLAN_REC;
	for (;;) {
		rec += PLAN_REC;
		const uint4 p0 = lds128u(rec);
		/* the end mark (the second test only bounds the walk should a plan ever lack it) */
		if ((p0.x & 0xffu) == P_STOP || rec - c.plan > PLAN_WALK_MAX) break;
		if (plan_record_generic<NS, CTAB, OTHER>(c, rec, p0, row_s, row_r, frame)) return;
	}
}

/* One steady stretch: its own function, so that the hot loop gets its own register
 * allocation whatever the general path around the call needs.  OTHER: the plan has
 * serial self-PM records (plan_other); feed-forward plans run in the other instance. */
template <bool CTAB, bool OTHER>
__device__ __noinline__ void run_block_fast(uint32_t sb, uint32_t plan, int lane, float coeff,
		uint32_t len, float *row_s, float *row_r, uint32_t frame) {
	HotCtx c;
	c.sb = sb; c.plan = plan; c.lane = lane; c.coeff = coeff;
	for (uint32_t oc = 0; oc < len; oc += FastCfg<FAST_NS>::CHUNKF) {
		c.oc = oc;
		run_chunk_plan<FAST_NS, CTAB, OTHER>(c, row_s, row_r, frame + oc);
	}
}
