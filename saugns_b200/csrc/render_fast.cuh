/* render_fast.cuh -- part of kernels.cu (one translation unit; included inside namespace saugen):
 * the LOWERED form of a steady-stretch plan and the chunk loop that runs it.
 *
 * steady_plan (render_plan.cuh) writes one record per bytecode instruction that does something
 * per chunk.  Interpreting those records costs a chain of flag tests per record and a round trip
 * through the warp's shared-memory work buffers for every value that passes from one operator
 * to the next.  plan_lower() rewrites, in place, the records whose shape is one of the common
 * ones into SPECIALISED kinds: the record's kind then names a straight-line variant (how the
 * frequency arrives x where the phase modulation comes from), its amplitude variant sits in two
 * bits, and operands that the record just before produced are taken from REGISTERS ("val": the
 * four samples per lane the previous record left) instead of a buffer; a result nobody reads
 * from its buffer later is not stored at all (one backward liveness pass over the records).
 * Everything else stays in the general form and runs through plan_record_generic().
 *
 * Covered: the fused wave operator (sauPhasor_fill + sauWOsc_run + amplitude + block_mix,
 * generator.c:548-602, wosc.h:135-169,238-266) with a uniform frequency, a frequency vector or a
 * constant ratio to one, with or without phase modulation, amplitude constant or on a lin / xpe /
 * lge trajectory; the range parameter with uniform ends (generator.c:465-467); the voice output
 * (generator.c:772-786).  Coefficient-plane launches only (CTAB).  The arithmetic is the same
 * statements as render_plan.cuh's, in the same order: results are bit-identical.
 */
#pragma once

enum : uint32_t {
	X_OSC0 = 32,           /* X_OSC0 + 3 * fs + pm: fs 0 = uniform increment (w6), 1 = frequency vector,
	                        * 2 = constant (w6) times a vector; pm 0 = none, 1 = buffer c, 2 = val */
	X_RANGE = X_OSC0 + 9,  /* uniform ends (w6, w7), modulator from buffer c or val */
	X_VOUT,                /* carrier from buffer a or val */
	X_COUNT1,              /* counting pass of a team (render_team.cuh): add the rounded increments of a
	                        * frequency vector to the accumulator ... */
	X_COUNT2,              /* ... of a constant (w6) times a vector */
	X_SAVE,                /* team phases: the value just produced (val, or buffer a: flag 1) -> the stretch's
	                        * cache in global memory (64-bit address in w2, w3; index = frame in the stretch) */
	X_LOAD,                /* ... and back: -> val (and buffer a: flag 1), in the place of the record that
	                        * produced it in an earlier phase */
	X_NOP,                 /* a counting record outside its window */
};
/* w1 bits 16..23 of a lowered record */
enum : uint32_t {
	XF_ST = 1,             /* the result is read from its buffer later: store it */
	XF_SRC_VAL = 4,        /* the frequency vector / RANGE modulator / VOUT carrier is val */
	XF_AMP_SHIFT = 3,      /* bits 3..4: 0 constant (w7), 1 lin, 2 xpe, 3 lge (second slot) */
};
constexpr uint32_t XNONE = 0xffffu;

/* lane 0 only; the caller syncs the warp afterwards.  `plan` = first record. */
__device__ __noinline__ void plan_lower(uint32_t plan, uint32_t nrec) {
	/* forward: classify, find operands the record before left in registers */
	uint32_t prev_out = XNONE;
	for (uint32_t r = 0; r < nrec; ++r) {
		const uint32_t a = plan + r * PLAN_REC;
		const uint32_t w0 = lds32(a), w1 = lds32(a + 4);
		const uint32_t kind = w0 & 0xffu, fl = (w0 >> 8) & 0xffu;
		const uint32_t bufa = (w0 >> 16) & 0xffu, bufb = w0 >> 24;
		uint32_t out = XNONE;
		if (kind == P_WLEAF || kind == P_WTAIL) {
			const uint32_t bufc = w1 & 0xffu, mb = (w1 >> 8) & 0xffu;
			uint32_t fs = 3, src = NO_BUF;
			if (fl & PF_FUNI) fs = 0;
			else if (kind == P_WTAIL) { fs = 1; src = bufb; }
			else if (fl & PF_FMUL) { fs = 2; src = mb; }
			uint32_t amp = 4;
			if (fl & PF_ACONST) amp = 0;
			else if (fl & PF_AEXT) {
				const uint32_t t = lds32(a + PLAN_REC) >> 8;
				amp = t == (uint32_t) sau::L_lin ? 1u : t == (uint32_t) sau::L_xpe ? 2u : 3u;
			}
			if (fs < 3 && amp < 4 && !(fl & PF_ABUF)) {
				uint32_t xf = amp << XF_AMP_SHIFT;
				const uint32_t pm = bufc == NO_BUF ? 0u : (bufc == prev_out ? 2u : 1u);
				if (fs && src == prev_out && pm != 2u) xf |= XF_SRC_VAL;
				sts32(a, (w0 & ~0xffu) | (X_OSC0 + 3u * fs + pm));
				sts32(a + 4, (w1 & 0xffffu) | xf << 16);
				/* second half reordered for the lowered code: {diff_scale, inc / v0} are needed early
				 * (one 64-bit load), {diff_offset, held amplitude} late (another) */
				const uint32_t doff = lds32(a + 20), incv = lds32(a + 24);
				sts32(a + 20, incv); sts32(a + 24, doff);
				out = bufa;
			}
			if (fl & PF_AEXT) ++r;             /* the second slot */
		} else if (kind == P_RANGE && (fl & PF_FUNI)) {
			const uint32_t m = w1 & 0xffu;
			sts32(a, (w0 & ~0xffu) | X_RANGE);
			sts32(a + 4, (w1 & 0xffffu) | (m == prev_out ? XF_SRC_VAL : 0u) << 16);
			out = bufa;
		} else if (kind == P_VOUT) {
			sts32(a, (w0 & ~0xffu) | X_VOUT);
			sts32(a + 4, (w1 & 0xffffu) | (bufa == prev_out ? XF_SRC_VAL : 0u) << 16);
		}
		prev_out = out;
	}
	/* backward: which results are read from their buffer later (buffers >= 32: always stored;
	 * a general record may read anything) */
	uint32_t needed = 0;
	for (uint32_t r = nrec; r-- > 0; ) {
		const uint32_t a = plan + r * PLAN_REC;
		const uint32_t w0 = lds32(a), w1 = lds32(a + 4);
		const uint32_t kind = w0 & 0xffu, fl = (w0 >> 8) & 0xffu, xf = (w1 >> 16) & 0xffu;
		const uint32_t bufa = (w0 >> 16) & 0xffu, bufb = w0 >> 24;
		auto need = [&](uint32_t b) { if (b < 32u) needed |= 1u << b; };
		auto is_needed = [&](uint32_t b) { return b >= 32u || ((needed >> b) & 1u); };
		if (kind == P_EXT) continue;
		if (kind >= X_OSC0 && kind < X_RANGE) {
			const uint32_t v = kind - X_OSC0, fs = v / 3u, pm = v % 3u;
			if (is_needed(bufa)) sts32(a + 4, w1 | XF_ST << 16);
			if (bufa < 32u) needed &= ~(1u << bufa);
			if (fl & PF_LAYER) need(bufa);
			if (fs == 1 && !(xf & XF_SRC_VAL)) need(bufb);
			if (fs == 2 && !(xf & XF_SRC_VAL)) need((w1 >> 8) & 0xffu);
			if (pm == 1) need(w1 & 0xffu);
		} else if (kind == X_RANGE) {
			if (is_needed(bufa)) sts32(a + 4, w1 | XF_ST << 16);
			if (bufa < 32u) needed &= ~(1u << bufa);
			if (!(xf & XF_SRC_VAL)) need(w1 & 0xffu);
		} else if (kind == X_VOUT) {
			if (!(xf & XF_SRC_VAL)) need(bufa);
		} else {
			needed = 0xffffffffu;
		}
	}
}

/* ---- the specialised fused wave operator ------------------------------------ */

/* phases of this lane's four samples (sauPhasor_fill, wosc.h:135-169); `st` = the operator's
 * {i0, i1, prev_Is} group, already loaded.  Returns the accumulator after the chunk (what
 * lane 31 writes back with the oscillator's look-back values, xosc_core). */
template <int FS, int PM>
__device__ __forceinline__ uint32_t xphase(const HotCtx &c, const uint4 p0, const uint32_t incv, const uint32_t xf,
		const uint4 st, const float val[4], uint32_t ph[4]) {
	uint32_t acc;
	if (FS == 0) {
		const uint32_t inc = incv;
		const uint32_t base = st.x + inc * (uint32_t) (c.lane * 4);
#pragma unroll
		for (int k = 0; k < 4; ++k) ph[k] = base + inc * (uint32_t) (k + 1);
		acc = ph[3];                   /* (lane 31's is the chunk's last) */
	} else {
		float fr[4];
		const uint32_t src = FS == 1 ? p0.x >> 24 : (p0.y >> 8) & 0xffu;
		if (xf & XF_SRC_VAL) {
#pragma unroll
			for (int k = 0; k < 4; ++k) fr[k] = val[k];
		} else {
			fld<4>(c, src, fr);
		}
		if (FS == 2) {                 /* a constant ratio to a modulated parent frequency (line.c:417-445, no goal) */
			const float v0 = __uint_as_float(incv);
#pragma unroll
			for (int k = 0; k < 4; ++k) fr[k] = v0 * fr[k];
		}
		const float coeff = lds32f(c.plan + PH_COEFF);
		uint32_t run = 0;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			run += ftoi_lo32(coeff * fr[k]);
			ph[k] = run;
		}
		const uint32_t incl = scan_incl_u32(run, c.lane);
		const uint32_t base = st.x + (incl - run);
#pragma unroll
		for (int k = 0; k < 4; ++k) ph[k] += base;
		acc = st.x + incl;
	}
	if (PM == 1) {
		float pm[4];
		fld<4>(c, p0.y & 0xffu, pm);
#pragma unroll
		for (int k = 0; k < 4; ++k) ph[k] += ftoi_lo32(pm[k] * 2147483648.f);
	} else if (PM == 2) {
#pragma unroll
		for (int k = 0; k < 4; ++k) ph[k] += ftoi_lo32(val[k] * 2147483648.f);
	}
	return acc;
}

/* sauWOsc_run (wosc.h:238-266) at the phases ph -> s; the accumulator `acc` and the look-back
 * values go back in one 128-bit store by lane 31 */
__device__ __forceinline__ float xosc_core(const HotCtx &c, const uint4 p0, const uint32_t rec, const uint4 st,
		const uint32_t acc, const float ds, const uint32_t ph[4], float s[4]) {
	const uint32_t op = p0.z;
	uint32_t pph = __shfl_up_sync(FULL, ph[3], 1);
	if (c.lane == 0) pph = lds32(op + OS_I1);
	int32_t d[4];
	d[0] = (int32_t) (ph[0] - pph);
#pragma unroll
	for (int k = 1; k < 4; ++k) d[k] = (int32_t) (ph[k] - ph[k - 1]);
	float xq[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) xq[k] = div_scale_by_int(ds, d[k]);        /* wosc.h:254-256 */
	double Is[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint32_t ind = ph[k] >> sau::WAVE_SLENBITS;
		const double2 hi = lds128d(p0.w + (ind << 4));
		const float2 lo = lds64f(p0.w + CTAB_PLANE_BYTES + (ind << 3));
		Is[k] = horner_frac(hi.x, hi.y, (double) lo.x, ph[k]) + (double) lo.y;
	}
	double pIs = __shfl_up_sync(FULL, Is[3], 1);
	if (c.lane == 0) {
		uint2 pg;
		asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(pg.x), "=r"(pg.y) : "r"(op + OS_PREV));
		pIs = __hiloint2double((int) pg.y, (int) pg.x);
	}
	__syncwarp();              /* lane 0 holds the look-back values before lane 31 rewrites them */
	const float2 late = lds64f(rec + 24);          /* diff_offset, held amplitude */
	bool z = false;
#pragma unroll
	for (int k = 0; k < 4; ++k) z |= (d[k] == 0);
	if (__any_sync(FULL, z)) {
		/* some phase difference is zero: the output repeats (wosc.h:251-252), out of line */
		if (c.lane == 31) sts32(op + OS_I0, acc);
		const uint4 h = lds128u(c.plan);
		ColdCtx cc;
		cc.tab = reinterpret_cast<const float*>((uint64_t) h.x | ((uint64_t) h.y << 32));
		cc.wc = reinterpret_cast<const WaveCoeffs*>((uint64_t) h.z | ((uint64_t) h.w << 32));
		cc.wave_mask = lds32(c.plan + PH_WAVE_MASK); cc.lane = c.lane;
		PhaseVec<4> pv;
#pragma unroll
		for (int k = 0; k < 4; ++k) pv.v[k] = ph[k];
		OpState *o = reinterpret_cast<OpState*>(__cvta_shared_to_generic(op));
		const SampVec<4> sv = wosc_zero_diff<4>(cc, o, pv);
#pragma unroll
		for (int k = 0; k < 4; ++k) s[k] = sv.v[k];
	} else {
		const double doff = (double) late.x;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const double dI = Is[k] - (k ? Is[k - 1] : pIs);
			s[k] = (float) (dI * (double) xq[k] + doff);
		}
		if (c.lane == 31) {
			asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(op + OS_I0), "r"(acc), "r"(ph[3]),
					"r"((uint32_t) __double2loint(Is[3])), "r"((uint32_t) __double2hiint(Is[3])) : "memory");
			sts32(op + OS_PREVS, __float_as_uint(s[3]));
		}
	}
	return late.y;
}

/* the operator's amplitude for this lane's samples: held, or on a lin / xpe / lge trajectory
 * (sauLine_fill_lin / _xpe / _lge, line.c:65-140, from the stretch constants in the second slot) */
template <int AMP>
__device__ __forceinline__ void xamp(const HotCtx &c, const uint32_t rec, const float av, float am[4]) {
	if (AMP == 0) {
#pragma unroll
		for (int k = 0; k < 4; ++k) am[k] = av;
		return;
	}
	const uint4 e = lds128u(rec + PLAN_REC);
	const float w4 = lds32f(rec + PLAN_REC + 16);
	const float inv = __uint_as_float(e.z), w3 = __uint_as_float(e.w);
	const uint32_t i0 = e.y + c.oc + (uint32_t) (c.lane * 4);
	if (AMP == 1) {
#pragma unroll
		for (int k = 0; k < 4; ++k) am[k] = (sau::i2f((int32_t) (i0 + k)) * w3) + w4;
	} else if (AMP == 2) {
#pragma unroll
		for (int k = 0; k < 4; ++k) am[k] = sau::expramp6(1.f - sau::u2f(i0 + k) * inv) * w3 + w4;
	} else {
#pragma unroll
		for (int k = 0; k < 4; ++k) am[k] = sau::expramp6(sau::u2f(i0 + k) * inv) * w3 + w4;
	}
}

/* A record in the general form, out of line: the lowered loop's registers are for its own
 * straight-line variants.  Returns the record's address after it (the second slot consumed). */
template <bool OTHER>
__device__ __noinline__ uint32_t lowered_generic(uint32_t sb, uint32_t plan, int lane, uint32_t oc, uint32_t rec) {
	HotCtx g;
	g.sb = sb; g.plan = plan; g.lane = lane; g.oc = oc;
	g.coeff = lds32f(plan + PH_COEFF);
	const uint4 p0 = lds128u(rec);
	/* (the voice output of a lowered plan is always X_VOUT: no rows needed here) */
	plan_record_generic<FAST_NS, true, OTHER>(g, rec, p0, nullptr, nullptr, 0u);
	return rec;
}

/* One chunk of a lowered plan.  val: the four samples the previous specialised record produced. */
template <bool OTHER, bool TEAM = false>
__device__ __forceinline__ void run_chunk_lowered(const HotCtx &c) {
	uint32_t rec = c.plan + PLAN_HDR - PLAN_REC;
	float val[4] = {0.f, 0.f, 0.f, 0.f};
	for (;;) {
		rec += PLAN_REC;
		const uint4 p0 = lds128u(rec);
		const uint32_t kind = p0.x & 0xffu;
		if (kind == P_STOP || rec - c.plan > PLAN_WALK_MAX) break;
		if (kind < X_OSC0) {
			rec = lowered_generic<OTHER>(c.sb, c.plan, c.lane, c.oc, rec);
			continue;
		}
		const uint32_t flags = (p0.x >> 8) & 0xffu, xf = (p0.y >> 16) & 0xffu;
		const uint32_t bufa = (p0.x >> 16) & 0xffu;
		if (kind < X_RANGE) {
			uint4 st;                                /* i0 now; i1, prev_Is where they are needed (xosc_core) */
			st.x = lds32(p0.z + OS_I0); st.y = st.z = st.w = 0u;
			const float2 early = lds64f(rec + 16);         /* diff_scale, inc / v0 */
			const uint32_t incv = __float_as_uint(early.y);
			uint32_t ph[4], acc;
			switch (kind - X_OSC0) {
			case 0: acc = xphase<0, 0>(c, p0, incv, xf, st, val, ph); break;
			case 1: acc = xphase<0, 1>(c, p0, incv, xf, st, val, ph); break;
			case 2: acc = xphase<0, 2>(c, p0, incv, xf, st, val, ph); break;
			case 3: acc = xphase<1, 0>(c, p0, incv, xf, st, val, ph); break;
			case 4: acc = xphase<1, 1>(c, p0, incv, xf, st, val, ph); break;
			case 5: acc = xphase<1, 2>(c, p0, incv, xf, st, val, ph); break;
			case 6: acc = xphase<2, 0>(c, p0, incv, xf, st, val, ph); break;
			case 7: acc = xphase<2, 1>(c, p0, incv, xf, st, val, ph); break;
			default: acc = xphase<2, 2>(c, p0, incv, xf, st, val, ph); break;
			}
			float s[4];
			const float av = xosc_core(c, p0, rec, st, acc, early.x, ph, s);
			float am[4];
			switch ((xf >> XF_AMP_SHIFT) & 3u) {
			case 0: xamp<0>(c, rec, av, am); break;
			case 1: xamp<1>(c, rec, av, am); break;
			case 2: xamp<2>(c, rec, av, am); break;
			default: xamp<3>(c, rec, av, am); break;
			}
			if (flags & PF_WAVEENV) {                                     /* generator.c:407-426 */
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const float s_amp = am[k] * 0.5f;
					val[k] = (s[k] * s_amp) + fabsf(s_amp);
				}
				if (flags & PF_LAYER) {
					float lay[4];
					fld<4>(c, bufa, lay);
#pragma unroll
					for (int k = 0; k < 4; ++k) val[k] = lay[k] * val[k];
				}
			} else {                                                      /* generator.c:384-397 */
#pragma unroll
				for (int k = 0; k < 4; ++k) val[k] = s[k] * am[k];
				if (flags & PF_LAYER) {
					float lay[4];
					fld<4>(c, bufa, lay);
#pragma unroll
					for (int k = 0; k < 4; ++k) val[k] = lay[k] + val[k];
				}
			}
			if (xf & XF_ST) fst<4>(c, bufa, val);
			if (flags & PF_AEXT) rec += PLAN_REC;                  /* the record's second slot */
			__syncwarp();
		} else if (kind == X_RANGE) {                                  /* generator.c:465-467 */
			float m[4];
			if (xf & XF_SRC_VAL) {
#pragma unroll
				for (int k = 0; k < 4; ++k) m[k] = val[k];
			} else {
				fld<4>(c, p0.y & 0xffu, m);
			}
			const float2 pr = lds64f(rec + 24);
#pragma unroll
			for (int k = 0; k < 4; ++k) { float p = pr.x; p += (pr.y - p) * m[k]; val[k] = p; }
			if (xf & XF_ST) fst<4>(c, bufa, val);
		} else if (TEAM && kind >= X_SAVE) {
			if (kind != X_NOP) {
				float *g = reinterpret_cast<float*>((uint64_t) p0.z | ((uint64_t) p0.w << 32)) + c.oc + c.lane * 4;
				if (kind == X_SAVE) {
					if (flags & 1u) fld<4>(c, bufa, val);
					__stcg(reinterpret_cast<float4*>(g), make_float4(val[0], val[1], val[2], val[3]));
				} else {
					const float4 v = __ldcg(reinterpret_cast<const float4*>(g));
					val[0] = v.x; val[1] = v.y; val[2] = v.z; val[3] = v.w;
					if (flags & 1u) { fst<4>(c, bufa, val); __syncwarp(); }
				}
			}
		} else if (kind >= X_COUNT1) {
			/* sauPhasor_fill's accumulation alone (wosc.h:145-166): sum of lrintf(coeff * f) */
			float fr[4];
			const uint32_t src = kind == X_COUNT1 ? p0.x >> 24 : (p0.y >> 8) & 0xffu;
			if (xf & XF_SRC_VAL) {
#pragma unroll
				for (int k = 0; k < 4; ++k) fr[k] = val[k];
			} else {
				fld<4>(c, src, fr);
			}
			if (kind == X_COUNT2) {
				const float v0 = lds32f(rec + 20);
#pragma unroll
				for (int k = 0; k < 4; ++k) fr[k] = v0 * fr[k];
			}
			const float coeff = lds32f(c.plan + PH_COEFF);
			uint32_t run = 0;
#pragma unroll
			for (int k = 0; k < 4; ++k) run += ftoi_lo32(coeff * fr[k]);
			const uint32_t tot = __reduce_add_sync(FULL, run);
			if (c.lane == 0) sts32(p0.z + OS_I0, lds32(p0.z + OS_I0) + tot);
			__syncwarp();
		} else {                                                       /* X_VOUT, generator.c:772-786 */
			float sv[4];
			if (xf & XF_SRC_VAL) {
#pragma unroll
				for (int k = 0; k < 4; ++k) sv[k] = val[k];
			} else {
				fld<4>(c, bufa, sv);
			}
			const uint4 h1 = lds128u(c.plan + PH_ROW_S);   /* the s row, tile stride, frame0 | write_r << 31 */
			const float2 h2 = lds64f(c.plan + PH_AMP_SCALE);       /* amp_scale, the static pan */
			const float amp_scale = h2.x, pan = h2.y;
			const uint32_t write_r = h1.w >> 31, tstride = h1.z;
			float *row_s = reinterpret_cast<float*>((uint64_t) h1.x | ((uint64_t) h1.y << 32));
			float *row_r = nullptr;
			if (write_r) {
				uint2 rr;
				asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rr.x), "=r"(rr.y) : "r"(c.plan + PH_ROW_R));
				row_r = reinterpret_cast<float*>((uint64_t) rr.x | ((uint64_t) rr.y << 32));
			}
			const uint32_t frame = (h1.w & 0x7fffffffu) + c.oc;
			float s[4], rv[4];
#pragma unroll
			for (int k = 0; k < 4; ++k) { s[k] = sv[k] * amp_scale; rv[k] = s[k] * pan; }
			const uint32_t fl = frame + c.lane * 4;
			if ((frame & 3u) == 0) {
				const size_t at = row_index(fl, tstride);       /* 128-bit streaming stores */
				__stcs(reinterpret_cast<float4*>(row_s + at), make_float4(s[0], s[1], s[2], s[3]));
				if (write_r) __stcs(reinterpret_cast<float4*>(row_r + at), make_float4(rv[0], rv[1], rv[2], rv[3]));
			} else {
				/* segment starting at an odd frame: rare, out of line through the buffers */
				const uint32_t bufb = p0.x >> 24;
				const uint32_t rb = bufb != NO_BUF ? bufb : bufa + 1u;
				fst<4>(c, bufa, s);
				fst<4>(c, rb, rv);
				__syncwarp();
				vout_unaligned(c.sb - c.lane * 16 + bufa * FastCfg<4>::FBUF_BYTES,
						c.sb - c.lane * 16 + rb * FastCfg<4>::FBUF_BYTES,
						row_s, row_r, c.lane, 4, write_r, frame, tstride);
			}
			return;
		}
	}
}

/* One steady stretch on a lowered plan (the coefficient-plane launches); see run_block_fast. */
/* chunks [oc0, oc0 + len) of the stretch; rows, frame of the stretch's first sample and the
 * rest of the context come from the plan header */
template <bool OTHER, bool TEAM = false>
__device__ __noinline__ void run_block_lowered(uint32_t sb, uint32_t plan, int lane, uint32_t oc0, uint32_t len) {
	HotCtx c;
	c.sb = sb; c.plan = plan; c.lane = lane; c.coeff = 0.f;
	for (uint32_t oc = oc0; oc < oc0 + len; oc += FastCfg<FAST_NS>::CHUNKF) {
		c.oc = oc;
		run_chunk_lowered<OTHER, TEAM>(c);
	}
}

/* ---- fused shapes ------------------------------------------------------------ *
 * Most voices of a script share a handful of plan SIGNATURES (the same operator graph with
 * other constants: C3 has two).  For the signatures listed below the chunk loop exists as ONE
 * straight-line function -- the same xphase / xosc_core / xamp statements, instantiated with
 * the record variants as template parameters: no record dispatch, no flag tests.  A plan whose
 * lowered records spell one of the listed signatures runs through it; every other plan through
 * run_chunk_lowered.  The list is a compile-time table; a signature is the sequence of the
 * records' 16-bit codes (below). */
__device__ __forceinline__ uint32_t rec_code(uint32_t w0, uint32_t w1) {
	const uint32_t kind = w0 & 0xffu, fl = (w0 >> 8) & 0xffu, xf = (w1 >> 16) & 0xffu;
	if (kind >= X_OSC0 && kind < X_RANGE)
		return 0x1000u | (kind - X_OSC0) | ((xf >> XF_AMP_SHIFT) & 3u) << 4 | ((fl & PF_WAVEENV) ? 0x40u : 0u) |
			((fl & PF_LAYER) ? 0x80u : 0u) | ((xf & XF_ST) ? 0x100u : 0u) | ((xf & XF_SRC_VAL) ? 0x200u : 0u);
	if (kind == X_RANGE) return 0x2000u | ((xf & XF_ST) ? 0x100u : 0u) | ((xf & XF_SRC_VAL) ? 0x200u : 0u);
	if (kind == X_VOUT) return 0x3000u | ((xf & XF_SRC_VAL) ? 0x200u : 0u);
	/* the records of a team's phase plans (render_team.cuh); a counting record keeps its code in and out of its window */
	if (kind == X_SAVE) return 0x4000u | (fl & 1u);
	if (kind == X_LOAD) return 0x5000u | (fl & 1u);
	if (kind == X_COUNT1 || kind == X_COUNT2) return 0x6000u | (kind == X_COUNT1 ? 1u : 2u) | ((xf & XF_SRC_VAL) ? 0x200u : 0u);
	if (kind == X_NOP) return 0x6000u | (fl & 3u) | ((xf & XF_SRC_VAL) ? 0x200u : 0u);
	return 0xffffu;            /* not a fusable record */
}

template <int FS, int PM, int AMP, bool ENV, bool LAYER, bool ST, bool SRCVAL>
struct FOsc {
	static constexpr uint32_t SLOTS = AMP ? 2 : 1;
	static constexpr uint16_t CODE = 0x1000u | (3 * FS + PM) | AMP << 4 | (ENV ? 0x40u : 0u) | (LAYER ? 0x80u : 0u) |
		(ST ? 0x100u : 0u) | (SRCVAL ? 0x200u : 0u);
	static __device__ __forceinline__ void exec(const HotCtx &c, const uint32_t rec, float val[4]) {
		const uint4 p0 = lds128u(rec);
		const float2 early = lds64f(rec + 16);             /* diff_scale, inc / v0 */
		uint4 st;
		st.x = lds32(p0.z + OS_I0); st.y = st.z = st.w = 0u;
		uint32_t ph[4];
		const uint32_t acc = xphase<FS, PM>(c, p0, __float_as_uint(early.y), SRCVAL ? XF_SRC_VAL : 0u, st, val, ph);
		float s[4], am[4];
		const float av = xosc_core(c, p0, rec, st, acc, early.x, ph, s);
		xamp<AMP>(c, rec, av, am);
		const uint32_t bufa = (p0.x >> 16) & 0xffu;
		if (ENV) {
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const float s_amp = am[k] * 0.5f;
				val[k] = (s[k] * s_amp) + fabsf(s_amp);
			}
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k) val[k] = s[k] * am[k];
		}
		if (LAYER) {
			float lay[4];
			fld<4>(c, bufa, lay);
#pragma unroll
			for (int k = 0; k < 4; ++k) val[k] = ENV ? lay[k] * val[k] : lay[k] + val[k];
		}
		if (ST) fst<4>(c, bufa, val);
		__syncwarp();
	}
};
template <bool ST, bool SRCVAL>
struct FRange {
	static constexpr uint32_t SLOTS = 1;
	static constexpr uint16_t CODE = 0x2000u | (ST ? 0x100u : 0u) | (SRCVAL ? 0x200u : 0u);
	static __device__ __forceinline__ void exec(const HotCtx &c, const uint32_t rec, float val[4]) {
		float m[4];
		if (SRCVAL) {
#pragma unroll
			for (int k = 0; k < 4; ++k) m[k] = val[k];
		} else {
			fld<4>(c, lds32(rec + 4) & 0xffu, m);
		}
		const float2 pr = lds64f(rec + 24);
#pragma unroll
		for (int k = 0; k < 4; ++k) { float p = pr.x; p += (pr.y - p) * m[k]; val[k] = p; }
		if (ST) fst<4>(c, (lds32(rec) >> 16) & 0xffu, val);
	}
};
template <bool SRCVAL>
struct FVout {
	static constexpr uint32_t SLOTS = 1;
	static constexpr uint16_t CODE = 0x3000u | (SRCVAL ? 0x200u : 0u);
	static __device__ __forceinline__ void exec(const HotCtx &c, const uint32_t rec, float val[4]) {
		float sv[4];
		if (SRCVAL) {
#pragma unroll
			for (int k = 0; k < 4; ++k) sv[k] = val[k];
		} else {
			fld<4>(c, (lds32(rec) >> 16) & 0xffu, sv);
		}
		/* two loads from the plan header's output slot (render_plan.cuh:PH_ROW_S), not the record */
		const uint4 h1 = lds128u(c.plan + PH_ROW_S);
		const float2 h2 = lds64f(c.plan + PH_AMP_SCALE);
		const float amp_scale = h2.x, pan = h2.y;
		const uint32_t write_r = h1.w >> 31, tstride = h1.z;
		float *row_s = reinterpret_cast<float*>((uint64_t) h1.x | ((uint64_t) h1.y << 32));
		const uint32_t frame = (h1.w & 0x7fffffffu) + c.oc;
		float s[4], rv[4];
#pragma unroll
		for (int k = 0; k < 4; ++k) { s[k] = sv[k] * amp_scale; rv[k] = s[k] * pan; }
		const size_t at = row_index(frame + c.lane * 4, tstride);       /* (fused plans: aligned frames only) */
		__stcs(reinterpret_cast<float4*>(row_s + at), make_float4(s[0], s[1], s[2], s[3]));
		if (write_r) {
			uint2 rr;
			asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rr.x), "=r"(rr.y) : "r"(c.plan + PH_ROW_R));
			float *row_r = reinterpret_cast<float*>((uint64_t) rr.x | ((uint64_t) rr.y << 32));
			__stcs(reinterpret_cast<float4*>(row_r + at), make_float4(rv[0], rv[1], rv[2], rv[3]));
		}
	}
};

/* a team's phase plans: value to / from the stretch's cache, counting records */
template <bool FROMBUF>
struct FSave {
	static constexpr uint32_t SLOTS = 1;
	static constexpr uint16_t CODE = 0x4000u | (FROMBUF ? 1u : 0u);
	static __device__ __forceinline__ void exec(const HotCtx &c, const uint32_t rec, float val[4]) {
		const uint4 p0 = lds128u(rec);
		float *g = reinterpret_cast<float*>((uint64_t) p0.z | ((uint64_t) p0.w << 32)) + c.oc + c.lane * 4;
		if (FROMBUF) fld<4>(c, (p0.x >> 16) & 0xffu, val);
		__stcg(reinterpret_cast<float4*>(g), make_float4(val[0], val[1], val[2], val[3]));
	}
};
template <bool TOBUF>
struct FLoad {
	static constexpr uint32_t SLOTS = 1;
	static constexpr uint16_t CODE = 0x5000u | (TOBUF ? 1u : 0u);
	static __device__ __forceinline__ void exec(const HotCtx &c, const uint32_t rec, float val[4]) {
		const uint4 p0 = lds128u(rec);
		const float *g = reinterpret_cast<const float*>((uint64_t) p0.z | ((uint64_t) p0.w << 32)) + c.oc + c.lane * 4;
		const float4 v = __ldcg(reinterpret_cast<const float4*>(g));
		val[0] = v.x; val[1] = v.y; val[2] = v.z; val[3] = v.w;
		if (TOBUF) { fst<4>(c, (p0.x >> 16) & 0xffu, val); __syncwarp(); }
	}
};
template <int WHICH, bool SRCVAL>
struct FCount {
	static constexpr uint32_t SLOTS = 1;
	static constexpr uint16_t CODE = 0x6000u | WHICH | (SRCVAL ? 0x200u : 0u);
	static __device__ __forceinline__ void exec(const HotCtx &c, const uint32_t rec, float val[4]) {
		const uint4 p0 = lds128u(rec);
		if ((p0.x & 0xffu) == X_NOP) return;           /* outside its window */
		float fr[4];
		if (SRCVAL) {
#pragma unroll
			for (int k = 0; k < 4; ++k) fr[k] = val[k];
		} else {
			fld<4>(c, WHICH == 1 ? p0.x >> 24 : (p0.y >> 8) & 0xffu, fr);
		}
		if (WHICH == 2) {
			const float v0 = lds32f(rec + 20);
#pragma unroll
			for (int k = 0; k < 4; ++k) fr[k] = v0 * fr[k];
		}
		const float coeff = lds32f(c.plan + PH_COEFF);
		uint32_t run = 0;
#pragma unroll
		for (int k = 0; k < 4; ++k) run += ftoi_lo32(coeff * fr[k]);
		const uint32_t tot = __reduce_add_sync(FULL, run);
		if (c.lane == 0) sts32(p0.z + OS_I0, lds32(p0.z + OS_I0) + tot);
		__syncwarp();
	}
};

template <class... R>
struct Shape {
	static constexpr uint32_t N = sizeof...(R);
	static __device__ __forceinline__ bool match(const uint16_t *codes, uint32_t n) {
		if (n != N) return false;
		const uint16_t want[N] = {R::CODE...};
		bool ok = true;
#pragma unroll
		for (uint32_t i = 0; i < N; ++i) ok = ok && codes[i] == want[i];
		return ok;
	}
	static __device__ __noinline__ void run(uint32_t sb, uint32_t plan, int lane, uint32_t oc0, uint32_t len) {
		HotCtx c;
		c.sb = sb; c.plan = plan; c.lane = lane; c.coeff = 0.f;
		for (uint32_t oc = oc0; oc < oc0 + len; oc += FastCfg<FAST_NS>::CHUNKF) {
			c.oc = oc;
			float val[4] = {0.f, 0.f, 0.f, 0.f};
			uint32_t rec = plan + PLAN_HDR;
			((R::exec(c, rec, val), rec += R::SLOTS * PLAN_REC), ...);
		}
	}
};

/* The listed signatures.  (FOsc<FS, PM, AMP, ENV, LAYER, ST, SRCVAL>: FS 0 uniform / 1 vector / 2
 * constant x vector; PM 0 none / 1 buffer / 2 val; AMP 0 held / 1 lin / 2 xpe / 3 lge.) */
using ShapeW1      = Shape<FOsc<0, 0, 0, false, false, false, false>, FVout<true>>;                 /* one plain wave operator */
using ShapeW1x     = Shape<FOsc<0, 0, 2, false, false, false, false>, FVout<true>>;                 /* ... with an xpe amplitude ramp */
using ShapePM2     = Shape<FOsc<0, 0, 0, false, false, false, false>,
                           FOsc<0, 2, 0, false, false, false, false>, FVout<true>>;                 /* carrier + one PM modulator */
using ShapeC3PM    = Shape<FOsc<0, 0, 0, false, false, false, false>, FOsc<0, 2, 1, false, false, false, false>,
                           FOsc<0, 2, 2, false, false, false, false>, FVout<true>>;                 /* 3-operator PM chain, lin / xpe ramps */
using ShapeC3PMh   = Shape<FOsc<0, 0, 0, false, false, false, false>, FOsc<0, 2, 0, false, false, false, false>,
                           FOsc<0, 2, 2, false, false, false, false>, FVout<true>>;                 /* ... once the modulator's ramp has ended */
using ShapePM3     = Shape<FOsc<0, 0, 0, false, false, false, false>, FOsc<0, 2, 0, false, false, false, false>,
                           FOsc<0, 2, 0, false, false, false, false>, FVout<true>>;                 /* ... all amplitudes held */
using ShapeC3FMh   = Shape<FOsc<0, 0, 0, true, false, false, false>, FRange<true, true>,
                           FOsc<2, 0, 0, false, false, false, true>, FOsc<1, 2, 2, false, false, false, false>,
                           FVout<true>>;                                                             /* range-FM, modulator ramp ended */
using ShapeFM3h    = Shape<FOsc<0, 0, 0, true, false, false, false>, FRange<true, true>,
                           FOsc<2, 0, 0, false, false, false, true>, FOsc<1, 2, 0, false, false, false, false>,
                           FVout<true>>;                                                             /* ... all amplitudes held */
using ShapeC3FM    = Shape<FOsc<0, 0, 1, true, false, false, false>, FRange<true, true>,
                           FOsc<2, 0, 0, false, false, false, true>, FOsc<1, 2, 2, false, false, false, false>,
                           FVout<true>>;                                                             /* range-FM carrier + ratio PM modulator */
/* the two phases of a team on the range-FM voice (render_team.cuh): modulator + range, kept, with the two
 * oscillators counted that it feeds; then those two and the voice output */
using ShapeC3FMq0  = Shape<FOsc<0, 0, 1, true, false, false, false>, FRange<true, true>, FSave<false>,
                           FCount<2, true>, FCount<1, false>>;
using ShapeC3FMq0h = Shape<FOsc<0, 0, 0, true, false, false, false>, FRange<true, true>, FSave<false>,
                           FCount<2, true>, FCount<1, false>>;                                       /* ... modulator ramp ended */
using ShapeC3FMq1  = Shape<FLoad<true>, FOsc<2, 0, 0, false, false, false, true>,
                           FOsc<1, 2, 2, false, false, false, false>, FVout<true>>;
constexpr uint32_t FUSED_NONE = 0;
__device__ uint32_t g_sig_dump[36];           /* developer aid: the first stretch's signature (saugen_debug_signature) */

/* lane 0: which listed shape the lowered plan spells (0 = none) */
__device__ __noinline__ uint32_t fused_match(uint32_t plan, uint32_t nrec, bool dump) {
	uint16_t codes[16];
	uint32_t n = 0;
	for (uint32_t r = 0; r < nrec; ++r) {
		const uint32_t a = plan + PLAN_HDR + r * PLAN_REC;
		const uint32_t w0 = lds32(a), w1 = lds32(a + 4);
		if ((w0 & 0xffu) == P_EXT) continue;
		if (n >= 16) return FUSED_NONE;
		codes[n++] = (uint16_t) rec_code(w0, w1);
	}
	if (dump) {
		g_sig_dump[0] = n;
		for (uint32_t i = 0; i < n; ++i) g_sig_dump[1 + i] = codes[i];
	}
	if (ShapeC3PMh::match(codes, n)) return 6;
	if (ShapeC3FMh::match(codes, n)) return 7;
	if (ShapeC3PM::match(codes, n)) return 1;
	if (ShapeC3FM::match(codes, n)) return 2;
	if (ShapePM2::match(codes, n)) return 3;
	if (ShapeW1::match(codes, n)) return 4;
	if (ShapeW1x::match(codes, n)) return 5;
	if (ShapePM3::match(codes, n)) return 8;
	if (ShapeFM3h::match(codes, n)) return 9;
	if (ShapeC3FMq0::match(codes, n)) return 10;
	if (ShapeC3FMq1::match(codes, n)) return 11;
	if (ShapeC3FMq0h::match(codes, n)) return 12;
	return FUSED_NONE;
}
__device__ __forceinline__ void fused_run(uint32_t which, uint32_t sb, uint32_t plan, int lane, uint32_t oc0, uint32_t len) {
	switch (which) {
	case 1: ShapeC3PM::run(sb, plan, lane, oc0, len); break;
	case 2: ShapeC3FM::run(sb, plan, lane, oc0, len); break;
	case 3: ShapePM2::run(sb, plan, lane, oc0, len); break;
	case 4: ShapeW1::run(sb, plan, lane, oc0, len); break;
	case 5: ShapeW1x::run(sb, plan, lane, oc0, len); break;
	case 6: ShapeC3PMh::run(sb, plan, lane, oc0, len); break;
	case 7: ShapeC3FMh::run(sb, plan, lane, oc0, len); break;
	case 8: ShapePM3::run(sb, plan, lane, oc0, len); break;
	case 10: ShapeC3FMq0::run(sb, plan, lane, oc0, len); break;
	case 11: ShapeC3FMq1::run(sb, plan, lane, oc0, len); break;
	case 12: ShapeC3FMq0h::run(sb, plan, lane, oc0, len); break;
	default: ShapeFM3h::run(sb, plan, lane, oc0, len); break;
	}
}
