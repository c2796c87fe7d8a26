/* render_kernel.cuh -- part of kernels.cu (one translation unit; included inside namespace saugen):
 * the render kernels: per-voice unit loop, schedulers, table staging. */
#pragma once

/* ---- render kernel ------------------------------------------------------ */

static_assert(PLAN_FBUF == FastCfg<FAST_NS>::FBUF_BYTES, "plan-time scratch words sit in the fast buffers");
constexpr uint32_t OP_VEC = sizeof(OpState) / 16;
static_assert(sizeof(OpState) == 192, "OpState layout (device_types.h)");

/* operator states of the current voice program: HBM <-> shared memory */
__device__ __forceinline__ void ops_load(Ctx &c, uint32_t cnt) {
	uint4 *dst = reinterpret_cast<uint4*>(c.sops);
	for (uint32_t i = c.lane; i < cnt * OP_VEC; i += 32) {
		const uint32_t slot = i / OP_VEC, w = i % OP_VEC;
		dst[i] = __ldcg(reinterpret_cast<const uint4*>(c.gops + c.prog_ops[slot]) + w);
	}
	__syncwarp();
}
__device__ __forceinline__ void ops_store(Ctx &c, uint32_t cnt) {
	__syncwarp();
	const uint4 *src = reinterpret_cast<const uint4*>(c.sops);
	for (uint32_t i = c.lane; i < cnt * OP_VEC; i += 32) {
		const uint32_t slot = i / OP_VEC, w = i % OP_VEC;
		__stcg(reinterpret_cast<uint4*>(c.gops + c.prog_ops[slot]) + w, src[i]);
	}
	__syncwarp();
}

/* Units [u0, u1) of one voice of one call.  A unit is a stretch of one
 * inter-event segment, starting at a multiple of REF_BLOCK inside it (the
 * reference's own block grid, generator.c:854-878). */
/* The position word of an amplitude trajectory's P_EXT slot (steady_plan, render_plan.cuh) at a stretch
 * start, from the operator's line: the position, centred for a linear one.  op = shared address. */
__device__ __forceinline__ uint32_t plan_ext_pos(uint32_t op, uint32_t type) {
	const uint32_t pos = lds32(op + OS_LINE + 16u * LINE_AMP + 8u), end = lds32(op + OS_LINE + 16u * LINE_AMP + 12u);
	return type == (uint32_t) sau::L_lin ? pos - (end / 2u) : pos;
}

/* developer aid: a time line of the launch's first warp (g_team_dump[18..]: cycles since the kernel started) */
#define SAUGEN_TRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_team_dump[i] = (uint32_t) (clock64() - g_trace_t0); } while (0)
__device__ long long g_trace_t0;

__device__ __noinline__ void render_units(Ctx &c, FastCtx &fc, const CallDesc *cd,
		const SegDesc *segs, const UnitDesc *units, uint32_t lv, uint32_t u0, uint32_t u1) {
	const GenDesc *g = cd->gen;
	const int lane = c.lane;
	const uint32_t v = g->voice_begin + lv;
	const uint32_t nlv = g->voice_end - g->voice_begin;
	c.g = g;
	c.gops = g->ops;
	c.coeff = g->coeff;
	fc.coeff = g->coeff; fc.amp_scale = g->amp_scale;
	VoiceState *vsp = &g->voices[v];
	VoiceState vs;
	uint32_t first_w;
	{
		/* lane 0 reads (from L2: another SM may have written it), every lane gets
		 * the same copy */
		const uint32_t *src = reinterpret_cast<const uint32_t*>(vsp);
		uint32_t *dst = reinterpret_cast<uint32_t*>(&vs);
		/* ONE round trip for everything whose address is known now: the voice's state (lanes 0..15), its
		 * event-list bounds (16, 17), the first unit (20..22) and, in case it is the unit's, segment 0 (24..26) */
		static_assert(sizeof(VoiceState) / 4 <= 16 && sizeof(UnitDesc) == 12 && sizeof(SegDesc) == 12, "lane map");
		const uint32_t *p = nullptr;
		if (lane < (int) (sizeof(VoiceState) / 4)) p = src + lane;
		else if (lane == 16 || lane == 17) p = g->vev_off + v + (lane - 16);
		else if (lane >= 20 && lane < 23 && u0 < u1) p = reinterpret_cast<const uint32_t*>(units + cd->unit_off + u0) + (lane - 20);
		else if (lane >= 24 && lane < 27 && u0 < u1) p = reinterpret_cast<const uint32_t*>(segs + cd->seg_off) + (lane - 24);
		const uint32_t w = p ? __ldcg(p) : 0u;
#pragma unroll
		for (uint32_t i = 0; i < sizeof(VoiceState) / 4; ++i) dst[i] = __shfl_sync(FULL, w, i);
		first_w = w;
	}
	/* the voice program's bytecode and operator list: towards L1 now, read one by one later */
	for (uint32_t i = lane; i < vs.code_len; i += 32)
		asm volatile("prefetch.global.L1 [%0];" :: "l"(g->code + vs.code_off + i));
	if (lane == 0) asm volatile("prefetch.global.L1 [%0];" :: "l"(g->prog_ops + vs.ops_off));
	SAUGEN_TRACE(19);                  /* voice state read */
	const uint32_t ev_lo = __shfl_sync(FULL, first_w, 16), ev_n = __shfl_sync(FULL, first_w, 17) - ev_lo;
	UnitDesc ud0;
	SegDesc sd0;
	ud0.seg = __shfl_sync(FULL, first_w, 20); ud0.off = __shfl_sync(FULL, first_w, 21); ud0.len = __shfl_sync(FULL, first_w, 22);
	sd0.start = __shfl_sync(FULL, first_w, 24); sd0.len = __shfl_sync(FULL, first_w, 25); sd0.ev_end = __shfl_sync(FULL, first_w, 26);
	float *row_s = g->rows_s + (size_t) lv * ROW_TILE;      /* the voice's piece of frame tile 0 */
	float *row_r = g->rows_r + (size_t) lv * ROW_TILE;
	c.tstride = g->row_stride;
	uint32_t loaded = 0;        // operator states currently held in shared memory

	for (uint32_t ui = u0; ui < u1; ++ui) {
		const UnitDesc ud = ui == u0 ? ud0 : units[cd->unit_off + ui];
		const uint32_t si = ud.seg;
		const SegDesc sd = (ui == u0 && si == 0u) ? sd0 : segs[cd->seg_off + si];
		/* this voice's events due at the segment start, in order */
		if (ud.off == 0 && vs.ev_cursor < ev_n && g->vev_idx[ev_lo + vs.ev_cursor] < sd.ev_end) {
			if (loaded) { ops_store(c, loaded); loaded = 0; }
			while (vs.ev_cursor < ev_n && g->vev_idx[ev_lo + vs.ev_cursor] < sd.ev_end) {
				if (lane == 0) {
					apply_event(g, c.wc, &g->events[g->vev_idx[ev_lo + vs.ev_cursor]], &vs);
					vs.ev_cursor++;
					vs.plan_gen = 0;           /* the kept plan is for the voice as it was */
				}
				__syncwarp();
				uint32_t *w = reinterpret_cast<uint32_t*>(&vs);
#pragma unroll
				for (uint32_t i = 0; i < sizeof(VoiceState) / 4; ++i) w[i] = __shfl_sync(FULL, w[i], 0);
			}
		}
		if (vs.duration == 0 || ud.len == 0) continue;
		/* this voice's pan in this segment (VoiceSeg): undecided at the segment's
		 * first unit, else what the unit that started the segment recorded */
		VoiceSeg *vsg = g->vlen + (size_t) si * nlv + lv;
		uint32_t pan_mode = PAN_UNSET;
		if (ud.off != 0) {
			const uint2 pv = __ldcg(reinterpret_cast<const uint2*>(vsg));
			if (pv.x != 0) pan_mode = pv.y;
		}
		c.write_r = pan_mode == PAN_DYNAMIC;
		fc.write_r = c.write_r ? 1u : 0u;
		c.prog_ops = g->prog_ops + vs.ops_off;
		if (!loaded && vs.ops_cnt > 0) {
			ops_load(c, vs.ops_cnt);
			loaded = vs.ops_cnt;
		}
		uint32_t run_total = 0;
		const uint32_t uend = ud.off + ud.len;
		for (uint32_t off = ud.off; off < uend && vs.duration != 0; off += CHUNK) {
			/* whole reference blocks in steady state: the fast path, for as many of the
			 * unit's blocks as one plan holds */
			uint32_t sp = 0;
			uint32_t which_kept = 0xffu;   /* the stretch's fused shape, where it was looked up */
			uint32_t kmax = 0;             /* blocks the plan holds for, from this stretch's start */
			uint32_t teamP = 0;            /* the teams' analysis of the plan: 0 unknown, 1 + P, 0x100 not eligible */
			bool teamP_kept = false;
			uint32_t kept = 0;         /* the plan came from the voice's kept one: 1 + its fused shape (0xff: not looked up) */
			if (off % REF_BLOCK == 0 && uend - off >= (uint32_t) REF_BLOCK &&
					vs.duration >= (uint32_t) REF_BLOCK && vs.code_len && !(fc.wave_mask & TAP_FLAG) &&
					op_ptr(c, vs.carr_slot)->time > 0) {
				uint32_t kb = (uend - off) / (uint32_t) REF_BLOCK;
				if (vs.duration / (uint32_t) REF_BLOCK < kb) kb = vs.duration / (uint32_t) REF_BLOCK;
				if (kb > 0x7fffu) kb = 0x7fffu;        /* 15 bits in steady_plan's result */
				SAUGEN_TRACE(20);          /* operator states loaded, events applied */
				/* A voice's plan holds, as it is, until an event touches the voice or one of the spans
				 * steady_plan found runs out (an operator's time, a line reaching its goal): it is kept,
				 * lowered, in global memory with the number of blocks it still holds for, and comes back
				 * with one coalesced read instead of the walk over the bytecode, the lowering and the
				 * shape look-up.  What moves inside it -- the position of an amplitude trajectory at the
				 * stretch start (the P_EXT slots) -- is put in again from the operator's line. */
				uint4 *kp = nullptr;
				bool have = false;
				if (fc.keep_plans && g->plan_cache && (fc.wave_mask & CTAB_FLAG)) {
					kp = g->plan_cache + (size_t) lv * (18u + 2u * g->plan_cache_recs);
					if (vs.plan_gen && vs.plan_left && vs.plan_so == fc.so && vs.plan_st == fc.st) {
						const uint4 kh = __ldcg(kp);
						if (kh.x == vs.plan_gen && kh.y <= g->plan_cache_recs && kh.y + 3u <= fc.plan_cap) {
							have = true;
							if (!(fc.wave_mask & VERIFY_FLAG)) {
								for (uint32_t i = lane; i < 2u * kh.y; i += 32) {
									const uint4 x = __ldcg(kp + 1 + i);
									asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(fc.plan + PLAN_HDR + 16u * i),
											"r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
								}
								if (lane == 0) sts32(fc.plan + PLAN_HDR + kh.y * PLAN_REC, P_STOP);
								__syncwarp();
								for (uint32_t r = 1 + lane; r < kh.y; r += 32) {
									const uint32_t a = fc.plan + PLAN_HDR + r * PLAN_REC, w0 = lds32(a);
									if ((w0 & 0xffu) == P_EXT) sts32(a + 4, plan_ext_pos(lds32(a - PLAN_REC + 8), w0 >> 8));
								}
								__syncwarp();
								kmax = vs.plan_left;
								sp = (kmax < kb ? kmax : kb) << 16 | kh.y;
								kept = 1u + (kh.z & 0xffu);
								teamP = kh.w; teamP_kept = kh.w != 0u;
							}
						}
					}
				}
				if (!kept) {
					sp = steady_plan(c.sops, fc.so, fc.st, fc.wave_mask, fc.wc, g->code + vs.code_off,
							vs.code_len, fc.plan, fc.plan_cap, 0x7fffu, fc.sb, fc.coeff);
					kmax = (sp >> 16) & 0x7fffu;       /* blocks the plan holds for; this unit has kb */
					if (sp) sp = (sp & 0x8000ffffu) | (kmax < kb ? kmax : kb) << 16;
				}
				if (!sp) vs.plan_gen = 0;
				SAUGEN_TRACE(21);          /* plan built */
			}
			if (sp) {
				const uint32_t nrec = sp & 0x7fffu, nb = (sp >> 16) & 0x7fffu, span = nb * (uint32_t) REF_BLOCK;
				const bool other = (sp >> 31) != 0;
				if (pan_mode == PAN_UNSET)       /* steady => the pan stands still */
					pan_mode = __float_as_uint(op_ptr(c, vs.carr_slot)->line[LINE_PAN].v0);
				{
					/* plan header: what the rare paths and VOUT need */
					const uint64_t tp = reinterpret_cast<uint64_t>(fc.tab), wp = reinterpret_cast<uint64_t>(fc.wc);
					plan_put(fc.plan, 0, (uint32_t) tp, (uint32_t) (tp >> 32), (uint32_t) wp, (uint32_t) (wp >> 32),
							__uint_as_float(fc.wave_mask), fc.coeff, 0.f, 0.f);
					const uint64_t rs = reinterpret_cast<uint64_t>(row_s), rr = reinterpret_cast<uint64_t>(row_r);
					plan_put(fc.plan, 1, (uint32_t) rs, (uint32_t) (rs >> 32), c.tstride,
							(sd.start + off) | (fc.write_r ? 0x80000000u : 0u),
							fc.amp_scale, op_ptr(c, vs.carr_slot)->line[LINE_PAN].v0,
							__uint_as_float((uint32_t) rr), __uint_as_float((uint32_t) (rr >> 32)));
					__syncwarp();
				}
				if (fc.wave_mask & CTAB_FLAG) {
					/* coefficient planes: the plan is lowered (render_fast.cuh) */
					if (lane == 0 && !kept) plan_lower(fc.plan + PLAN_HDR, nrec);
					__syncwarp();
					SAUGEN_TRACE(22);      /* lowered */
					const bool aligned = ((sd.start + off) & 3u) == 0u;
					if ((fc.wave_mask & VERIFY_FLAG) && fc.keep_plans && g->plan_cache && vs.plan_gen && vs.plan_left &&
							vs.plan_so == fc.so && vs.plan_st == fc.st) {
						/* developer knob: the kept plan, had it been used, against the fresh one */
						const uint4 *kq = g->plan_cache + (size_t) lv * (18u + 2u * g->plan_cache_recs);
						const uint4 kh = __ldcg(kq);
						if (kh.x == vs.plan_gen) {
							uint32_t bad = kh.y != nrec || (kmax < 0x7fffu && kmax != vs.plan_left) ? 1u : 0u;
							for (uint32_t i = lane; i < 2u * nrec && !bad; i += 32) {
								uint4 x = __ldcg(kq + 1 + i);
								const uint4 y = lds128u(fc.plan + PLAN_HDR + 16u * i);
								uint32_t m = (i & 1u) ? 0xffffffffu : 0x00ffffffu;      /* (w1's top byte: the teams' levels) */
								if (!(i & 1u) && (x.x & 0xffu) == P_EXT) {
									x.y = plan_ext_pos(lds32(fc.plan + PLAN_HDR + 16u * i - PLAN_REC + 8), x.x >> 8);
									m = 0xffffffffu;
								}
								if (x.x != y.x || (x.y & m) != (y.y & m) || x.z != y.z || x.w != y.w) bad = 1u;
							}
							bad = __any_sync(FULL, bad != 0u) ? 1u : 0u;
							if (lane == 0) { atomicAdd(&g_team_dump[31], 1u); if (bad) atomicAdd(&g_team_dump[30], 1u); }
						}
					}
					uint32_t which = 0;
					if (!other && kept && kept != 0x100u) {
						which = aligned ? kept - 1u : 0u;
					} else if (!other) {
						if (lane == 0 && aligned && !(fc.wave_mask & NOFUSE_FLAG))
							which = fused_match(fc.plan, nrec, blockIdx.x == 0 && threadIdx.x == 0);
						which = __shfl_sync(FULL, which, 0);
					}
					SAUGEN_TRACE(23);      /* matched */
					if (aligned && !(fc.wave_mask & NOFUSE_FLAG) && !other) which_kept = which;
					/* few voices: the stretch is split along time over the voice's team of warps */
					const bool shared = fc.team && !other &&
						team_stretch(*fc.team, fc.sb, lane, fc.plan, nrec, vs.ops_cnt, span, which,
								g->team_cache ? g->team_cache + (size_t) lv * TEAM_SLOTS * g->team_cache_stride : nullptr,
								g->team_cache_stride,
								g->plan_cache ? g->plan_cache + (size_t) lv * (18u + 2u * g->plan_cache_recs) + 1u + 2u * g->plan_cache_recs : nullptr,
								teamP);
					if (shared) { }
					else if (other) run_block_lowered<true>(fc.sb, fc.plan, lane, 0u, span);
					else if (which) fused_run(which, fc.sb, fc.plan, lane, 0u, span);
					else run_block_lowered<false>(fc.sb, fc.plan, lane, 0u, span);
				} else {
					if (other) run_block_fast<false, true>(fc.sb, fc.plan, lane, fc.coeff, span, row_s, row_r, sd.start + off);
					else run_block_fast<false, false>(fc.sb, fc.plan, lane, fc.coeff, span, row_s, row_r, sd.start + off);
				}
				__syncwarp();
				SAUGEN_TRACE(24);          /* stretch rendered */
				steady_update(c.sops, g->code + vs.code_off, vs.code_len, nb, lane);
				__syncwarp();
				SAUGEN_TRACE(25);
				if ((fc.wave_mask & CTAB_FLAG) && fc.keep_plans && g->plan_cache) {
					const uint32_t left = kmax - nb;
					if (kept) {
						vs.plan_left = left;
						if (fc.team && teamP && !teamP_kept) {
							/* the teams' analysis of the kept plan, made in this stretch: kept too (levels -> the records) */
							uint4 *kq = g->plan_cache + (size_t) lv * (18u + 2u * g->plan_cache_recs);
							__syncwarp();
							for (uint32_t i = lane; i < 2u * nrec; i += 32) __stcg(kq + 1 + i, lds128u(fc.plan + PLAN_HDR + 16u * i));
							for (uint32_t i = lane; i < 17u; i += 32) __stcg(kq + 1 + 2u * g->plan_cache_recs + i, lds128u(fc.team->cmd + TC_INFO + 16u * i));
							if (lane == 0) __stcg(reinterpret_cast<uint32_t*>(kq) + 3, teamP);
							__syncwarp();
						}
					} else if (left && !other && nrec <= g->plan_cache_recs) {
						/* keep the plan for the stretches to come (a new generation: a call that is undone
						 * later -- run-ahead -- leaves the voice's older generation number behind, not this plan) */
						uint4 *kq = g->plan_cache + (size_t) lv * (18u + 2u * g->plan_cache_recs);
						const uint32_t gen = (__ldcg(kq).x & 0x7fffffffu) + 1u;
						__syncwarp();
						for (uint32_t i = lane; i < 2u * nrec; i += 32) __stcg(kq + 1 + i, lds128u(fc.plan + PLAN_HDR + 16u * i));
						if (fc.team && teamP)
							for (uint32_t i = lane; i < 17u; i += 32) __stcg(kq + 1 + 2u * g->plan_cache_recs + i, lds128u(fc.team->cmd + TC_INFO + 16u * i));
						if (lane == 0) __stcg(kq, make_uint4(gen, nrec, which_kept, fc.team ? teamP : 0u));
						__syncwarp();          /* every lane has read the plan before the area is used again */
						vs.plan_gen = gen; vs.plan_so = fc.so; vs.plan_st = fc.st; vs.plan_left = left;
					} else {
						vs.plan_gen = 0;
					}
				}
				vs.duration -= span;
				run_total += span;
				off += span - CHUNK;
				continue;
			}
			uint32_t clen = uend - off;
			if (clen > (uint32_t) CHUNK) clen = CHUNK;
			const uint32_t time = vs.duration < clen ? vs.duration : clen;
			vs.plan_gen = 0;           /* (the general interpreter may leave the voice in another shape) */
			__syncwarp();              /* (its len stacks overlay the plan area) */
			c.oc = off % REF_BLOCK;
			uint32_t rem0 = vs.duration;
			if (sd.len - off < rem0) rem0 = sd.len - off;
			if (REF_BLOCK - c.oc < rem0) rem0 = REF_BLOCK - c.oc;
			uint32_t out_len = 0;
			if (vs.code_len && op_ptr(c, vs.carr_slot)->time > 0)     /* run_voice, :833-846 */
				out_len = run_chunk(c, g->code + vs.code_off, vs.code_len, time, rem0,
						row_s, row_r, sd.start + off);
			__syncwarp();
			if (out_len && pan_mode == PAN_UNSET) {
				/* first rendered chunk of the segment decides (run_chunk wrote r if moving) */
				pan_mode = c.pan_dyn ? PAN_DYNAMIC :
					__float_as_uint(op_ptr(c, vs.carr_slot)->line[LINE_PAN].v0);
				c.write_r = c.pan_dyn;
				fc.write_r = c.write_r ? 1u : 0u;
			}
			vs.duration -= time;
			run_total += out_len;
		}
		if (lane == 0 && run_total) {
			/* frames this voice has run in the segment so far (units of a voice are
			 * rendered in order, by one warp at a time) */
			const uint32_t tot = __ldcg(&vsg->len) + run_total;
			__stcg(reinterpret_cast<uint2*>(vsg), make_uint2(tot, pan_mode));
			/* the maximum only grows: skip the atomic when it is already there */
			if (__ldcg(&cd->status[1 + si]) < tot) atomicMax(&cd->status[1 + si], tot);
		}
	}
	if (loaded) ops_store(c, loaded);
	SAUGEN_TRACE(26);
	if (lane == 0) {
		const uint32_t *w = reinterpret_cast<const uint32_t*>(&vs);
#pragma unroll
		for (uint32_t i = 0; i < sizeof(VoiceState) / 4; ++i)
			__stcg(reinterpret_cast<uint32_t*>(vsp) + i, w[i]);
		if (u1 == cd->nunits && !cd->more_launches && vs.duration != 0) atomicOr(&cd->status[0], 1u);
	}
}

__device__ __forceinline__ void render_body(const CallDesc *calls, uint32_t ncalls,
		const SegDesc *segs, const UnitDesc *units, uint32_t ntasks, const float *tables,
		const double *coefs, uint32_t wave_mask, uint32_t nbufs, uint32_t nslots_ops, uint32_t nplan, uint32_t warps_per_cta,
		uint32_t ticketed, uint32_t team_arg) {
	extern __shared__ __align__(128) unsigned char smem[];
	const uint32_t team = team_arg & 0xffu;        /* warps per voice in a CTA ... */
	const uint32_t multi = team_arg >> 8 ? team_arg >> 8 : 1u;   /* ... and CTAs per voice (render_team.cuh) */
	if (blockIdx.x == 0 && threadIdx.x == 0) g_trace_t0 = clock64();
	uint64_t *bar = reinterpret_cast<uint64_t*>(smem);
	float *tab = reinterpret_cast<float*>(smem + 128);
	const bool ctab = (wave_mask & CTAB_FLAG) != 0;
	const uint32_t nslots = __popc(wave_mask & 0xfffu);
	const uint32_t slot_bytes = ctab ? CTAB_WAVE_BYTES : TAB_STRIDE * (uint32_t) sizeof(float);
	unsigned char *warp_area = smem + 128 + nslots * slot_bytes;
	const uint32_t per_warp = warp_smem_bytes(nbufs, nslots_ops, nplan);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	/* stage the tables this launch needs: TMA bulk copies, one mbarrier */
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0 && nslots) {
		mbar_expect_tx(bar, nslots * (ctab ? CTAB_WAVE_BYTES : WAVE_LEN * (uint32_t) sizeof(float)));
		uint32_t slot = 0;
		for (uint32_t w = 0; w < NUM_WAVES; ++w) {
			if (!(wave_mask & (1u << w))) continue;
			if (ctab)
				tma_bulk_g2s(smem + 128 + slot * CTAB_WAVE_BYTES,
						reinterpret_cast<const unsigned char*>(coefs) + (size_t) w * CTAB_WAVE_BYTES,
						CTAB_WAVE_BYTES, bar);
			else
				tma_bulk_g2s(tab + slot * TAB_STRIDE + 4, tables + w * WAVE_LEN,
						WAVE_LEN * sizeof(float), bar);
			++slot;
		}
	}
	if (nslots) {
		mbar_wait(bar, 0);
		/* wrapped neighbours: lut[-1], lut[2048], lut[2049] */
		if (!ctab && threadIdx.x < nslots) {
			float *t = tab + threadIdx.x * TAB_STRIDE + 4;
			t[-1] = t[WAVE_LEN - 1];
			t[WAVE_LEN] = t[0];
			t[WAVE_LEN + 1] = t[1];
		}
		__syncthreads();
	}

	SAUGEN_TRACE(18);                  /* tables staged */
	Ctx c;
	/* team > 1 (no ticketed scheduler then): per team the voice's part, then one part per member (kernels.cu) */
	unsigned char *team_base = warp_area + (warp / team) * team_smem_bytes(nbufs, nslots_ops, nplan, team);
	unsigned char *member = team_base + team_lead_bytes(nslots_ops, nplan) +
		(warp % team) * team_member_bytes(nbufs, nslots_ops, nplan);
	if (team > 1u) {
		c.sops = reinterpret_cast<OpState*>(team_base);
		c.bufs = reinterpret_cast<float*>(member + nslots_ops * sizeof(OpState));
		c.stk_len = reinterpret_cast<uint32_t*>(team_base + nslots_ops * sizeof(OpState));
	} else {
		c.sops = reinterpret_cast<OpState*>(warp_area + warp * per_warp);
		c.bufs = reinterpret_cast<float*>(c.sops + nslots_ops);
		c.stk_len = reinterpret_cast<uint32_t*>(c.bufs + nbufs * BUF_FLOATS);
	}
	c.stk_rem = c.stk_len + MAX_NEST;
	c.stk_layer = c.stk_rem + MAX_NEST;
	c.tab = tab;                     /* staged float tables, or the coefficient planes */
	c.wc = reinterpret_cast<const WaveCoeffs*>(tables + NUM_WAVES * WAVE_LEN);
	c.wave_mask = wave_mask;
	c.lane = lane;
	FastCtx fc;
	fc.so = smem_u32(c.sops);
	fc.sb = smem_u32(c.bufs) + lane * 16;
	fc.st = smem_u32(tab);           /* staged float tables, or the coefficient tables */
	fc.tab = c.tab; fc.wc = c.wc;
	fc.wave_mask = wave_mask; fc.lane = lane;
	fc.plan = smem_u32(c.stk_len);   /* the plan overlays the len stacks */
	fc.plan_cap = nplan * 32u > STACK_BYTES ? nplan : STACK_BYTES / 32u;
	fc.team = nullptr;
	fc.keep_plans = !ticketed;

	if (!ticketed) {
		/* one warp (or, with few voices, a team of warps: render_team.cuh) renders every unit
		 * of one voice; task -> (call, voice) by binary search on task_base */
		const uint32_t per_cta = warps_per_cta / team;
		uint32_t team_i = warp / team;
		const uint32_t rank = warp - team_i * team;
		if (team == 1u) {
			/* Voices of a script tend to alternate in kind (C3: PM chain / range-FM) while a warp's scheduler
			 * is its number mod 4: every other PAIR of warps swaps its voices, so that the dearer kind does
			 * not pile up on two of the SM's four schedulers. */
			const uint32_t sw = team_i ^ (((team_i >> 1) ^ (team_i >> 2)) & 1u);
			if (sw < per_cta) team_i = sw;
		}
		if (team_i >= per_cta) return;
		const uint32_t task = (blockIdx.x / multi) * per_cta + team_i;
		if (task >= ntasks) return;
		uint32_t ci = 0, hi = ncalls;
		while (hi - ci > 1) {
			const uint32_t mid = (ci + hi) >> 1;
			if (calls[mid].task_base <= task) ci = mid; else hi = mid;
		}
		const CallDesc *cd = &calls[ci];
		TeamCtx tc;
		if (team > 1u) {
			tc.T = team; tc.rank = rank; tc.bar = 1u + team_i;
			tc.K = multi; tc.part = blockIdx.x % multi;
			tc.hdr = nullptr; tc.mail = nullptr;
			tc.plan_bytes = warp_plan_bytes(nplan); tc.max_ops = nslots_ops;
			if (multi > 1u) {
				const GenDesc *gd = cd->gen;
				const uint32_t lvt = task - cd->task_base;
				tc.hdr = gd->team_hdr + (size_t) lvt * 16u;
				tc.mail = gd->team_mail + (size_t) lvt * gd->team_mail_stride;
			}
			tc.so_a = fc.so; tc.so_b = smem_u32(member);
			tc.plan_x = smem_u32(member) + nslots_ops * (uint32_t) sizeof(OpState) + nbufs * BUF_FLOATS * (uint32_t) sizeof(float);
			tc.per_warp = team_member_bytes(nbufs, nslots_ops, nplan);
			tc.cmd = fc.plan + warp_plan_bytes(nplan);
			tc.lead_so = fc.so; tc.lead_plan = fc.plan;
			if (rank == 0u && lane == 0) sts32(tc.cmd + TC_SYNCS, 0u);     /* (before the team's first barrier) */
			if (tc.part) { team_remote(tc, fc.sb, lane); return; }
			if (rank) { team_helper(tc, fc.sb, lane); return; }
			fc.team = &tc;
		}
		render_units(c, fc, cd, segs, units, task - cd->task_base, 0, cd->nunits);
		if (team > 1u) team_dismiss(tc, lane);
		return;
	}
	const CallDesc *cd = &calls[0];
	const GenDesc *g = cd->gen;
	const uint32_t nlv = g->voice_end - g->voice_begin;
	if (ticketed == 2) {
		/* Balanced: more voices than resident warps, all of them alike.  The
		 * (voice, unit) items of the call, voice-major, are cut into one contiguous
		 * range per warp of a grid that is resident all at once, so every warp gets
		 * the same amount of work (+-1 unit) and there is no second, partly filled
		 * wave.  A range covers the tail of one voice, whole voices, and the head
		 * of another.  The head comes FIRST (it depends on nothing), the tail LAST:
		 * it continues what the previous warp rendered as its first action, handed
		 * over through L2 (progress[], release / acquire).  Warp ranks are taken
		 * from a counter, so the warp holding the previous rank has already started. */
		const uint32_t U = cd->nunits;
		const uint64_t items = (uint64_t) nlv * U;
		const uint64_t S = (uint64_t) gridDim.x * warps_per_cta;
		uint32_t rank = 0;
		if (lane == 0) rank = atomicAdd(g->ticket, 1u);
		rank = __shfl_sync(FULL, rank, 0);
		const uint64_t begin = rank * items / S, end = (rank + 1ull) * items / S;
		if (begin >= end) return;
		const uint32_t vA = (uint32_t) (begin / U), uA = (uint32_t) (begin - (uint64_t) vA * U);
		const uint32_t vB = (uint32_t) ((end - 1) / U), uB = (uint32_t) (end - (uint64_t) vB * U);
		auto publish = [&](uint32_t lv, uint32_t u) {
			__threadfence();
			__syncwarp();
			if (lane == 0) *(volatile uint32_t*) (g->progress + lv) = u;
		};
		auto await = [&](uint32_t lv, uint32_t u) {
			if (lane == 0) {
				volatile uint32_t *pr = g->progress + lv;
				while (*pr != u) __nanosleep(64);
				__threadfence();
			}
			__syncwarp();
		};
		if (vA == vB) {
			if (uA) await(vA, uA);
			render_units(c, fc, cd, segs, units, vA, uA, uB);
			if (uB < U) publish(vA, uB);
			return;
		}
		uint32_t v_hi = vB;                    /* whole voices are [v_lo, v_hi] */
		if (uB < U) {
			render_units(c, fc, cd, segs, units, vB, 0, uB);
			publish(vB, uB);
			--v_hi;
		}
		const uint32_t v_lo = uA ? vA + 1 : vA;
		for (uint32_t v = v_lo; v <= v_hi; ++v)
			render_units(c, fc, cd, segs, units, v, 0, U);
		if (uA) {
			await(vA, uA);
			render_units(c, fc, cd, segs, units, vA, uA, U);
		}
		return;
	}
	/* Ticketed: a persistent grid hands out (unit, voice) pairs in time order, so
	 * that SMs stay evenly loaded when there are more voices than resident warps.
	 * Unit u of a voice may start once its unit u-1 is done (progress[], release /
	 * acquire through global memory); the holder of every earlier ticket is
	 * already running, so the wait always ends. */
	const uint32_t total = nlv * cd->nunits;
	for (;;) {
		uint32_t t = 0;
		if (lane == 0) t = atomicAdd(g->ticket, 1u);
		t = __shfl_sync(FULL, t, 0);
		if (t >= total) break;
		const uint32_t u = t / nlv, lv = t - u * nlv;
		if (lane == 0) {
			volatile uint32_t *pr = g->progress + lv;
			while (*pr != u) __nanosleep(32);
			__threadfence();
		}
		__syncwarp();
		render_units(c, fc, cd, segs, units, lv, u, u + 1);
		__threadfence();
		__syncwarp();
		if (lane == 0) *(volatile uint32_t*) (g->progress + lv) = u + 1;
	}
}

__global__ void __launch_bounds__(256, 2)
render_kernel(const CallDesc *calls, uint32_t ncalls, const SegDesc *segs, const UnitDesc *units,
		uint32_t ntasks, const float *tables, const double *coefs, uint32_t wave_mask, uint32_t nbufs,
		uint32_t nslots_ops, uint32_t nplan, uint32_t warps_per_cta, uint32_t ticketed, uint32_t team) {
	render_body(calls, ncalls, segs, units, ntasks, tables, coefs, wave_mask, nbufs, nslots_ops, nplan,
			warps_per_cta, ticketed, team);
}

/* same body for CTAs of up to 32 warps (64 registers), one per SM (coefficient-table
 * mode: the planes take 48 KiB per wave, so one large CTA shares them among all the
 * warps an SM can hold -- the path is latency-bound, resident warps are what counts) */
__global__ void __launch_bounds__(WIDE_WARPS * 32, 1)
render_kernel_wide(const CallDesc *calls, uint32_t ncalls, const SegDesc *segs, const UnitDesc *units,
		uint32_t ntasks, const float *tables, const double *coefs, uint32_t wave_mask, uint32_t nbufs,
		uint32_t nslots_ops, uint32_t nplan, uint32_t warps_per_cta, uint32_t ticketed, uint32_t team) {
	render_body(calls, ncalls, segs, units, ntasks, tables, coefs, wave_mask, nbufs, nslots_ops, nplan,
			warps_per_cta, ticketed, team);
}
