/* render_interp.cuh -- part of kernels.cu (one translation unit; included inside namespace saugen):
 * the general bytecode interpreter: one 128-sample chunk of one voice, any operator type, any state. */
#pragma once

/* ---- bytecode interpreter: one chunk of one voice ----------------------- */

__device__ __noinline__ uint32_t run_chunk(Ctx &c, const Instr *code, uint32_t code_len, uint32_t time,
		uint32_t rem0, float *row_s, float *row_r, uint32_t frame) {
	c.sp = 0;
	c.frame = frame;
	if (c.lane == 0) { c.stk_len[0] = time; c.stk_rem[0] = rem0; c.stk_layer[0] = 0; }
	__syncwarp();
	c.pma_flag = false; c.pan_dyn = false;
	c.last_len = 0; c.last_rem = 0;
	uint32_t pc = 0;
	while (pc < code_len) {
		const uint4 raw = __ldg(reinterpret_cast<const uint4*>(code + pc));
		Instr in;
		memcpy(&in, &raw, sizeof(in));
		++pc;
		const uint32_t n = c.stk_len[c.sp];
		const uint32_t i0 = c.lane * SPL;
		switch (in.opcode) {
		case I_WLEAF: wop<true, true>(c, in, pc); break;
		case I_WHEAD: wop<true, false>(c, in, pc); break;
		case I_WTAIL: wop<false, true>(c, in, pc); break;
		case I_ENTER: {                                            /* generator.c:675-698 */
			const OpState *o = op_ptr(c, in.op);
			const uint32_t flags = o->flags, t = o->time;
			uint32_t rem = c.stk_rem[c.sp];
			if (!(flags & ON_TIME_INF) && t < rem) rem = t;
			const uint32_t len = rem < n ? rem : n;
			const uint32_t layer = (in.flags & F_LAYER) ? 1u :
				((in.flags & F_LAYER_PMA) ? (c.pma_flag ? 1u : 0u) : 0u);
			++c.sp;
			if (c.lane == 0) { c.stk_len[c.sp] = len; c.stk_rem[c.sp] = rem; c.stk_layer[c.sp] = layer; }
			__syncwarp();
			if (len == 0) pc = in.aux;     /* nothing to render: go to the LEAVE */
			break; }
		case I_LEAVE: {                                            /* generator.c:716-728 */
			OpState *o = op_ptr(c, in.op);
			const uint32_t len = n, layer = c.stk_layer[c.sp];
			c.last_len = len; c.last_rem = c.stk_rem[c.sp];
			--c.sp;
			leave_eval(c, o, in.a, len, c.stk_len[c.sp], layer);
			__syncwarp();
			if (c.wave_mask & TAP_FLAG) { tap_store(c, in.op, in.a, c.stk_len[c.sp]); __syncwarp(); }
			break; }
		case I_ZERO:
			*B4(c, in.a) = make_float4(0.f, 0.f, 0.f, 0.f);
			__syncwarp();
			break;
		case I_LINE: {
			OpState *o = op_ptr(c, in.op);
			if (in.d) {
				const bool has_mul = in.b != NO_BUF;
				float out[SPL], m[SPL];
				bool done = false;
				if (n == (uint32_t) CHUNK) {
					const LineRegs r = line_load(o, in.c);
					if (has_mul) ld4(c, in.b, m);
					__syncwarp();
					done = line_eval_full(c.oc, c.lane, o, in.c, r, has_mul ? m : nullptr, out);
				}
				if (done) st4(c, in.a, out);
				else *B4(c, in.a) = line_eval_any(c.oc, c.lane, o, in.c,
						has_mul ? c.bufs + in.b * CHUNK : nullptr, n, c.stk_rem[c.sp]);
			} else {
				line_skip(c.oc, c.lane, o, in.c, n);
			}
			__syncwarp();
			break; }
		case I_RANGE: {                                            /* generator.c:465-467 */
			float4 p = *B4(c, in.a);
			const float4 r = *B4(c, in.b), m = *B4(c, in.c);
			if (i0 + 0 < n) p.x += (r.x - p.x) * m.x;
			if (i0 + 1 < n) p.y += (r.y - p.y) * m.y;
			if (i0 + 2 < n) p.z += (r.z - p.z) * m.z;
			if (i0 + 3 < n) p.w += (r.w - p.w) * m.w;
			*B4(c, in.a) = p;
			__syncwarp();
			break; }
		case I_PHASOR: {
			float f[SPL], pm[SPL], fpm[SPL];
			uint32_t ph[SPL];
			ld4(c, in.b, f);
			if (in.c != NO_BUF) ld4(c, in.c, pm);
			if (in.d != NO_BUF) ld4(c, in.d, fpm);
			{
				OpState *o = op_ptr(c, in.op);
				const uint32_t phase0 = o->i0;
				__syncwarp();
				phasor_eval<false>(c, o, phase0, f, in.c != NO_BUF ? pm : nullptr,
						in.d != NO_BUF ? fpm : nullptr, n, ph);
			}
			*U4(c, in.a) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
			__syncwarp();
			break; }
		case I_PMA: {                                              /* generator.c:485-490 */
			OpState *o = op_ptr(c, in.op);
			const bool run = pma_decide(c, o);
			if (run) *B4(c, in.a) = line_eval_any(c.oc, c.lane, o, LINE_PMA, nullptr, n, c.stk_rem[c.sp]);
			else line_skip(c.oc, c.lane, o, LINE_PMA, n);
			c.pma_flag = run;
			__syncwarp();
			break; }
		case I_WOSC:
			if (n) {
				OpState *o = op_ptr(c, in.op);
				if ((in.flags & F_HAS_APMODS) || c.pma_flag) {
					wosc_selfmod(cold(c), o, reinterpret_cast<const uint32_t*>(c.bufs + in.b * CHUNK),
							c.bufs + in.c * CHUNK, c.bufs + in.a * CHUNK, n);
				} else {
					*B4(c, in.a) = wosc_eval_any(cold(c), o, *U4(c, in.b), n);
				}
			}
			__syncwarp();
			break;
		case I_CYCLOR:
			cyclor_fill(c, in, n);
			__syncwarp();
			break;
		case I_RASG:
			if (n) rasg_run(c, in, n, c.oc + c.stk_rem[c.sp]);
			__syncwarp();
			break;
		case I_NOISE:
			noise_run(c, in, n);
			__syncwarp();
			break;
		case I_MIX: {                                              /* generator.c:384-440 */
			float x[SPL] = {1.f, 1.f, 1.f, 1.f}, a[SPL];
			if (in.b != NO_BUF) ld4(c, in.b, x);
			ld4(c, in.c, a);
			mix_eval<false>(c, in.a, x, a, n, c.stk_layer[c.sp], (in.flags & F_WAVEENV) != 0);
			__syncwarp();
			break; }
		case I_VPAN: {                                             /* generator.c:756-762 */
			/* the voice-level part runs over the carrier's out_len */
			__syncwarp();              /* every lane has read this iteration's length */
			if (c.lane == 0) { c.stk_len[0] = c.last_len; c.stk_rem[0] = c.last_rem; }
			__syncwarp();
			if (c.last_len == 0) return 0;
			OpState *po = op_ptr(c, in.op);
			const bool run = in.d || (LM_FLAGS(po->lmeta[LINE_PAN]) & SAUABI_LINEP_GOAL);
			__syncwarp();
			if (run) *B4(c, in.a) = line_eval_any(c.oc, c.lane, po, LINE_PAN, nullptr, c.last_len, c.last_rem);
			else line_skip(c.oc, c.lane, po, LINE_PAN, c.last_len);
			c.pan_dyn = run;
			__syncwarp();
			break; }
		case I_VOUT: {                                             /* generator.c:772-786 */
			const uint32_t vn = c.stk_len[0];
			const float amp_scale = c.g->amp_scale;
			const float4 sv = *B4(c, in.a);
			float4 pv;
			if (c.pan_dyn) pv = *B4(c, in.b);
			else { const float p = op_ptr(c, in.op)->line[LINE_PAN].v0; pv = make_float4(p, p, p, p); }
			float4 s, r;
			s.x = sv.x * amp_scale; r.x = s.x * pv.x;
			s.y = sv.y * amp_scale; r.y = s.y * pv.y;
			s.z = sv.z * amp_scale; r.z = s.z * pv.z;
			s.w = sv.w * amp_scale; r.w = s.w * pv.w;
			const bool wr = c.write_r || c.pan_dyn;      /* see VoiceSeg */
			/* row_s / row_r: this voice's piece of frame tile 0 (device_types.h:ROW_TILE) */
			const uint32_t fl = frame + i0;
			if (i0 + 3 < vn && (fl & 3u) == 0) {
				const size_t at = row_index(fl, c.tstride);
				__stcs(reinterpret_cast<float4*>(row_s + at), s);   /* coalesced 128-bit stores */
				if (wr) __stcs(reinterpret_cast<float4*>(row_r + at), r);
			} else {
				const float sa[4] = {s.x, s.y, s.z, s.w}, ra[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
				for (int k = 0; k < 4; ++k)
					if (i0 + k < vn) {
						const size_t at = row_index(fl + k, c.tstride);
						row_s[at] = sa[k];
						if (wr) row_r[at] = ra[k];
					}
			}
			return vn; }
		case I_END:
		default:
			return c.stk_len[0];
		}
	}
	return 0;
}
