/* kernels.cu -- sm_100a kernels of the saugns generator back end.
 *
 *   render_kernel : one warp per (call, voice).  The warp owns its voice's
 *     timeline for the call: it applies the voice's due events
 *     (handle_event, sau/generator.c:348-377), then renders every inter-event
 *     segment by interpreting the voice's bytecode (the flattened run_block
 *     recursion, generator.c:448-729): in 128-sample chunks through the
 *     general interpreter (run_chunk: any operator, any state), or -- for
 *     stretches of whole 1024-sample blocks in which nothing changes shape,
 *     almost all of a render -- from a PLAN the warp compiles once per stretch
 *     into its shared memory (steady_plan -> run_block_fast -> steady_update).
 *     Each lane owns 4 consecutive samples of a chunk; integer phase
 *     accumulation is a warp shuffle scan over the rounded increments, or a
 *     closed form when the frequency is uniform (bit-exact with
 *     sauPhasor_fill / sauCyclor_fill either way); self-PM operators run as a
 *     serial loop on one lane with their state in registers.  Wave tables --
 *     or their per-index cubic coefficients -- are staged into shared memory
 *     with TMA bulk copies (cp.async.bulk + mbarrier).  The carrier chunk,
 *     scaled, goes to the voice's 512-byte piece of the frame tile in HBM
 *     (device_types.h:ROW_TILE) with 128-bit streaming stores.
 *   mix_kernel : fused epilogue.  One CTA per frame tile, one thread per
 *     output frame; the tile's voice pieces stream through a TMA-filled ring
 *     and are summed in voice order (same float summation order as mix_add,
 *     generator.c:749-788), clamped and rounded to int16
 *     (mix_write_stereo/mono, generator.c:795-825).
 *
 * Compiled with -fmad=false: every float/double operation is a separately
 * rounded IEEE operation, in the order sau_arith.h states.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "device_types.h"
#include "sau_arith.h"
#include "../../include/sau_program_abi.h"

namespace saugen {

#define FULL 0xffffffffu

/* per-warp shared memory: operator-state cache, work buffers, len stack */
constexpr int FAST_NS = 4;            // samples per lane in the steady-block fast path (8 measured
                                      // slower: its registers allow 16 resident warps per SM, 4 allows 32)
constexpr uint32_t BUF_FLOATS = 32 * (FAST_NS > SPL ? FAST_NS : SPL);   // per work buffer
/* The len stacks (general interpreter, live inside one chunk only) and the block
 * plan (fast path, live across one steady block, 32 bytes per record) share one
 * area: whichever is larger. */
constexpr int WIDE_WARPS = 28;        // most warps of a render_kernel_wide CTA: 72 registers each
constexpr uint32_t STACK_BYTES = 3 * MAX_NEST * (uint32_t) sizeof(uint32_t);
__host__ __device__ inline uint32_t warp_smem_bytes(uint32_t nbufs, uint32_t nslots, uint32_t nplan) {
	const uint32_t plan_bytes = nplan * 32u;
	return nslots * (uint32_t) sizeof(OpState) + nbufs * BUF_FLOATS * (uint32_t) sizeof(float) +
		(plan_bytes > STACK_BYTES ? plan_bytes : STACK_BYTES);
}

/* ---- TMA 1-D bulk copy + mbarrier (PTX) --------------------------------- */

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
	return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
			:: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
		uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
			" [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra WAIT_DONE;\n"
		"bra WAIT_LOOP;\n"
		"WAIT_DONE:\n"
		"}\n" :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}

/* ---- shared-window loads / stores by 32-bit address ------------------------ */

__device__ __forceinline__ float4 lds128(uint32_t a) {
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
			: "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
	return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
			: "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
	return v;
}
__device__ __forceinline__ double2 lds128d(uint32_t a) {
	double2 v;
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
	return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t a) {
	float2 v;
	asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
	return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ float lds32f(uint32_t a) {
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
	asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
			:: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
	asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
/* ---- per-warp interpreter context --------------------------------------- */

struct Ctx {
	float *bufs;               // shared: nbufs x CHUNK floats of this warp
	uint32_t *stk_len;         // shared: MAX_NEST entries each
	uint32_t *stk_rem;
	uint32_t *stk_layer;
	OpState *sops;             // shared: this voice's operator states
	const float *tab;          // shared: staged wave tables (or coefficient planes, CTAB_FLAG)
	const WaveCoeffs *wc;      // global
	const GenDesc *g;          // global
	OpState *gops;             // global operator states
	const uint32_t *prog_ops;  // global: slot -> operator id of the current program
	float coeff;               // g->coeff
	uint32_t wave_mask;        // tables staged by this launch
	uint32_t oc;               // chunk offset inside the reference's 1024-block
	int lane;
	int sp;
	bool pma_flag, pan_dyn;
	bool write_r;              // the segment's pan moves: r rows are written (VoiceSeg)
	uint32_t tstride;          // g->row_stride: floats between frame tiles of the carrier rows
	uint32_t last_len, last_rem;
};

/* Instr::op is a slot of the voice program's operator list. */
__device__ __forceinline__ OpState *op_ptr(const Ctx &c, uint32_t slot) {
	return c.sops + slot;
}
__device__ __forceinline__ float4 *B4(const Ctx &c, uint32_t i) {
	return reinterpret_cast<float4*>(c.bufs + i * CHUNK) + c.lane;
}
__device__ __forceinline__ uint4 *U4(const Ctx &c, uint32_t i) {
	return reinterpret_cast<uint4*>(c.bufs + i * CHUNK) + c.lane;
}
__device__ __forceinline__ void ld4(const Ctx &c, uint32_t buf, float v[SPL]) {
	const float4 t = *B4(c, buf);
	v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void st4(const Ctx &c, uint32_t buf, const float v[SPL]) {
	*B4(c, buf) = make_float4(v[0], v[1], v[2], v[3]);
}
/* Staged tables: slot stride TAB_STRIDE floats, table at +4 (16-byte aligned
 * for the bulk copy), lut[-1] at +3 and lut[2048], lut[2049] after it, so the
 * four Hermite taps of an index are consecutive without masking. */
constexpr uint32_t TAB_STRIDE = WAVE_LEN + 8;
/* Coefficient-table mode (flag in the top bit of the wave mask a launch carries):
 * shared memory holds, for every wave the launch uses, the per-index cubic
 * coefficients of sauWave_get_herp in double precision (see coef_kernel) instead
 * of the float tables; every table evaluation, hot or rare, goes through them. */
constexpr uint32_t CTAB_FLAG = 0x80000000u;
constexpr uint32_t CTAB_WAVE_BYTES = WAVE_LEN * 24;       // {c3,c2} double plane + {c1,c0} float plane
constexpr uint32_t CTAB_PLANE_BYTES = WAVE_LEN * 16;      // offset of the float plane
/* What the out-of-line (rare path) functions need, passed by value. */
struct ColdCtx {
	const float *tab;
	const WaveCoeffs *wc;
	uint32_t wave_mask;
	int lane;
};
__device__ __forceinline__ ColdCtx cold(const Ctx &c) {
	ColdCtx k; k.tab = c.tab; k.wc = c.wc; k.wave_mask = c.wave_mask; k.lane = c.lane;
	return k;
}
/* A staged wave: the float table (wrapped neighbours around it), or its
 * coefficient planes. */
struct WaveRef {
	const void *p;
	bool ct;
};
template <typename C>
__device__ __forceinline__ WaveRef wave_ref(const C &c, uint32_t wave) {
	const uint32_t slot = __popc(c.wave_mask & ((1u << wave) - 1u));
	WaveRef r;
	r.ct = (c.wave_mask & CTAB_FLAG) != 0;
	if (r.ct) r.p = reinterpret_cast<const unsigned char*>(c.tab) + (size_t) slot * CTAB_WAVE_BYTES;
	else r.p = c.tab + slot * TAB_STRIDE + 4;
	return r;
}
/* sauWave_get_herp (wave.h:127-141) on either form; poly_out / c0_out as sau::herp */
__device__ __forceinline__ double herp_ref(const WaveRef &w, uint32_t phase, double *poly_out,
		double *c0_out) {
	if (w.ct) {
		const uint32_t ind = phase >> sau::WAVE_SLENBITS;
		const double2 hi = reinterpret_cast<const double2*>(w.p)[ind];
		const float2 lo = reinterpret_cast<const float2*>(
				reinterpret_cast<const unsigned char*>(w.p) + CTAB_PLANE_BYTES)[ind];
		const double p = sau::herp_horner(hi.x, hi.y, (double) lo.x, phase);
		if (poly_out) { *poly_out = p; *c0_out = (double) lo.y; }
		return p + (double) lo.y;
	}
	/* staged float table: taps lut[ind-1 .. ind+2] are consecutive, no masking */
	const float *t = reinterpret_cast<const float*>(w.p) - 1 + (phase >> sau::WAVE_SLENBITS);
	const float s0 = t[0], s1 = t[1], s2 = t[2], s3 = t[3];
	const double p = sau::herp_poly(s0, s1, s2, s3, phase);
	if (poly_out) { *poly_out = p; *c0_out = (double) s1; }
	return p + (double) s1;
}
__device__ __forceinline__ uint32_t scan_incl_u32(uint32_t v, int lane) {
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t y = __shfl_up_sync(FULL, v, d);
		if (lane >= d) v += y;
	}
	return v;
}
__device__ __forceinline__ uint64_t scan_incl_u64(uint64_t v, int lane) {
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint64_t y = __shfl_up_sync(FULL, v, d);
		if (lane >= d) v += y;
	}
	return v;
}

/* ---- sauLine state machine (sau/line.c:349-473) on a chunk -------------- */

__device__ __forceinline__ void line_advance(uint32_t &pos, uint32_t end, uint32_t &flags,
		uint32_t n, bool &expired) {                             /* line.c:385-398 */
	if (pos < end) {
		uint32_t l = end - pos;
		if (l > n) l = n;
		pos += l;
	}
	expired = false;
	if (pos >= end) {
		pos = 0;
		flags &= ~SAUABI_LINEP_TIME;
		expired = true;
	}
}

/* Line state in registers (every lane holds the same copy). */
struct LineRegs {
	float v0, vt, inv;
	uint32_t pos, end, meta;
};
__device__ __forceinline__ LineRegs line_load(const OpState *o, int li) {
	LineRegs r;
	const float4 t = *reinterpret_cast<const float4*>(&o->line[li]);
	r.v0 = t.x; r.vt = t.y; r.pos = __float_as_uint(t.z); r.end = __float_as_uint(t.w);
	r.meta = o->lmeta[li];
	r.inv = o->linv[li];
	return r;
}

/* sauLine_run(line, out, n, mulbuf) -- line.c:417-445 -- any n, any state.
 * mulbuf: shared-memory buffer of ratio multipliers or nullptr; rem: samples
 * the visit still has in the reference's 1024-block (for gcc's cub tail).
 * Ends with the state write-back by lane 0; the caller syncs the warp. */
__device__ __noinline__ float4 line_eval_any(uint32_t oc, int lane, OpState *o, int li,
		const float *mulbuf, uint32_t n, uint32_t rem) {
	float out[SPL] = {0.f, 0.f, 0.f, 0.f};
	LineState *ls = &o->line[li];
	float v0 = ls->v0, vt = ls->vt;
	uint32_t pos = ls->pos, end = ls->end;
	const uint32_t meta = o->lmeta[li];
	uint32_t type = LM_TYPE(meta), flags = LM_FLAGS(meta);
	/* The reference advances a line once per 1024-block: when the position
	 * reaches `end` (wrap, or goal reached) the rest of that block is not
	 * counted (line.c:385-398,426-443).  blk_done carries that across our
	 * 128-sample chunks so pos/flags stay bit-identical at any later event. */
	uint32_t blk_done = oc == 0 ? 0u : LM_BLK(meta);
	const bool has_mul = (mulbuf != nullptr);
	float m[SPL] = {1.f, 1.f, 1.f, 1.f};
	if (has_mul) {
		const float4 t = reinterpret_cast<const float4*>(mulbuf)[lane];
		m[0] = t.x; m[1] = t.y; m[2] = t.z; m[3] = t.w;
	}
	const uint32_t i0 = lane * SPL;
	if (!(flags & SAUABI_LINEP_GOAL)) {
		if (!blk_done) {
			bool ex;
			line_advance(pos, end, flags, n, ex);
			if (ex) blk_done = 1;
		}
		const bool um = has_mul && (flags & SAUABI_LINEP_STATE_RATIO);
#pragma unroll
		for (int k = 0; k < SPL; ++k) out[k] = um ? v0 * m[k] : v0;
	} else {
		bool fillmul = has_mul;                                   /* sauLine_get, line.c:349-378 */
		if (flags & SAUABI_LINEP_GOAL_RATIO) {
			if (!(flags & SAUABI_LINEP_STATE_RATIO)) {
				if (has_mul) v0 = v0 / mulbuf[0];
				flags |= SAUABI_LINEP_STATE_RATIO;
			}
		} else {
			if (flags & SAUABI_LINEP_STATE_RATIO) {
				if (has_mul) v0 = v0 * mulbuf[0];
				flags &= ~SAUABI_LINEP_STATE_RATIO;
			}
			fillmul = false;
		}
		uint32_t flen = 0;
		if (pos < end) { flen = end - pos; if (flen > n) flen = n; }
		if (flen > 0) {
			sau::LineFill f = sau::line_fill_setup((int) type, v0, vt, pos, end);
			/* gcc's scalar tail of sauLine_fill_cub: the last element of an
			 * odd-length fill call, counted in the reference's 1024-block. */
			uint32_t tail_idx = 0xffffffffu;
			if (f.type == sau::L_cub) {
				uint32_t F = end - pos;
				if (F > rem) F = rem;
				if (F <= (uint32_t) CHUNK && ((oc + F) & 1u)) tail_idx = F - 1;
			}
#pragma unroll
			for (int k = 0; k < SPL; ++k) {
				uint32_t idx = i0 + k;
				float v = sau::line_fill_at(f, idx, idx == tail_idx);
				out[k] = fillmul ? v * m[k] : v;
			}
		}
		pos += flen;
		if (pos >= end) {
			v0 = vt;
			pos = 0;
			blk_done = 1;
			flags &= ~(SAUABI_LINEP_GOAL | SAUABI_LINEP_GOAL_RATIO | SAUABI_LINEP_TIME);
			const bool um = has_mul && (flags & SAUABI_LINEP_STATE_RATIO);
#pragma unroll
			for (int k = 0; k < SPL; ++k)
				if (i0 + k >= flen) out[k] = um ? v0 * m[k] : v0;
		}
	}
	__syncwarp();   /* every lane has read the state before lane 0 rewrites it */
	if (lane == 0) {
		ls->v0 = v0; ls->pos = pos;
		o->lmeta[li] = LM_PACK(type, flags, blk_done);
	}
	return make_float4(out[0], out[1], out[2], out[3]);
}

/* The common cases of the above on a FULL chunk (n == CHUNK), from registers:
 * no goal (hold v0), or a goal whose trajectory covers the whole chunk with
 * no ratio reconciliation due.  Returns false (nothing touched) otherwise.
 * Lane 0 writes the state back; the caller has synced after line_load and
 * syncs again before anything re-reads the state. */
template <int TYPE>
__device__ __forceinline__ void line_fill4(const sau::LineFill &f, uint32_t i0, float out[SPL]) {
	sau::LineFill g = f;
	g.type = TYPE;
#pragma unroll
	for (int k = 0; k < SPL; ++k) out[k] = sau::line_fill_at(g, i0 + k, false);
}
__device__ __forceinline__ bool line_eval_full(uint32_t oc, int lane, OpState *o, int li,
		const LineRegs &r, const float *m /* SPL multipliers or nullptr */, float out[SPL]) {
	const uint32_t type = LM_TYPE(r.meta);
	uint32_t flags = LM_FLAGS(r.meta);
	const uint32_t blk0 = LM_BLK(r.meta);
	uint32_t blk = oc == 0 ? 0u : blk0;
	if (!(flags & SAUABI_LINEP_GOAL)) {
		uint32_t pos = r.pos;
		if (!blk) {
			bool ex;
			line_advance(pos, r.end, flags, CHUNK, ex);
			if (ex) blk = 1;
		}
		const bool um = m && (flags & SAUABI_LINEP_STATE_RATIO);
#pragma unroll
		for (int k = 0; k < SPL; ++k) out[k] = um ? r.v0 * m[k] : r.v0;
		if (lane == 0) {
			if (pos != r.pos) o->line[li].pos = pos;
			const uint32_t meta = LM_PACK(type, flags, blk);
			if (meta != r.meta) o->lmeta[li] = meta;
		}
		return true;
	}
	const bool gr = (flags & SAUABI_LINEP_GOAL_RATIO) != 0, sr = (flags & SAUABI_LINEP_STATE_RATIO) != 0;
	if (gr != sr) return false;
	if (!(r.pos < r.end && r.end - r.pos > (uint32_t) CHUNK)) return false;
	/* line_fill_setup with the reciprocal kept in the state */
	sau::LineFill f;
	int t = (int) type;
	if (t == sau::L_exp) t = (r.v0 > r.vt) ? sau::L_xpe : sau::L_lge;
	else if (t == sau::L_log) t = (r.v0 < r.vt) ? sau::L_xpe : sau::L_lge;
	f.type = t;
	f.v0 = r.v0; f.vt = r.vt; f.pos = r.pos;
	f.adj_pos = (int32_t) (r.pos - (r.end / 2));
	f.inv = r.inv;
	f.vm = (r.v0 + r.vt) * 0.5f;
	f.vd = r.vt - r.v0;
	f.c = 0.f;
	const uint32_t i0 = lane * SPL;
	switch (t) {
	default:
	case sau::L_sah: line_fill4<sau::L_sah>(f, i0, out); break;
	case sau::L_lin: f.c = f.vd * f.inv; line_fill4<sau::L_lin>(f, i0, out); break;
	case sau::L_cos: line_fill4<sau::L_cos>(f, i0, out); break;
	case sau::L_xpe: f.c = r.v0 - r.vt; line_fill4<sau::L_xpe>(f, i0, out); break;
	case sau::L_lge: line_fill4<sau::L_lge>(f, i0, out); break;
	case sau::L_sqe: f.c = r.v0 - r.vt; line_fill4<sau::L_sqe>(f, i0, out); break;
	case sau::L_cub: f.inv = -2.f * f.inv; f.c = (r.v0 - r.vt) * 0.5f; line_fill4<sau::L_cub>(f, i0, out); break;
	case sau::L_smo: line_fill4<sau::L_smo>(f, i0, out); break;
	case sau::L_uwh: f.c = f.vd * (0.5f / 2147483648.f); line_fill4<sau::L_uwh>(f, i0, out); break;
	case sau::L_ncl: line_fill4<sau::L_ncl>(f, i0, out); break;
	case sau::L_nhl: line_fill4<sau::L_nhl>(f, i0, out); break;
	}
	if (m && gr) {
#pragma unroll
		for (int k = 0; k < SPL; ++k) out[k] = out[k] * m[k];
	}
	if (lane == 0) {
		o->line[li].pos = r.pos + CHUNK;
		if (blk != blk0) o->lmeta[li] = LM_PACK(type, flags, blk);
	}
	return true;
}

/* sauLine_skip -- line.c:456-473 */
__device__ __noinline__ void line_skip(uint32_t oc, int lane, OpState *o, int li, uint32_t n) {
	if (lane != 0) return;
	LineState *ls = &o->line[li];
	const uint32_t meta = o->lmeta[li];
	uint32_t pos = ls->pos, end = ls->end, flags = LM_FLAGS(meta);
	uint32_t blk_done = oc == 0 ? 0u : LM_BLK(meta);
	if (!blk_done) {
		bool ex;
		line_advance(pos, end, flags, n, ex);
		if (ex) {
			blk_done = 1;
			if (flags & SAUABI_LINEP_GOAL) {
				ls->v0 = ls->vt;
				if (flags & SAUABI_LINEP_GOAL_RATIO) flags |= SAUABI_LINEP_STATE_RATIO;
				else flags &= ~SAUABI_LINEP_STATE_RATIO;
				flags &= ~(SAUABI_LINEP_GOAL | SAUABI_LINEP_GOAL_RATIO);
			}
		}
	}
	ls->pos = pos;
	o->lmeta[li] = LM_PACK(LM_TYPE(meta), flags, blk_done);
}

/* ---- sauPhasor_fill (wosc.h:135-169): scan over rounded increments ------ */

/* low 32 bits of sau_ftoi(x) (generator.c:16-17): F2I.S64 saturates where
 * x86 returns INT64_MIN; only the positive overflow differs in the low word */
__device__ __forceinline__ uint32_t ftoi_lo32(float x) {
	/* every float >= 2^55 is a multiple of 2^32 (low word 0), so capping at 2^62 changes no
	 * low word below the overflow and gives 0 above it (and for NaN, which min() drops): one
	 * FMNMX instead of a compare and a select (checked over all floats by saugen_selftest) */
	return (uint32_t) __float2ll_rn(fminf(x, 4611686018427387904.f));
}

template <bool FULLC, typename C>
__device__ __forceinline__ void phasor_eval(const C &c, OpState *o, uint32_t phase0,
		const float f[SPL], const float *pm, const float *fpm, uint32_t n, uint32_t ph[SPL]) {
	const float coeff = c.coeff;
	const uint32_t i0 = c.lane * SPL;
	uint32_t p[SPL], ofs[SPL];
	uint32_t run = 0;
#pragma unroll
	for (int k = 0; k < SPL; ++k) {
		uint32_t inc = ftoi_lo32(coeff * f[k]);
		if (!FULLC && !(i0 + k < n)) inc = 0u;
		run += inc;
		p[k] = run;
	}
	if (pm && fpm) {
#pragma unroll
		for (int k = 0; k < SPL; ++k)
			ofs[k] = ftoi_lo32((((fpm[k] * f[k]) * SAU_FPM_SCALE) + pm[k]) * 2147483648.f);
	} else if (pm) {
#pragma unroll
		for (int k = 0; k < SPL; ++k) ofs[k] = ftoi_lo32(pm[k] * 2147483648.f);
	} else if (fpm) {
#pragma unroll
		for (int k = 0; k < SPL; ++k)
			ofs[k] = ftoi_lo32((fpm[k] * f[k]) * (SAU_FPM_SCALE * 2147483648.f));
	} else {
#pragma unroll
		for (int k = 0; k < SPL; ++k) ofs[k] = 0u;
	}
	const uint32_t incl = scan_incl_u32(run, c.lane);
	const uint32_t base = phase0 + (incl - run);
#pragma unroll
	for (int k = 0; k < SPL; ++k) ph[k] = base + p[k] + ofs[k];
	if (c.lane == 31) o->i0 = phase0 + incl;
}

/* ---- sauWOsc_run / sauWOsc_run_selfmod (wosc.h:215-310) ----------------- */

/* Differentiation (re)start, wosc.h:215-230; ph0 = phase of the chunk's sample 0. */
__device__ __forceinline__ void wosc_reset(const ColdCtx &c, const WaveRef &lut, uint32_t wave,
		uint32_t ph0, uint32_t &prev_phase, double &prev_Is, float &prev_s) {
	double poly, c0;
	herp_ref(lut, ph0 - sau::WAVE_SLEN, &poly, &c0);
	const double Is = herp_ref(lut, ph0, (double*) 0, (double*) 0);
	prev_s = (float) (((Is - poly) - c0) * (double) c.wc->amp256[wave] +
			(double) c.wc->diff_offset[wave]);
	prev_Is = Is;
	prev_phase = ph0;
}

/* diff_scale / (float) phase_diff, IEEE round-to-nearest: the instruction
 * sequence of div.rn.f32 without its operand-range check -- the divisor is a
 * non-zero int32 and the dividend amp_scale * 2^29, far from any exponent
 * limit (checked over the whole divisor range by saugen_selftest). */
__device__ __forceinline__ float div_scale_by_int(float a, int32_t d) {
	const float b = (float) d;
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
	const float e = __fmaf_rn(-b, r, 1.f);
	r = __fmaf_rn(r, e, r);
	float q = __fmaf_rn(a, r, 0.f);
	float rem = __fmaf_rn(-b, q, a);
	q = __fmaf_rn(r, rem, q);
	rem = __fmaf_rn(-b, q, a);
	return __fmaf_rn(r, rem, q);
}

/* Parallel form: sample i needs phase[i], phase[i-1] only.
 * Any n; handles zero phase differences and the differentiator restart. */
__device__ __noinline__ float4 wosc_eval_any(const ColdCtx c, OpState *o, const uint4 ph4,
		uint32_t n) {
	const uint32_t ph[SPL] = {ph4.x, ph4.y, ph4.z, ph4.w};
	float s[SPL];
	const uint32_t wave = o->mode;
	const WaveRef lut = wave_ref(c, wave);
	const float ds = c.wc->diff_scale[wave], doff = c.wc->diff_offset[wave];
	uint32_t prev_phase = o->i1;
	double prev_Is = o->prev_Is;
	float prev_s = o->prev_s;
	uint32_t oscflags = o->oscflags;
	if (oscflags & OSC_RESET_DIFF) {
		const uint32_t ph0 = __shfl_sync(FULL, ph[0], 0);
		wosc_reset(c, lut, wave, ph0, prev_phase, prev_Is, prev_s);
		oscflags &= ~OSC_RESET_DIFF;
	}
	const uint32_t i0 = c.lane * SPL;
	double Is[SPL];
#pragma unroll
	for (int k = 0; k < SPL; ++k) Is[k] = herp_ref(lut, ph[k], (double*) 0, (double*) 0);
	/* sample before this lane's first: previous lane's last, or carried state */
	uint32_t pph = __shfl_up_sync(FULL, ph[SPL - 1], 1);
	double pIs = __shfl_up_sync(FULL, Is[SPL - 1], 1);
	if (c.lane == 0) { pph = prev_phase; pIs = prev_Is; }
	bool zd[SPL];               // valid sample with zero phase difference
	bool lead_zero = false;     // has zero-difference samples before its first computed one
	bool has_nz = false;
	float s_run = 0.f;
#pragma unroll
	for (int k = 0; k < SPL; ++k) {
		const bool valid = (i0 + k) < n;
		const int32_t d = (int32_t) (ph[k] - pph);
		zd[k] = valid && d == 0;
		if (valid && d != 0) {
			s_run = sau::wosc_diff(Is[k], pIs, d, ds, doff);
			has_nz = true;
		}
		if (zd[k] && !has_nz) lead_zero = true;
		s[k] = s_run;
		pph = ph[k]; pIs = Is[k];
	}
	/* zero-difference samples repeat the last computed output (wosc.h:251-252):
	 * fetch it from the nearest lower lane that computed one, else carried state */
	const uint32_t any_lead = __ballot_sync(FULL, lead_zero);
	if (any_lead) {
		const uint32_t nzmask = __ballot_sync(FULL, has_nz);
		const uint32_t lower = nzmask & ((1u << c.lane) - 1u);
		const int src = lower ? (31 - __clz(lower)) : 0;
		float inc = __shfl_sync(FULL, s_run, src);
		if (!lower) inc = prev_s;
		bool seen = false;
#pragma unroll
		for (int k = 0; k < SPL; ++k) {
			if (!zd[k] && (i0 + k) < n) seen = true;
			if (!seen) s[k] = inc;
		}
		if (!has_nz) s_run = inc;
	}
	/* carried state = last valid sample (n >= 1 here) */
	const uint32_t li = n - 1;
	const int src_lane = (int) (li / SPL), src_k = (int) (li % SPL);
	uint32_t e_ph = ph[0]; double e_Is = Is[0]; float e_s = s[0];
#pragma unroll
	for (int k = 1; k < SPL; ++k) if (src_k == k) { e_ph = ph[k]; e_Is = Is[k]; e_s = s[k]; }
	e_ph = __shfl_sync(FULL, e_ph, src_lane);
	e_Is = __shfl_sync(FULL, e_Is, src_lane);
	e_s = __shfl_sync(FULL, e_s, src_lane);
	__syncwarp();
	if (c.lane == 0) {
		o->i1 = e_ph; o->prev_Is = e_Is; o->prev_s = e_s;
		o->oscflags = (uint8_t) oscflags;
	}
	return make_float4(s[0], s[1], s[2], s[3]);
}

/* FULL chunk, no restart pending, carried state in registers.  Returns false
 * (nothing written) when some phase difference is zero: the caller then runs
 * wosc_eval_any on the same phases. */
template <typename C>
__device__ __forceinline__ bool wosc_eval_full(const C &c, OpState *o, uint32_t wave,
		uint32_t prev_phase, double prev_Is, const uint32_t ph[SPL], float s[SPL]) {
	const WaveRef lut = wave_ref(c, wave);
	const float ds = c.wc->diff_scale[wave];
	const double doff = (double) c.wc->diff_offset[wave];
	double Is[SPL];
#pragma unroll
	for (int k = 0; k < SPL; ++k) Is[k] = herp_ref(lut, ph[k], (double*) 0, (double*) 0);
	uint32_t pph = __shfl_up_sync(FULL, ph[SPL - 1], 1);
	double pIs = __shfl_up_sync(FULL, Is[SPL - 1], 1);
	if (c.lane == 0) { pph = prev_phase; pIs = prev_Is; }
	int32_t d[SPL];
	d[0] = (int32_t) (ph[0] - pph);
#pragma unroll
	for (int k = 1; k < SPL; ++k) d[k] = (int32_t) (ph[k] - ph[k - 1]);
	bool z = false;
#pragma unroll
	for (int k = 0; k < SPL; ++k) z |= (d[k] == 0);
	if (__any_sync(FULL, z)) return false;
#pragma unroll
	for (int k = 0; k < SPL; ++k) {                               /* wosc.h:254-256 */
		const float xq = div_scale_by_int(ds, d[k]);
		const double dI = Is[k] - (k ? Is[k - 1] : pIs);
		s[k] = (float) (dI * (double) xq + doff);
	}
	if (c.lane == 31) {
		o->i1 = ph[SPL - 1]; o->prev_Is = Is[SPL - 1]; o->prev_s = s[SPL - 1];
	}
	return true;
}

/* Self-PM: a non-linear recurrence through fb_s, truly serial (wosc.h:273-310).
 * One lane runs it with the state in registers; what matters is the length of
 * the dependent chain per sample (fb_s -> phase -> table -> differentiate ->
 * fb_s), so phases and pm_a amounts come in four at a time with one 128-bit
 * shared load each, outputs leave the same way, the table step is two 128-bit
 * loads of the coefficient planes (or four taps) and the float division is the
 * expanded div_scale_by_int.  The output may replace the pm_a buffer in place
 * (dst == pma): each group of four is read before it is written. */
/* lut_s: shared-window address of the wave's coefficient planes (CT) or of tap
 * lut[-1] of its staged float table.  No branch on a zero phase difference:
 * the step is computed regardless and discarded by selects (wosc.h:251-252). */
template <bool CT>
__device__ __forceinline__ float selfmod_step(uint32_t lut_s, uint32_t phase_in, float pm_a,
		float ds, double doff, uint32_t &prev_phase, double &prev_Is, float &prev_s, float &fb_s) {
	const uint32_t phase = phase_in + (uint32_t) sau::ftoi64(fb_s * pm_a * 2147483648.f);
	const int32_t d = (int32_t) (phase - prev_phase);
	double Is;
	if (CT) {
		const uint32_t ind = phase >> sau::WAVE_SLENBITS;
		const double2 hi = lds128d(lut_s + (ind << 4));
		const float2 lo = lds64f(lut_s + CTAB_PLANE_BYTES + (ind << 3));
		Is = sau::herp_horner(hi.x, hi.y, (double) lo.x, phase) + (double) lo.y;
	} else {
		const uint32_t a = lut_s + ((phase >> sau::WAVE_SLENBITS) << 2);
		const float s0 = lds32f(a), s1 = lds32f(a + 4), s2 = lds32f(a + 8), s3 = lds32f(a + 12);
		Is = sau::herp_poly(s0, s1, s2, s3, phase) + (double) s1;
	}
	const float xq = div_scale_by_int(ds, d);                      /* wosc.h:254-256 */
	const float s_new = (float) ((Is - prev_Is) * (double) xq + doff);
	const bool moved = d != 0;
	const float s = moved ? s_new : prev_s;
	prev_Is = moved ? Is : prev_Is;
	prev_phase = phase;                                            /* d == 0: the same value */
	prev_s = s;
	fb_s = (fb_s + s) * 0.5f;
	return s;
}
template <bool CT>
__device__ __forceinline__ void selfmod_loop(uint32_t lut_s, const uint32_t *phase_buf, const float *pma,
		float *dst, uint32_t n, float ds, double doff, uint32_t &prev_phase, double &prev_Is,
		float &prev_s, float &fb_s) {
	uint32_t i = 0;
	for (; i + 4 <= n; i += 4) {
		const uint4 ph = *reinterpret_cast<const uint4*>(phase_buf + i);
		const float4 pa = *reinterpret_cast<const float4*>(pma + i);
		float4 out;
		out.x = selfmod_step<CT>(lut_s, ph.x, pa.x, ds, doff, prev_phase, prev_Is, prev_s, fb_s);
		out.y = selfmod_step<CT>(lut_s, ph.y, pa.y, ds, doff, prev_phase, prev_Is, prev_s, fb_s);
		out.z = selfmod_step<CT>(lut_s, ph.z, pa.z, ds, doff, prev_phase, prev_Is, prev_s, fb_s);
		out.w = selfmod_step<CT>(lut_s, ph.w, pa.w, ds, doff, prev_phase, prev_Is, prev_s, fb_s);
		*reinterpret_cast<float4*>(dst + i) = out;
	}
	for (; i < n; ++i)
		dst[i] = selfmod_step<CT>(lut_s, phase_buf[i], pma[i], ds, doff, prev_phase, prev_Is, prev_s, fb_s);
}
__device__ __noinline__ void wosc_selfmod(const ColdCtx c, OpState *o, const uint32_t *phase_buf,
		const float *pma, float *dst, uint32_t n) {
	__syncwarp();
	if (c.lane == 0) {
		const uint32_t wave = o->mode;
		const WaveRef lut = wave_ref(c, wave);
		const float ds = c.wc->diff_scale[wave];
		const double doff = (double) c.wc->diff_offset[wave];
		uint32_t prev_phase = o->i1;
		double prev_Is = o->prev_Is;
		float prev_s = o->prev_s, fb_s = o->fb_s;
		uint32_t oscflags = o->oscflags;
		if (oscflags & OSC_RESET_DIFF) {
			wosc_reset(c, lut, wave, phase_buf[0], prev_phase, prev_Is, prev_s);
			oscflags &= ~OSC_RESET_DIFF;
		}
		if (lut.ct)
			selfmod_loop<true>(smem_u32(lut.p), phase_buf, pma, dst, n, ds, doff,
					prev_phase, prev_Is, prev_s, fb_s);
		else
			selfmod_loop<false>(smem_u32(lut.p) - 4u, phase_buf, pma, dst, n, ds, doff,
					prev_phase, prev_Is, prev_s, fb_s);
		o->fb_s = fb_s;
		o->i1 = prev_phase; o->prev_Is = prev_Is; o->prev_s = prev_s;
		o->oscflags = (uint8_t) oscflags;
	}
	__syncwarp();
}

/* pm_a decision, generator.c:485-490: made once per reference 1024-block */
__device__ __forceinline__ bool pma_decide(const Ctx &c, OpState *o) {
	const LineState *ls = &o->line[LINE_PMA];
	uint32_t of = o->flags;
	bool run;
	if (c.oc == 0) {
		run = (ls->v0 != 0.f) || (LM_FLAGS(o->lmeta[LINE_PMA]) & SAUABI_LINEP_GOAL);
		of = run ? (of | ON_PMA_RUN) : (of & ~ON_PMA_RUN);
	} else {
		run = (of & ON_PMA_RUN) != 0;
	}
	__syncwarp();
	if (c.lane == 0) o->flags = (uint8_t) of;
	return run;
}

/* block_mix_add / block_mix_mul_waveenv, generator.c:384-440, on registers */
template <bool FULLC>
__device__ __forceinline__ void mix_eval(const Ctx &c, uint32_t out_buf, const float x[SPL],
		const float a[SPL], uint32_t n, uint32_t layer, bool waveenv) {
	const uint32_t i0 = c.lane * SPL;
	float o[SPL];
	if (!FULLC || layer) ld4(c, out_buf, o);
	if (waveenv) {
#pragma unroll
		for (int k = 0; k < SPL; ++k) {
			if (!FULLC && i0 + k >= n) continue;
			const float s_amp = a[k] * 0.5f;
			const float s = (x[k] * s_amp) + fabsf(s_amp);
			o[k] = layer ? o[k] * s : s;
		}
	} else {
#pragma unroll
		for (int k = 0; k < SPL; ++k) {
			if (!FULLC && i0 + k >= n) continue;
			const float v = x[k] * a[k];
			o[k] = layer ? o[k] + v : v;
		}
	}
	st4(c, out_buf, o);
}

/* end of run_block, generator.c:716-728: zero the unfilled tail, count time */
__device__ __forceinline__ void leave_eval(const Ctx &c, OpState *o, uint32_t out_buf,
		uint32_t len, uint32_t plen, uint32_t layer) {
	const uint32_t i0 = c.lane * SPL;
	if (!(o->flags & ON_TIME_INF)) {
		if (!layer && len < plen) {
			float4 v = *B4(c, out_buf);
			if (i0 + 0 >= len) v.x = 0.f;
			if (i0 + 1 >= len) v.y = 0.f;
			if (i0 + 2 >= len) v.z = 0.f;
			if (i0 + 3 >= len) v.w = 0.f;
			*B4(c, out_buf) = v;
		}
		__syncwarp();
		if (c.lane == 0) o->time -= len;
	}
}

/* ---- fused wave operator (run_block_wosc, generator.c:548-602) ---------- *
 * HEAD = run_block entry + frequency line (no FM lists); children (PM / fPM
 * modulators) run between HEAD and TAIL; TAIL = phase fill + amplitude line
 * (no AM lists, no self-PM modulators) + oscillator + block_mix + run_block
 * exit.  A leaf operator does both in one pass with everything in registers.
 * A full chunk in a steady state (the common case) runs from registers with
 * one warp sync after the state loads and one at the end; everything else
 * goes through the *_any forms. */
template <bool HEAD, bool TAIL>
__device__ __forceinline__ void wop(Ctx &c, const Instr &in, uint32_t &pc) {
	OpState *o = op_ptr(c, in.op);
	/* every piece of operator state this instruction needs, then ONE warp sync:
	 * all lanes hold their copy before lane 0 / lane 31 start writing back */
	const uint4 og = *reinterpret_cast<const uint4*>(&o->time);   /* time, type|flags|mode|oscflags, i0, i1 */
	const uint32_t otime = og.x, oflags = (og.y >> 8) & 0xffu, wave = (og.y >> 16) & 0xffu;
	const uint32_t oscflags = og.y >> 24;
	LineRegs rf, ra;
	float4 pg;
	if (HEAD) rf = line_load(o, LINE_FREQ);
	if (TAIL) {
		ra = line_load(o, LINE_AMP);
		pg = *reinterpret_cast<const float4*>(&o->prev_Is);         /* prev_Is, prev_s, fb_s */
	}
	__syncwarp();
	uint32_t len, rem, layer, plen;
	float fr[SPL];
	if (HEAD) {                                                    /* generator.c:675-698 */
		plen = c.stk_len[c.sp];
		rem = c.stk_rem[c.sp];
		if (!(oflags & ON_TIME_INF) && otime < rem) rem = otime;
		len = rem < plen ? rem : plen;
		layer = (in.flags & F_LAYER) ? 1u :
			((in.flags & F_LAYER_PMA) ? (c.pma_flag ? 1u : 0u) : 0u);
		if (!TAIL) {
			++c.sp;
			c.stk_len[c.sp] = len; c.stk_rem[c.sp] = rem; c.stk_layer[c.sp] = layer;
			if (len == 0) { pc = in.aux; return; }
		}
		if (len > 0) {
			const bool has_mul = in.e != NO_BUF;
			float m[SPL];
			if (has_mul) ld4(c, in.e, m);
			if (!(len == (uint32_t) CHUNK &&
					line_eval_full(c.oc, c.lane, o, LINE_FREQ, rf, has_mul ? m : nullptr, fr))) {
				const float4 t = line_eval_any(c.oc, c.lane, o, LINE_FREQ,
						has_mul ? c.bufs + in.e * CHUNK : nullptr, len, rem);
				fr[0] = t.x; fr[1] = t.y; fr[2] = t.z; fr[3] = t.w;
			}
			if (in.flags & F_SKIP_FREQ2) line_skip(c.oc, c.lane, o, LINE_FREQ2, len);
			if (!TAIL || (in.flags & F_KEEP_FREQ)) st4(c, in.b, fr);
		}
		if (!TAIL) { __syncwarp(); return; }
	} else {
		len = c.stk_len[c.sp]; rem = c.stk_rem[c.sp]; layer = c.stk_layer[c.sp];
		plen = c.stk_len[c.sp - 1];
		if (len > 0) ld4(c, in.b, fr);
	}
	if (len > 0) {
		const bool full = len == (uint32_t) CHUNK;
		float pm[SPL], fpm[SPL];
		if (in.c != NO_BUF) ld4(c, in.c, pm);
		if (in.d != NO_BUF) ld4(c, in.d, fpm);
		const double prev_Is = __hiloint2double(__float_as_int(pg.y), __float_as_int(pg.x));
		uint32_t ph[SPL];
		if (full)
			phasor_eval<true>(c, o, og.z, fr, in.c != NO_BUF ? pm : nullptr,
					in.d != NO_BUF ? fpm : nullptr, len, ph);
		else
			phasor_eval<false>(c, o, og.z, fr, in.c != NO_BUF ? pm : nullptr,
					in.d != NO_BUF ? fpm : nullptr, len, ph);
		float am[SPL];
		if (!(full && line_eval_full(c.oc, c.lane, o, LINE_AMP, ra, nullptr, am))) {
			const float4 t = line_eval_any(c.oc, c.lane, o, LINE_AMP, nullptr, len, rem);
			am[0] = t.x; am[1] = t.y; am[2] = t.z; am[3] = t.w;
		}
		if (in.flags & F_SKIP_AMP2) line_skip(c.oc, c.lane, o, LINE_AMP2, len);
		bool selfmod = false;
		if (in.flags & F_MAY_SELFMOD) { __syncwarp(); selfmod = pma_decide(c, o); }
		float s[SPL];
		if (!selfmod) {
			if (in.flags & F_MAY_SELFMOD) line_skip(c.oc, c.lane, o, LINE_PMA, len);
			bool done = false;
			if (full && !(oscflags & OSC_RESET_DIFF))
				done = wosc_eval_full(c, o, wave, og.w, prev_Is, ph, s);
			if (!done) {
				const float4 t = wosc_eval_any(cold(c), o, make_uint4(ph[0], ph[1], ph[2], ph[3]), len);
				s[0] = t.x; s[1] = t.y; s[2] = t.z; s[3] = t.w;
			}
		} else {
			/* scratch: phases over the (consumed) freq buffer, pm_a amounts and
			 * then the output over the buffer after it */
			const float4 pa = line_eval_any(c.oc, c.lane, o, LINE_PMA, nullptr, len, rem);
			*U4(c, in.b) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
			*B4(c, in.b + 1u) = pa;
			wosc_selfmod(cold(c), o, reinterpret_cast<const uint32_t*>(c.bufs + in.b * CHUNK),
					c.bufs + (in.b + 1u) * CHUNK, c.bufs + (in.b + 1u) * CHUNK, len);
			ld4(c, in.b + 1u, s);
			if (in.flags & F_KEEP_FREQ) { __syncwarp(); st4(c, in.b, fr); }   /* scratch over: freq back */
		}
		c.pma_flag = selfmod;
		if (full) mix_eval<true>(c, in.a, s, am, len, layer, (in.flags & F_WAVEENV) != 0);
		else mix_eval<false>(c, in.a, s, am, len, layer, (in.flags & F_WAVEENV) != 0);
	}
	if (!(oflags & ON_TIME_INF)) {                                 /* generator.c:716-728 */
		if (!layer && len < plen) {
			const uint32_t i0 = c.lane * SPL;
			float4 v = *B4(c, in.a);
			if (i0 + 0 >= len) v.x = 0.f;
			if (i0 + 1 >= len) v.y = 0.f;
			if (i0 + 2 >= len) v.z = 0.f;
			if (i0 + 3 >= len) v.w = 0.f;
			*B4(c, in.a) = v;
		}
		if (c.lane == 0) o->time = otime - len;
	}
	c.last_len = len; c.last_rem = rem;
	if (!HEAD) --c.sp;
	__syncwarp();
}

/* ---- sauCyclor_fill (rasg.h:165-222) ------------------------------------ */

__device__ void cyclor_fill(Ctx &c, const Instr &in, uint32_t n) {
	OpState *o = op_ptr(c, in.op);
	float coeff = c.coeff, ps = 2147483648.f;
	if (o->oscflags & 1) { coeff *= 2; ps *= 2; }
	const uint64_t cp0 = ((uint64_t) o->i1 << 32) | o->i0;
	const float4 f4 = *B4(c, in.c);
	const float f[SPL] = {f4.x, f4.y, f4.z, f4.w};
	float pm[SPL] = {0, 0, 0, 0}, fpm[SPL] = {0, 0, 0, 0};
	const bool has_pm = in.d != NO_BUF, has_fpm = in.e != NO_BUF;
	if (has_pm) ld4(c, in.d, pm);
	if (has_fpm) ld4(c, in.e, fpm);
	const uint32_t i0 = c.lane * SPL;
	uint64_t pre[SPL], ofs[SPL];
	uint64_t run = 0;
#pragma unroll
	for (int k = 0; k < SPL; ++k) {
		pre[k] = run;                                              /* post-increment */
		uint64_t inc = (i0 + k < n) ? (uint64_t) sau::ftoi64(coeff * f[k]) : 0ull;
		run += inc;
		int64_t of = 0;
		if (has_pm && has_fpm) of = sau::pofs_pm_fpm(pm[k], fpm[k], f[k], ps);
		else if (has_pm) of = sau::pofs_pm(pm[k], ps);
		else if (has_fpm) of = sau::pofs_fpm(fpm[k], f[k], ps);
		ofs[k] = (uint64_t) of;
	}
	const uint64_t incl = scan_incl_u64(run, c.lane);
	const uint64_t base = cp0 + (incl - run);
	uint32_t cyc[SPL]; float phf[SPL];
#pragma unroll
	for (int k = 0; k < SPL; ++k) {
		const uint64_t cp = base + pre[k] + ofs[k];
		cyc[k] = (uint32_t) (cp >> 32);
		const uint32_t phase = ((uint32_t) cp) >> 1;
		phf[k] = sau::i2f((int32_t) phase) * (1.f / 2147483648.f);
	}
	*U4(c, in.a) = make_uint4(cyc[0], cyc[1], cyc[2], cyc[3]);
	st4(c, in.b, phf);
	const uint64_t total = __shfl_sync(FULL, incl, 31);
	__syncwarp();                      /* every lane holds cp0 before lane 0 rewrites it */
	if (c.lane == 0) {
		const uint64_t cp = cp0 + total;
		o->i0 = (uint32_t) cp; o->i1 = (uint32_t) (cp >> 32);
	}
}

/* ---- sauRasG_run / sauRasG_run_selfmod (rasg.h:692-772) ----------------- */

/* one sample of sauRasG_run_selfmod's loop, rasg.h:248-280 */
__device__ __forceinline__ float rasg_self_step(unsigned func, unsigned flags, int sr, uint32_t alpha,
		int line, float phase_in, uint32_t cycle_in, float pma, float &fb_s, float &prev_s) {
	const float pm_a = fb_s * pma * 0.5f;
	float phase = phase_in + pm_a;
	const int32_t cycle_adj = (int32_t) floorf(phase);
	const uint32_t cycle = cycle_in + (uint32_t) cycle_adj;
	phase -= (float) cycle_adj;
	const float s = sau::rasg_sample(func, flags, sr, alpha, line, cycle, phase, true, false);
	fb_s = ((fb_s + prev_s) + s) * 0.5f;
	prev_s = s;
	return s;
}
/* FUNC folded in (0xff: taken from func_dyn), no option flags; inputs and outputs
 * four at a time (the output replaces the phase buffer in place) */
template <unsigned FUNC>
__device__ __noinline__ void rasg_self_loop(float *main_buf, const uint32_t *cycle_buf, const float *pma,
		uint32_t n, int sr, uint32_t alpha, int line, float &fb_s_io, float &prev_s_io,
		unsigned func_dyn = 0) {
	const unsigned func = FUNC == 0xffu ? func_dyn : FUNC;
	float fb_s = fb_s_io, prev_s = prev_s_io;
	uint32_t i = 0;
	for (; i + 4 <= n; i += 4) {
		const float4 ph = *reinterpret_cast<const float4*>(main_buf + i);
		const uint4 cy = *reinterpret_cast<const uint4*>(cycle_buf + i);
		const float4 pa = *reinterpret_cast<const float4*>(pma + i);
		float4 out;
		out.x = rasg_self_step(func, 0u, sr, alpha, line, ph.x, cy.x, pa.x, fb_s, prev_s);
		out.y = rasg_self_step(func, 0u, sr, alpha, line, ph.y, cy.y, pa.y, fb_s, prev_s);
		out.z = rasg_self_step(func, 0u, sr, alpha, line, ph.z, cy.z, pa.z, fb_s, prev_s);
		out.w = rasg_self_step(func, 0u, sr, alpha, line, ph.w, cy.w, pa.w, fb_s, prev_s);
		*reinterpret_cast<float4*>(main_buf + i) = out;
	}
	for (; i < n; ++i)
		main_buf[i] = rasg_self_step(func, 0u, sr, alpha, line, main_buf[i], cycle_buf[i], pma[i], fb_s, prev_s);
	fb_s_io = fb_s; prev_s_io = prev_s;
}

__device__ void rasg_run(Ctx &c, const Instr &in, uint32_t n, uint32_t blk_len) {
	OpState *o = op_ptr(c, in.op);
	const unsigned flags = o->ras_flags, func = o->ras_func;
	const int sr = o->ras_level, line = o->mode;
	const uint32_t alpha = o->ras_alpha;
	const bool selfmod = (in.flags & F_HAS_APMODS) || c.pma_flag;
	if (selfmod) {
		__syncwarp();
		if (c.lane == 0) {                                         /* rasg.h:242-280 */
			float fb_s = o->fb_s, prev_s = o->prev_s;
			float *main_buf = c.bufs + in.a * CHUNK;
			const uint32_t *cycle_buf = reinterpret_cast<const uint32_t*>(c.bufs + in.b * CHUNK);
			const float *pma = c.bufs + in.c * CHUNK;
			/* the plain modes (no option flags) get a loop with the function folded in:
			 * the serial chain per sample is what this path costs */
			if ((flags & 0x3ffu & ~(SAUABI_RAS_O_LINE_SET | SAUABI_RAS_O_FUNC_SET | SAUABI_RAS_O_LEVEL_SET |
					SAUABI_RAS_O_ASUBVAL_SET)) == 0) {
				switch (func) {
				case SAUABI_RAS_F_URAND: rasg_self_loop<SAUABI_RAS_F_URAND>(main_buf, cycle_buf, pma, n, sr, alpha, line, fb_s, prev_s); break;
				case SAUABI_RAS_F_GAUSS: rasg_self_loop<SAUABI_RAS_F_GAUSS>(main_buf, cycle_buf, pma, n, sr, alpha, line, fb_s, prev_s); break;
				case SAUABI_RAS_F_BIN: rasg_self_loop<SAUABI_RAS_F_BIN>(main_buf, cycle_buf, pma, n, sr, alpha, line, fb_s, prev_s); break;
				case SAUABI_RAS_F_TERN: rasg_self_loop<SAUABI_RAS_F_TERN>(main_buf, cycle_buf, pma, n, sr, alpha, line, fb_s, prev_s); break;
				case SAUABI_RAS_F_FIXED: rasg_self_loop<SAUABI_RAS_F_FIXED>(main_buf, cycle_buf, pma, n, sr, alpha, line, fb_s, prev_s); break;
				default: rasg_self_loop<0xffu>(main_buf, cycle_buf, pma, n, sr, alpha, line, fb_s, prev_s, func); break;
				}
			} else {
				for (uint32_t i = 0; i < n; ++i)
					main_buf[i] = rasg_self_step(func, flags, sr, alpha, line, main_buf[i], cycle_buf[i],
							pma[i], fb_s, prev_s);
			}
			o->fb_s = fb_s; o->prev_s = prev_s;
		}
		return;
	}
	const uint4 cy4 = *U4(c, in.b);
	const uint32_t cy[SPL] = {cy4.x, cy4.y, cy4.z, cy4.w};
	float ph[SPL];
	ld4(c, in.a, ph);
	/* sauLine_map_cub: 4-wide body + scalar tail, counted in the 1024-block */
	const uint32_t tail_from = blk_len & ~3u;
	float out[SPL];
#pragma unroll
	for (int k = 0; k < SPL; ++k) {
		const uint32_t idx = c.lane * SPL + k;
		out[k] = sau::rasg_sample(func, flags, sr, alpha, line, cy[k], ph[k], false,
				(c.oc + idx) >= tail_from);
	}
	st4(c, in.a, out);
}

/* ---- sauNoiseG_run_* (noise.h:41-185) ----------------------------------- */

__device__ __forceinline__ int32_t noise_tern(uint32_t n) {       /* bv's s1, noise.h:165-167 */
	int32_t s1 = sau::sar32((int32_t) sau::ranfast32(n), 31);
	return (n & 1) ? (s1 * 2 + 1) : 0;
}
__device__ void noise_run(Ctx &c, const Instr &in, uint32_t n) {
	OpState *o = op_ptr(c, in.op);
	const uint32_t n0 = o->i0, prev = o->i1, type = o->mode;
	const float scale = 1.f / 2147483648.f;
	const uint32_t i0 = c.lane * SPL;
	float out[SPL];
	uint32_t new_prev = prev;
	switch (type) {
	default:
	case SAUABI_NOISE_wh:
#pragma unroll
		for (int k = 0; k < SPL; ++k) out[k] = sau::fscalei(sau::ranfast32(n0 + i0 + k), scale);
		break;
	case SAUABI_NOISE_gw:
#pragma unroll
		for (int k = 0; k < SPL; ++k) out[k] = sau::franssgauss32(n0 + i0 + k);
		break;
	case SAUABI_NOISE_bw:
#pragma unroll
		for (int k = 0; k < SPL; ++k)
			out[k] = (float) (sau::sar32((int32_t) sau::ranfast32(n0 + i0 + k), 31) * 2 + 1);
		break;
	case SAUABI_NOISE_tw:
#pragma unroll
		for (int k = 0; k < SPL; ++k) {
			const uint32_t nn = n0 + i0 + k;
			const int32_t s = sau::sar32((int32_t) sau::ranfast32(nn), 31) * 2 + 1;
			out[k] = (nn & 1) ? (float) s : 0.f;
		}
		break;
	case SAUABI_NOISE_re: {                                        /* integer prefix sum */
		uint32_t p[SPL], run = 0;
#pragma unroll
		for (int k = 0; k < SPL; ++k) {
			const int32_t s = (int32_t) sau::ranfast32(n0 + i0 + k);
			run += (i0 + k < n) ? (uint32_t) (s >> 6) : 0u;
			p[k] = run;
		}
		const uint32_t incl = scan_incl_u32(run, c.lane);
		const uint32_t base = prev + (incl - run);
#pragma unroll
		for (int k = 0; k < SPL; ++k)
			out[k] = sau::fscalei((uint32_t) sau::foldhd32((int32_t) (base + p[k])), scale);
		new_prev = prev + __shfl_sync(FULL, incl, 31);
		break; }
	case SAUABI_NOISE_vi:                                          /* 1-sample shift */
#pragma unroll
		for (int k = 0; k < SPL; ++k) {
			const uint32_t idx = i0 + k;
			const uint32_t s1 = sau::ranfast32(n0 + idx);
			const uint32_t s0 = idx ? sau::ranfast32(n0 + idx - 1) : prev;
			out[k] = sau::fscalei((s1 / 2) - (s0 / 2), scale);
		}
		if (n) new_prev = sau::ranfast32(n0 + n - 1);
		break;
	case SAUABI_NOISE_bv:
#pragma unroll
		for (int k = 0; k < SPL; ++k) {
			const uint32_t idx = i0 + k;
			const int32_t s1 = noise_tern(n0 + idx);
			const int32_t s0 = idx ? noise_tern(n0 + idx - 1) : (int32_t) prev;
			out[k] = (float) (s1 - s0);
		}
		if (n) new_prev = (uint32_t) noise_tern(n0 + n - 1);
		break;
	}
	st4(c, in.a, out);
	__syncwarp();
	if (c.lane == 0) { o->i0 = n0 + n; o->i1 = new_prev; }
}

/* ---- event application (generator.c:233-377, line.c:287-332) ------------ */

__device__ void dev_line_copy(OpState *n, int li, const LineDelta *src) {
	if (!src->present) return;
	LineState *o = &n->line[li];
	const uint32_t meta = n->lmeta[li];
	uint32_t mask = 0, flags = LM_FLAGS(meta), type = LM_TYPE(meta);
	const uint32_t sf = src->flags;
	if (sf & SAUABI_LINEP_STATE) {
		o->v0 = src->v0;
		mask |= SAUABI_LINEP_STATE | SAUABI_LINEP_STATE_RATIO;
	} else if (flags & SAUABI_LINEP_GOAL) {
		if (sf & SAUABI_LINEP_GOAL) {
			/* sauLine_get(o, &f, 1, NULL): one value on the old trajectory */
			if (flags & SAUABI_LINEP_GOAL_RATIO) flags |= SAUABI_LINEP_STATE_RATIO;
			else flags &= ~SAUABI_LINEP_STATE_RATIO;
			if (o->pos < o->end) {
				sau::LineFill f = sau::line_fill_setup((int) type, o->v0, o->vt, o->pos, o->end);
				o->v0 = sau::line_fill_at(f, 0, true);   /* 1-element fill = gcc's tail */
			}
		}
	}
	if (sf & SAUABI_LINEP_GOAL) {
		o->vt = src->vt;
		if (sf & SAUABI_LINEP_TIME_IF_NEW) o->end -= o->pos;
		o->pos = 0;
		mask |= SAUABI_LINEP_GOAL | SAUABI_LINEP_GOAL_RATIO;
	}
	if (sf & SAUABI_LINEP_TYPE) {
		type = src->type;
		mask |= SAUABI_LINEP_TYPE;
	}
	if (!(flags & SAUABI_LINEP_TIME) || !(sf & SAUABI_LINEP_TIME_IF_NEW)) {
		if (sf & SAUABI_LINEP_TIME) {
			o->end = src->end_samples;
			mask |= SAUABI_LINEP_TIME;
		}
	}
	flags &= ~mask;
	flags |= (sf & mask);
	n->lmeta[li] = LM_PACK(type, flags, LM_BLK(meta));
	n->linv[li] = 1.f / sau::u2f(o->end);       /* line_fill_setup's reciprocal, kept current */
}

/* R oscillator setters, rasg.h:59-119 */
__device__ __forceinline__ uint64_t ras_cp(const OpState *o) { return ((uint64_t) o->i1 << 32) | o->i0; }
__device__ __forceinline__ void ras_store(OpState *o, uint64_t cp) { o->i0 = (uint32_t) cp; o->i1 = (uint32_t) (cp >> 32); }
__device__ __forceinline__ uint32_t ras_get_cycle(const OpState *o) { return o->i1 & ~1u; }
__device__ __forceinline__ uint32_t ras_get_phase(const OpState *o) {
	return (o->oscflags & 1) ? (uint32_t) (ras_cp(o) >> 1) : o->i0;
}
__device__ void ras_set_cycle(OpState *o, uint32_t cycle) {
	const uint32_t phase = ras_get_phase(o);
	const uint64_t p64 = (o->oscflags & 1) ? ((uint64_t) phase) << 1 : phase;
	ras_store(o, ((uint64_t) (cycle & ~1u)) << 32 | p64);
}
__device__ void ras_set_phase(OpState *o, uint32_t phase) {
	const uint32_t cycle = ras_get_cycle(o);
	const uint64_t p64 = (o->oscflags & 1) ? ((uint64_t) phase) << 1 : phase;
	ras_store(o, ((uint64_t) cycle) << 32 | p64);
}

__device__ __noinline__ void apply_event(const GenDesc *g, const WaveCoeffs *wc, const EventRec *ev,
		VoiceState *vs) {
	for (uint32_t i = 0; i < ev->opdata_count; ++i) {
		const OpDataRec *od = &g->opdata[ev->opdata_off + i];
		/* work on a copy read from / written to L2: under the ticketed scheduler the
		 * operator may last have been stored by another SM */
		OpState *gn = &g->ops[od->id];
		OpState stv;
		{
			uint4 *d = reinterpret_cast<uint4*>(&stv);
			for (uint32_t w = 0; w < sizeof(OpState) / 16; ++w)
				d[w] = __ldcg(reinterpret_cast<const uint4*>(gn) + w);
		}
		OpState *n = &stv;
		if (!(n->flags & ON_INIT)) {                               /* prepare_op, :245-278 */
			OpState z;
			memset(&z, 0, sizeof(z));
			z.type = od->type;
			z.flags = ON_INIT;
			if (od->type == SAUABI_POPT_wave) {                    /* wosc.h:55-71 */
				z.i0 = (uint32_t) wc->phase_adj[SAUABI_WAVE_sin];
				z.mode = SAUABI_WAVE_sin;
				z.oscflags = OSC_RESET_DIFF;
			} else if (od->type == SAUABI_POPT_raseg) {            /* rasg.h:44-57 */
				z.oscflags = 1;   /* rate2x */
				z.mode = SAUABI_LINE_lin;
				z.ras_func = SAUABI_RAS_F_URAND;
				z.ras_level = 27;
				z.ras_alpha = 0x9e3779b9u;
			}
			*n = z;
		}
		const uint32_t params = od->params;                        /* update_op, :283-343 */
		bool osc = false;
		switch (od->type) {
		case SAUABI_POPT_noise:
			if (params & SAUABI_POPP_MODE) { n->mode = od->mode_main; n->i1 = 0; }
			if (params & SAUABI_POPP_SEED) n->i0 = od->seed;
			break;
		case SAUABI_POPT_wave:
			if (params & SAUABI_POPP_MODE) {                       /* wosc.h:81-87 */
				const uint32_t wave = od->mode_main;
				n->i0 += (uint32_t) wc->phase_adj[wave] - (uint32_t) wc->phase_adj[n->mode];
				n->mode = (uint8_t) wave;
				n->oscflags |= OSC_RESET_DIFF;
			}
			if (params & SAUABI_POPP_PHASE)
				n->i0 = od->phase + (uint32_t) wc->phase_adj[n->mode];
			osc = true;
			break;
		case SAUABI_POPT_raseg:
			if (params & SAUABI_POPP_MODE) {                       /* rasg.h:97-119 */
				unsigned flags = od->ras_flags;
				if (flags & SAUABI_RAS_O_LINE_SET) n->mode = od->mode_main;
				if (flags & SAUABI_RAS_O_FUNC_SET) n->ras_func = od->ras_func;
				else flags |= n->ras_flags;
				if (od->ras_flags & SAUABI_RAS_O_LEVEL_SET) n->ras_level = od->ras_level;
				if (od->ras_flags & SAUABI_RAS_O_ASUBVAL_SET) n->ras_alpha = od->ras_alpha;
				n->ras_flags = (uint16_t) (flags & 0x3ff);
				const bool rate2x = !(flags & SAUABI_RAS_O_HALFSHAPE);
				if (rate2x != (bool) (n->oscflags & 1)) {
					const uint32_t cycle = ras_get_cycle(n);
					const uint32_t phase = ras_get_phase(n);
					n->oscflags = rate2x ? 1 : 0;
					ras_set_cycle(n, cycle);
					ras_set_phase(n, phase);
				}
			}
			if (params & SAUABI_POPP_PHASE) ras_set_phase(n, od->phase);
			if (params & SAUABI_POPP_SEED) ras_set_cycle(n, od->seed);
			osc = true;
			break;
		}
		if (osc) {
			dev_line_copy(n, LINE_FREQ, &od->line[LINE_FREQ]);
			dev_line_copy(n, LINE_FREQ2, &od->line[LINE_FREQ2]);
			dev_line_copy(n, LINE_PMA, &od->line[LINE_PMA]);
		}
		if (params & SAUABI_POPP_TIME) {
			if (od->time_flags & SAUABI_TIMEP_IMPLICIT) {
				n->time = 0;
				n->flags |= ON_TIME_INF;
			} else {
				n->time = od->time_samples;
				n->flags &= ~ON_TIME_INF;
			}
		}
		dev_line_copy(n, LINE_AMP, &od->line[LINE_AMP]);
		dev_line_copy(n, LINE_AMP2, &od->line[LINE_AMP2]);
		dev_line_copy(n, LINE_PAN, &od->line[LINE_PAN]);
		{
			const uint4 *s = reinterpret_cast<const uint4*>(&stv);
			for (uint32_t w = 0; w < sizeof(OpState) / 16; ++w)
				__stcg(reinterpret_cast<uint4*>(gn) + w, s[w]);
		}
	}
	vs->carr_op = ev->carr_op_id;
	vs->flags |= VN_INIT;
	vs->code_off = ev->code_off;
	vs->code_len = ev->code_len;
	vs->ops_off = ev->ops_off;
	vs->ops_cnt = ev->ops_cnt;
	vs->carr_slot = ev->carr_slot;
	vs->duration = __ldcg(&g->ops[vs->carr_op].time);              /* set_voice_duration */
	__threadfence();
}

/* ---- bytecode interpreter: one chunk of one voice ----------------------- */

__device__ __noinline__ uint32_t run_chunk(Ctx &c, const Instr *code, uint32_t code_len, uint32_t time,
		uint32_t rem0, float *row_s, float *row_r, uint32_t frame) {
	c.sp = 0;
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
enum : uint32_t { P_LINE = 1, P_WHEAD, P_WTAIL, P_WLEAF, P_PHASE, P_WOSC, P_RANGE, P_VOUT,
	P_NOISE, P_CYCLE, P_RASG, P_MIX, P_WSELF };
enum : uint32_t {
	PF_LAYER = 1, PF_WAVEENV = 2,
	PF_FUNI = 4,       /* frequency (or the LINE's value) is uniform over the block: w6 holds the
	                    * value (LINE, WHEAD) or the phase increment (WTAIL, WLEAF, PHASE) */
	PF_FMUL = 8,       /* WHEAD / WLEAF, not uniform: the frequency is the constant w6 times the
	                    * (varying) multiplier buffer: a ratio to a modulated parent frequency */
	PF_ACONST = 16,    /* amplitude line holds av */
	PF_ABUF = 32,      /* amplitude comes from work buffer c (the operator has amplitude modulators) */
};
constexpr uint32_t PLAN_FBUF = 32 * FAST_NS * 4;    /* FastCfg<FAST_NS>::FBUF_BYTES */
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
	plan += PLAN_REC;          /* slot 0 is the header (render_units) */
	if (cap) --cap;
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
			const uint32_t slot = __popc(wave_mask & ((1u << wave) - 1u));
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
			const uint32_t slot = __popc(wave_mask & ((1u << wave) - 1u));
			const uint32_t ct = (wave_mask & CTAB_FLAG) ? st + slot * CTAB_WAVE_BYTES :
				st + slot * (TAB_STRIDE * 4) + 12;                   /* planes, or &lut[-1] */
			const bool aconst = !(LM_FLAGS(o->lmeta[LINE_AMP]) & SAUABI_LINEP_GOAL);
			const uint32_t fl = ((in.flags & F_LAYER) ? PF_LAYER : 0u) | ((in.flags & F_WAVEENV) ? PF_WAVEENV : 0u) |
				(funi ? PF_FUNI : 0u) | (rmul ? PF_FMUL : 0u) | (aconst ? PF_ACONST : 0u);
			plan_put(plan, n++, (head ? P_WLEAF : P_WTAIL) | fl << 8 | (uint32_t) in.a << 16 | (uint32_t) in.b << 24,
					(uint32_t) in.c | (uint32_t) in.e << 8, opa, ct,
					wc->diff_scale[wave], wc->diff_offset[wave], rmul ? fval : __uint_as_float(inc),
					o->line[LINE_AMP].v0);
			dirty(in.a);
		}
	}
	return finish(n);
}

/* sauLine_run's bookkeeping for nb whole blocks of a steady run line, block by block */
__device__ __forceinline__ void line_block_update(OpState *o, int li, uint32_t nb) {
	const uint32_t meta = o->lmeta[li];
	uint32_t flags = LM_FLAGS(meta), pos = o->line[li].pos;
	if (flags & SAUABI_LINEP_GOAL) {
		pos += nb * (uint32_t) REF_BLOCK;
	} else {
		for (uint32_t b = 0; b < nb; ++b) {
			bool ex;
			line_advance(pos, o->line[li].end, flags, REF_BLOCK, ex);
		}
	}
	o->line[li].pos = pos;
	o->lmeta[li] = LM_PACK(LM_TYPE(meta), flags, 0u);
}
__device__ __forceinline__ void line_skip_blocks(OpState *o, int li, uint32_t nb) {
	for (uint32_t b = 0; b < nb; ++b) line_skip(0, 0, o, li, REF_BLOCK);
}

/* lane 0 only */
__device__ __noinline__ void steady_update(OpState *sops, const Instr *code, uint32_t code_len, uint32_t nb) {
	uint4 raw_next = __ldg(reinterpret_cast<const uint4*>(code));
	for (uint32_t pc = 0; pc < code_len; ++pc) {
		const uint4 raw = raw_next;
		if (pc + 1 < code_len) raw_next = __ldg(reinterpret_cast<const uint4*>(code + pc + 1));
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
/* plan header (the first 32-byte slot): the cold paths' context and VOUT's constants */
constexpr uint32_t PH_TAB = 0, PH_WC = 8, PH_WAVE_MASK = 16, PH_AMP_SCALE = 20, PH_WRITE_R = 24,
	PH_TSTRIDE = 28;
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
constexpr uint32_t OS_LINE = 0, OS_LMETA = 96, OS_LINV = 120, OS_TIME = 144, OS_I0 = 152,
	OS_I1 = 156, OS_PREV = 160;
static_assert(offsetof(OpState, lmeta) == OS_LMETA && offsetof(OpState, linv) == OS_LINV &&
		offsetof(OpState, time) == OS_TIME && offsetof(OpState, i0) == OS_I0 &&
		offsetof(OpState, i1) == OS_I1 && offsetof(OpState, prev_Is) == OS_PREV, "OpState offsets");

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
				sts32(op + OS_PREV + 8, __float_as_uint(s[NS - 1]));
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

template <int NS, bool CTAB, bool OTHER>
__device__ __forceinline__ void run_chunk_plan(const HotCtx &c, const uint32_t nrec,
		float *row_s, float *row_r, const uint32_t frame) {
	uint32_t rec = c.plan;
	for (uint32_t r = 0; r < nrec; ++r) {
		rec += PLAN_REC;
		const uint4 p0 = lds128u(rec);
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
					if (kind != P_WLEAF) { fst<NS>(c, is_line ? bufa : bufb, fr); continue; }
				}
				const uint32_t bufc = p0.y & 0xffu;
				phase_plan<NS>(c, op, bufc, funi, inc, fr, ph);
				if (kind == P_PHASE) {
					float pf[NS];
#pragma unroll
					for (int k = 0; k < NS; ++k) pf[k] = __uint_as_float(ph[k]);
					fst<NS>(c, bufa, pf);
					continue;
				}
				pure = funi && bufc == NO_BUF;
			} else {
				float pf[NS];
				fld<NS>(c, bufb, pf);
#pragma unroll
				for (int k = 0; k < NS; ++k) ph[k] = __float_as_uint(pf[k]);
			}
			osc_plan<NS, CTAB>(c, p0, rec, pure, inc, ph);
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
			const uint32_t write_r = lds32(c.plan + PH_WRITE_R);
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
			return;
		}
	}
}

/* One steady stretch: its own function, so that the hot loop gets its own register
 * allocation whatever the general path around the call needs.  OTHER: the plan has
 * serial self-PM records (plan_other); feed-forward plans run in the other instance. */
template <bool CTAB, bool OTHER>
__device__ __noinline__ void run_block_fast(uint32_t sb, uint32_t plan, int lane, float coeff, uint32_t nrec,
		uint32_t len, float *row_s, float *row_r, uint32_t frame) {
	HotCtx c;
	c.sb = sb; c.plan = plan; c.lane = lane; c.coeff = coeff;
	for (uint32_t oc = 0; oc < len; oc += FastCfg<FAST_NS>::CHUNKF) {
		c.oc = oc;
		run_chunk_plan<FAST_NS, CTAB, OTHER>(c, nrec, row_s, row_r, frame + oc);
	}
}

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
	{
		/* lane 0 reads (from L2: another SM may have written it), every lane gets
		 * the same copy */
		const uint32_t *src = reinterpret_cast<const uint32_t*>(vsp);
		uint32_t *dst = reinterpret_cast<uint32_t*>(&vs);
#pragma unroll
		for (uint32_t i = 0; i < sizeof(VoiceState) / 4; ++i) {
			uint32_t w = 0;
			if (lane == 0) w = __ldcg(src + i);
			dst[i] = __shfl_sync(FULL, w, 0);
		}
	}
	const uint32_t ev_lo = g->vev_off[v], ev_n = g->vev_off[v + 1] - ev_lo;
	float *row_s = g->rows_s + (size_t) lv * ROW_TILE;      /* the voice's piece of frame tile 0 */
	float *row_r = g->rows_r + (size_t) lv * ROW_TILE;
	c.tstride = g->row_stride;
	uint32_t loaded = 0;        // operator states currently held in shared memory

	for (uint32_t ui = u0; ui < u1; ++ui) {
		const UnitDesc ud = units[cd->unit_off + ui];
		const uint32_t si = ud.seg;
		const SegDesc sd = segs[cd->seg_off + si];
		/* this voice's events due at the segment start, in order */
		if (ud.off == 0 && vs.ev_cursor < ev_n && g->vev_idx[ev_lo + vs.ev_cursor] < sd.ev_end) {
			if (loaded) { ops_store(c, loaded); loaded = 0; }
			while (vs.ev_cursor < ev_n && g->vev_idx[ev_lo + vs.ev_cursor] < sd.ev_end) {
				if (lane == 0) {
					apply_event(g, c.wc, &g->events[g->vev_idx[ev_lo + vs.ev_cursor]], &vs);
					vs.ev_cursor++;
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
			if (off % REF_BLOCK == 0 && uend - off >= (uint32_t) REF_BLOCK &&
					vs.duration >= (uint32_t) REF_BLOCK && vs.code_len &&
					op_ptr(c, vs.carr_slot)->time > 0) {
				uint32_t kb = (uend - off) / (uint32_t) REF_BLOCK;
				if (vs.duration / (uint32_t) REF_BLOCK < kb) kb = vs.duration / (uint32_t) REF_BLOCK;
				if (kb > 0x7fffu) kb = 0x7fffu;        /* 15 bits in steady_plan's result */
				sp = steady_plan(c.sops, fc.so, fc.st, fc.wave_mask, fc.wc, g->code + vs.code_off,
						vs.code_len, fc.plan, fc.plan_cap, kb, fc.sb, fc.coeff);
			}
			if (sp) {
				const uint32_t nrec = sp & 0xffffu, nb = (sp >> 16) & 0x7fffu, span = nb * (uint32_t) REF_BLOCK;
				const bool other = (sp >> 31) != 0;
				if (pan_mode == PAN_UNSET)       /* steady => the pan stands still */
					pan_mode = __float_as_uint(op_ptr(c, vs.carr_slot)->line[LINE_PAN].v0);
				{
					/* plan header: what the rare paths and VOUT need */
					const uint64_t tp = reinterpret_cast<uint64_t>(fc.tab), wp = reinterpret_cast<uint64_t>(fc.wc);
					plan_put(fc.plan, 0, (uint32_t) tp, (uint32_t) (tp >> 32), (uint32_t) wp, (uint32_t) (wp >> 32),
							__uint_as_float(fc.wave_mask), fc.amp_scale, __uint_as_float(fc.write_r),
							__uint_as_float(c.tstride));
					__syncwarp();
				}
				if (fc.wave_mask & CTAB_FLAG) {
					if (other) run_block_fast<true, true>(fc.sb, fc.plan, lane, fc.coeff, nrec, span, row_s, row_r, sd.start + off);
					else run_block_fast<true, false>(fc.sb, fc.plan, lane, fc.coeff, nrec, span, row_s, row_r, sd.start + off);
				} else {
					if (other) run_block_fast<false, true>(fc.sb, fc.plan, lane, fc.coeff, nrec, span, row_s, row_r, sd.start + off);
					else run_block_fast<false, false>(fc.sb, fc.plan, lane, fc.coeff, nrec, span, row_s, row_r, sd.start + off);
				}
				__syncwarp();
				if (lane == 0) steady_update(c.sops, g->code + vs.code_off, vs.code_len, nb);
				__syncwarp();
				vs.duration -= span;
				run_total += span;
				off += span - CHUNK;
				continue;
			}
			uint32_t clen = uend - off;
			if (clen > (uint32_t) CHUNK) clen = CHUNK;
			const uint32_t time = vs.duration < clen ? vs.duration : clen;
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
			if (__ldcg(&g->status[1 + si]) < tot) atomicMax(&g->status[1 + si], tot);
		}
	}
	if (loaded) ops_store(c, loaded);
	if (lane == 0) {
		const uint32_t *w = reinterpret_cast<const uint32_t*>(&vs);
#pragma unroll
		for (uint32_t i = 0; i < sizeof(VoiceState) / 4; ++i)
			__stcg(reinterpret_cast<uint32_t*>(vsp) + i, w[i]);
		if (u1 == cd->nunits && vs.duration != 0) atomicOr(&g->status[0], 1u);
	}
}

__device__ __forceinline__ void render_body(const CallDesc *calls, uint32_t ncalls,
		const SegDesc *segs, const UnitDesc *units, uint32_t ntasks, const float *tables,
		const double *coefs, uint32_t wave_mask, uint32_t nbufs, uint32_t nslots_ops, uint32_t nplan, uint32_t warps_per_cta,
		uint32_t ticketed) {
	extern __shared__ __align__(128) unsigned char smem[];
	uint64_t *bar = reinterpret_cast<uint64_t*>(smem);
	float *tab = reinterpret_cast<float*>(smem + 128);
	const bool ctab = (wave_mask & CTAB_FLAG) != 0;
	const uint32_t nslots = __popc(wave_mask & ~CTAB_FLAG);
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

	Ctx c;
	c.sops = reinterpret_cast<OpState*>(warp_area + warp * per_warp);
	c.bufs = reinterpret_cast<float*>(c.sops + nslots_ops);
	c.stk_len = reinterpret_cast<uint32_t*>(c.bufs + nbufs * BUF_FLOATS);
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

	if (!ticketed) {
		/* one warp renders every unit of one voice; task -> (call, voice) by
		 * binary search on task_base */
		const uint32_t task = blockIdx.x * warps_per_cta + warp;
		if (task >= ntasks) return;
		uint32_t ci = 0, hi = ncalls;
		while (hi - ci > 1) {
			const uint32_t mid = (ci + hi) >> 1;
			if (calls[mid].task_base <= task) ci = mid; else hi = mid;
		}
		const CallDesc *cd = &calls[ci];
		render_units(c, fc, cd, segs, units, task - cd->task_base, 0, cd->nunits);
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
		uint32_t nslots_ops, uint32_t nplan, uint32_t warps_per_cta, uint32_t ticketed) {
	render_body(calls, ncalls, segs, units, ntasks, tables, coefs, wave_mask, nbufs, nslots_ops, nplan,
			warps_per_cta, ticketed);
}

/* same body for CTAs of up to 32 warps (64 registers), one per SM (coefficient-table
 * mode: the planes take 48 KiB per wave, so one large CTA shares them among all the
 * warps an SM can hold -- the path is latency-bound, resident warps are what counts) */
__global__ void __launch_bounds__(WIDE_WARPS * 32, 1)
render_kernel_wide(const CallDesc *calls, uint32_t ncalls, const SegDesc *segs, const UnitDesc *units,
		uint32_t ntasks, const float *tables, const double *coefs, uint32_t wave_mask, uint32_t nbufs,
		uint32_t nslots_ops, uint32_t nplan, uint32_t warps_per_cta, uint32_t ticketed) {
	render_body(calls, ncalls, segs, units, ntasks, tables, coefs, wave_mask, nbufs, nslots_ops, nplan,
			warps_per_cta, ticketed);
}

/* ---- mix + clip epilogue ------------------------------------------------- */

/* One CTA mixes one frame tile (ROW_TILE = 128 consecutive frames), one thread
 * per frame: the sum over voices must run in voice order in ONE thread (float
 * addition is not associative and mix_add adds voice after voice,
 * generator.c:773-786).  The tile's voice pieces lie side by side in HBM
 * (device_types.h:ROW_TILE), so the CTA reads ONE contiguous stream: the
 * producer warp moves MIX_TV voices (8 KiB) per stage with a single TMA bulk copy
 * (cp.async.bulk + mbarrier transaction count) into a ring of MIX_STAGES
 * stages, the four consumer warps add behind it.
 * A voice whose pan stands still contributes r = s * pan, computed here
 * (VoiceSeg); only moving pans have an r piece, which the consumers read straight
 * from HBM (a coalesced 128-byte line per warp; rare). */
constexpr int MIX_FRAMES = ROW_TILE;           // = consumer threads (one per frame)
constexpr int MIX_TV = 32;                     // voices per stage (16 KiB per bulk copy)
constexpr int MIX_STAGES = 5;
constexpr int MIX_CWARPS = MIX_FRAMES / 32;    // consumer warps; one more warp produces
struct MixSmem {
	float s[MIX_STAGES][MIX_TV][MIX_FRAMES];
	uint2 vi[MIX_STAGES][MIX_TV];              // the tile's VoiceSeg records
	uint64_t full[MIX_STAGES], empty[MIX_STAGES];
	uint32_t ndyn[MIX_STAGES];                 // moving-pan voices in the stage's tile
};

__device__ __forceinline__ void mix_store(const GenDesc *g, const CallDesc *cd, uint32_t mode,
		uint32_t f, float L, float R) {
	if (mode == 1) {
		g->mix[f] = L;
		g->mix[g->row_len + f] = R;
		return;
	}
	/* CallDesc::stereo: bit 0 = two channels, bit 1 = big-endian samples (the AU stream
	 * of `saugns -o -`, player/sndfile.c:160-168: the byte swap folded into the epilogue) */
	const bool be = (cd->stereo & 2u) != 0;
	if (cd->stereo & 1u) {                                         /* generator.c:795-810 */
		L = sau::fclampf(L, -1.f, 1.f);
		R = sau::fclampf(R, -1.f, 1.f);
		uint32_t w = ((uint32_t) (uint16_t) (short) __float2int_rn(L * 32767.f)) |
			((uint32_t) (uint16_t) (short) __float2int_rn(R * 32767.f) << 16);
		if (be) w = __byte_perm(w, 0u, 0x2301);
		reinterpret_cast<uint32_t*>(g->pcm)[f] = w;
	} else {                                                       /* generator.c:812-825 */
		float m = (L + R) * 0.5f;
		m = sau::fclampf(m, -1.f, 1.f);
		uint32_t w = (uint16_t) (short) __float2int_rn(m * 32767.f);
		if (be) w = __byte_perm(w, 0u, 0x3201);
		reinterpret_cast<uint16_t*>(g->pcm)[f] = (uint16_t) w;
	}
}

__global__ void __launch_bounds__(MIX_FRAMES + 32)
mix_kernel(const CallDesc *calls, const SegDesc *segs, uint32_t mode /*0 pcm, 1 float planes*/) {
	extern __shared__ __align__(128) unsigned char mix_smem_raw[];
	MixSmem &sm = *reinterpret_cast<MixSmem*>(mix_smem_raw);
	const CallDesc *cd = &calls[blockIdx.y];
	const GenDesc *g = cd->gen;
	const uint32_t f0 = blockIdx.x * MIX_FRAMES;
	if (f0 >= cd->call_len) return;
	const uint32_t tid = threadIdx.x;
	const bool producer = tid >= (uint32_t) MIX_FRAMES;          /* the last warp */
	const uint32_t f = f0 + (producer ? 0u : tid);
	const bool valid = !producer && f < cd->call_len;
	const uint32_t nlv = g->voice_end - g->voice_begin;
	const uint32_t tstride = g->row_stride;
	/* the segment holding each thread's frame */
	uint32_t si = 0;
	for (; si < cd->nseg; ++si) {
		const SegDesc sd = segs[cd->seg_off + si];
		if (f >= sd.start && f < sd.start + sd.len) break;
	}
	const bool in_seg = valid && si < cd->nseg;
	const uint32_t fi = in_seg ? f - segs[cd->seg_off + si].start : 0u;
	__shared__ uint32_t seg0, mixed, active;
	if (tid == 0) { seg0 = si; mixed = 0; active = 0; }
	__syncthreads();
	if (valid && si != seg0) mixed = 1;
	if (in_seg && fi < g->status[1 + si]) active = 1;
	if (tid == 0) {
		for (int st = 0; st < MIX_STAGES; ++st) {
			mbar_init(&sm.full[st], 1);                /* the producer's arrive.expect_tx */
			mbar_init(&sm.empty[st], MIX_CWARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	float L = 0.f, R = 0.f;
	if (mixed || seg0 >= cd->nseg) {
		/* an event boundary inside these frames: each thread walks its own segment's
		 * voice list straight from global memory (rare) */
		if (in_seg && fi < g->status[1 + si]) {
			const uint2 *vl = reinterpret_cast<const uint2*>(g->vlen + (size_t) si * nlv);
			for (uint32_t lv = 0; lv < nlv; ++lv) {
				const uint2 v = vl[lv];
				if (fi < v.x) {
					const size_t at = (size_t) lv * ROW_TILE + row_index(f, tstride);
					const float s = g->rows_s[at];
					const float r = (v.y == PAN_DYNAMIC) ? g->rows_r[at] : s * __uint_as_float(v.y);
					L = (L + s) - r;
					R = (R + s) + r;
				}
			}
		}
		if (valid) mix_store(g, cd, mode, f, L, R);
		return;
	}
	if (!active) {                       /* nothing was rendered for these frames */
		if (valid) mix_store(g, cd, mode, f, 0.f, 0.f);
		return;
	}
	const uint2 *vl = reinterpret_cast<const uint2*>(g->vlen + (size_t) seg0 * nlv);
	const uint32_t ntiles = (nlv + MIX_TV - 1) / MIX_TV;
	if (producer) {
		/* The producer warp: per stage, the VoiceSeg records (one lane each), then lane 0
		 * posts the transaction count and issues the bulk copies: the MIX_TV voices' s
		 * pieces are contiguous (one copy), r pieces only for moving pans. */
		const uint32_t lane = tid & 31u;
		const float *tile_s = g->rows_s + (size_t) blockIdx.x * tstride;
		/* the records are fetched three stages ahead of their use (their L2 latency
		 * would otherwise sit in this loop's critical path) */
		auto fetch = [&](uint32_t t) {
			const uint32_t v = t * MIX_TV + lane;
			return (lane < (uint32_t) MIX_TV && v < nlv) ? __ldg(vl + v) : make_uint2(0u, 0u);
		};
		uint2 pre0 = fetch(0), pre1 = fetch(1), pre2 = fetch(2);
		for (uint32_t t = 0; t < ntiles; ++t) {
			const uint32_t st = t % MIX_STAGES, v0 = t * MIX_TV;
			const uint32_t nv = nlv - v0 < (uint32_t) MIX_TV ? nlv - v0 : (uint32_t) MIX_TV;
			const uint2 info = pre0;
			pre0 = pre1; pre1 = pre2; pre2 = fetch(t + 3);
			if (t >= (uint32_t) MIX_STAGES) mbar_wait(&sm.empty[st], ((t / MIX_STAGES) - 1u) & 1u);
			const bool has = lane < nv;
			if (has) sm.vi[st][lane] = info;
			const uint32_t dynmask = __ballot_sync(FULL, has && info.y == PAN_DYNAMIC && info.x);
			if (lane == 0) sm.ndyn[st] = __popc(dynmask);
			__syncwarp();                      /* vi, ndyn written before lane 0's arrive publishes them */
			if (lane == 0) {
				const uint32_t piece = ROW_TILE * (uint32_t) sizeof(float);
				mbar_expect_tx(&sm.full[st], nv * piece);
				tma_bulk_g2s(&sm.s[st][0][0], tile_s + (size_t) v0 * ROW_TILE, nv * piece, &sm.full[st]);
			}
		}
		return;
	}
	const uint32_t fx = in_seg ? fi : 0xffffffffu;               /* frames outside take nothing */
	for (uint32_t t = 0; t < ntiles; ++t) {
		const uint32_t st = t % MIX_STAGES, v0 = t * MIX_TV;
		const uint32_t nv = nlv - v0 < (uint32_t) MIX_TV ? nlv - v0 : (uint32_t) MIX_TV;
		mbar_wait(&sm.full[st], (t / MIX_STAGES) & 1u);
		const float *sp = &sm.s[st][0][tid];
		const float *rp = g->rows_r + (size_t) blockIdx.x * tstride + (size_t) v0 * ROW_TILE + tid;
		const uint2 *ip = &sm.vi[st][0];
		if (nv == (uint32_t) MIX_TV && sm.ndyn[st] == 0) {
			/* the common tile: all pans stand still */
#pragma unroll
			for (int k = 0; k < MIX_TV; ++k) {
				const uint2 info = ip[k];
				const float s = fx < info.x ? sp[k * MIX_FRAMES] : 0.f;
				const float rr = s * __uint_as_float(info.y);
				L = (L + s) - rr;                              /* as compiled, Appendix B.3; */
				R = (R + s) + rr;                              /* adding 0 is exact */
			}
		} else {
			for (uint32_t k = 0; k < nv; ++k) {
				const uint2 info = ip[k];
				const bool on = fx < info.x;
				const float s = on ? sp[k * MIX_FRAMES] : 0.f;
				float rr;
				if (info.y == PAN_DYNAMIC) rr = on ? rp[k * MIX_FRAMES] : 0.f;
				else rr = s * __uint_as_float(info.y);
				L = (L + s) - rr;
				R = (R + s) + rr;
			}
		}
		__syncwarp();
		if ((tid & 31u) == 0) mbar_arrive(&sm.empty[st]);        /* this warp is done with the stage */
	}
	if (valid) mix_store(g, cd, mode, f, L, R);
}

/* float planes (already reduced over ranks) -> int16, for voice-sharded runs */
__global__ void planes_to_pcm_kernel(const float *mix, uint32_t plane_stride, uint32_t n,
		uint32_t stereo, int16_t *pcm) {
	const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n) return;
	float L = mix[f], R = mix[plane_stride + f];
	const bool be = (stereo & 2u) != 0;          /* flags as CallDesc::stereo */
	uint16_t *out = reinterpret_cast<uint16_t*>(pcm);
	auto put = [be](uint16_t *p, int v) {
		const uint16_t u = (uint16_t) (short) v;
		*p = be ? (uint16_t) ((u << 8) | (u >> 8)) : u;
	};
	if (stereo & 1u) {
		L = sau::fclampf(L, -1.f, 1.f);
		R = sau::fclampf(R, -1.f, 1.f);
		put(out + 2 * f, __float2int_rn(L * 32767.f));
		put(out + 2 * f + 1, __float2int_rn(R * 32767.f));
	} else {
		float m = sau::fclampf((L + R) * 0.5f, -1.f, 1.f);
		put(out + f, __float2int_rn(m * 32767.f));
	}
}

/* ---- arithmetic self-test ------------------------------------------------ *
 * The hand-expanded primitives of the fast path against the plain statements
 * they replace: div_scale_by_int vs IEEE `/` for every non-zero int32 divisor
 * (strided over the grid) and each wave's diff_scale; ftoi_lo32 vs the low
 * word of sau::ftoi64 over a float bit-pattern sweep. */
__global__ void selftest_kernel(const float *tables, unsigned long long *bad) {
	const WaveCoeffs *wc = reinterpret_cast<const WaveCoeffs*>(tables + NUM_WAVES * WAVE_LEN);
	const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t nth = (uint64_t) gridDim.x * blockDim.x;
	unsigned long long nbad = 0;
	for (uint64_t u = tid; u < 0x100000000ull; u += nth) {
		const int32_t d = (int32_t) (uint32_t) u;
		if (d != 0) {
			const uint32_t w = (uint32_t) (u % NUM_WAVES);
			const float ds = wc->diff_scale[w];
			const float want = ds / (float) d;
			const float got = div_scale_by_int(ds, d);
			if (__float_as_uint(want) != __float_as_uint(got)) ++nbad;
		}
		const float x = __uint_as_float((uint32_t) u);
		if ((uint32_t) sau::ftoi64(x) != ftoi_lo32(x)) ++nbad;
	}
	if (nbad) atomicAdd(bad, nbad);
}

cudaError_t launch_selftest(const float *d_tables, unsigned long long *d_bad, cudaStream_t stream) {
	selftest_kernel<<<148 * 8, 256, 0, stream>>>(d_tables, d_bad);
	return cudaGetLastError();
}

/* ---- host-callable launchers -------------------------------------------- */

/* wave_mask may carry CTAB_FLAG (coefficient tables in shared memory) */
size_t render_smem_bytes(uint32_t wave_mask, uint32_t nbufs, uint32_t nslots_ops, uint32_t nplan,
		uint32_t warps) {
	uint32_t nslots = 0;
	for (uint32_t w = 0; w < NUM_WAVES; ++w) if (wave_mask & (1u << w)) ++nslots;
	const size_t slot = (wave_mask & CTAB_FLAG) ? CTAB_WAVE_BYTES : TAB_STRIDE * sizeof(float);
	return 128 + (size_t) nslots * slot + (size_t) warps * warp_smem_bytes(nbufs, nslots_ops, nplan);
}

/* ---- per-index cubic coefficients of every wave table -------------------- *
 * sauWave_get_herp (wave.h:127-141, as compiled: sau::herp_poly) forms c1, c2,
 * c3 from the four taps around an index before it touches the phase fraction;
 * they depend on the index alone.  One thread per (wave, index) evaluates the
 * SAME expressions once; the fast path then loads {c3, c2} (doubles) with one
 * 128-bit and {c1, c0} with one 64-bit shared-memory load.  c1 = 0.5 * (s2 - s0)
 * and c0 = s1 are float values held as floats (24 bytes per index instead of
 * 32: two waves leave room for 32 warps per SM); the kernel checks that the
 * float form of c1 is exact and flags the table set otherwise (the runtime then
 * stays with the float tables).  Layout per wave: 2048 x {c3, c2}, then
 * 2048 x {c1, c0}. */
__global__ void coef_kernel(const float *tables, double *coefs, uint32_t *inexact) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= NUM_WAVES * WAVE_LEN) return;
	const uint32_t w = t / WAVE_LEN, i = t % WAVE_LEN;
	const float *lut = tables + w * WAVE_LEN;
	const float s0 = lut[(i - 1) & sau::WAVE_LENMASK], s1 = lut[i];
	const float s2 = lut[(i + 1) & sau::WAVE_LENMASK], s3 = lut[(i + 2) & sau::WAVE_LENMASK];
	double c1, c2, c3;
	sau::herp_coefs(s0, s1, s2, s3, &c1, &c2, &c3);
	unsigned char *base = reinterpret_cast<unsigned char*>(coefs) + (size_t) w * CTAB_WAVE_BYTES;
	reinterpret_cast<double2*>(base)[i] = make_double2(c3, c2);
	const float c1f = (float) c1;
	if ((double) c1f != c1) atomicOr(inexact, 1u);
	reinterpret_cast<float2*>(base + CTAB_PLANE_BYTES)[i] = make_float2(c1f, s1);
}
size_t coef_table_bytes() { return (size_t) NUM_WAVES * CTAB_WAVE_BYTES; }
cudaError_t launch_coefs(const float *d_tables, double *d_coefs, uint32_t *d_inexact, cudaStream_t stream) {
	coef_kernel<<<(NUM_WAVES * WAVE_LEN + 255) / 256, 256, 0, stream>>>(d_tables, d_coefs, d_inexact);
	return cudaGetLastError();
}

/* grid: one warp per voice task, or a persistent grid of `ticketed_ctas` CTAs with
 * sched_mode 1 = (unit, voice) tickets, 2 = balanced contiguous ranges */
static cudaError_t ensure_smem(bool wide, size_t smem) {
	int dev = 0;
	cudaGetDevice(&dev);
	static size_t configured[2][64] = {{0}, {0}};
	if (dev >= 0 && dev < 64 && smem > configured[wide][dev]) {
		cudaError_t e = wide ?
			cudaFuncSetAttribute(render_kernel_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) :
			cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if (e != cudaSuccess) return e;
		configured[wide][dev] = smem;
	}
	return cudaSuccess;
}
int render_ctas_per_sm(size_t smem, uint32_t warps) {
	int n = 0;
	if (ensure_smem(warps > 8, smem) != cudaSuccess) return 0;
	cudaError_t e = warps > 8 ?
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, render_kernel_wide, (int) warps * 32, smem) :
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, render_kernel, (int) warps * 32, smem);
	return e == cudaSuccess ? n : 0;
}
cudaError_t launch_render(const CallDesc *d_calls, uint32_t ncalls, const SegDesc *d_segs,
		const UnitDesc *d_units, uint32_t ntasks, const float *d_tables, const double *d_coefs,
		uint32_t wave_mask, uint32_t nbufs, uint32_t nslots_ops, uint32_t nplan, uint32_t warps,
		uint32_t ticketed_ctas, uint32_t sched_mode, cudaStream_t stream) {
	if (ntasks == 0) return cudaSuccess;
	if (nslots_ops == 0) nslots_ops = 1;
	const size_t smem = render_smem_bytes(wave_mask, nbufs, nslots_ops, nplan, warps);
	const bool wide = warps > 8;
	{
		cudaError_t e = ensure_smem(wide, smem);
		if (e != cudaSuccess) return e;
	}
	const uint32_t grid = ticketed_ctas ? ticketed_ctas : (ntasks + warps - 1) / warps;
	if (wide)
		render_kernel_wide<<<grid, warps * 32, smem, stream>>>(d_calls, ncalls, d_segs, d_units, ntasks,
				d_tables, d_coefs, wave_mask, nbufs, nslots_ops, nplan, warps, ticketed_ctas ? sched_mode : 0u);
	else
		render_kernel<<<grid, warps * 32, smem, stream>>>(d_calls, ncalls, d_segs, d_units, ntasks,
				d_tables, d_coefs, wave_mask, nbufs, nslots_ops, nplan, warps, ticketed_ctas ? sched_mode : 0u);
	return cudaGetLastError();
}

cudaError_t launch_mix(const CallDesc *d_calls, uint32_t ncalls, const SegDesc *d_segs,
		uint32_t max_call_len, uint32_t mode, cudaStream_t stream) {
	if (ncalls == 0 || max_call_len == 0) return cudaSuccess;
	dim3 grid((max_call_len + MIX_FRAMES - 1) / MIX_FRAMES, ncalls);
	static bool configured[64] = {false};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 0 && dev < 64 && !configured[dev]) {
		cudaError_t e = cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
				(int) sizeof(MixSmem));
		if (e != cudaSuccess) return e;
		configured[dev] = true;
	}
	mix_kernel<<<grid, MIX_FRAMES + 32, sizeof(MixSmem), stream>>>(d_calls, d_segs, mode);
	return cudaGetLastError();
}

cudaError_t launch_planes_to_pcm(const float *d_mix, uint32_t plane_stride, uint32_t n,
		uint32_t stereo, int16_t *d_pcm, cudaStream_t stream) {
	if (!n) return cudaSuccess;
	planes_to_pcm_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_mix, plane_stride, n, stereo, d_pcm);
	return cudaGetLastError();
}

} // namespace saugen
