/* sau_arith.h -- per-sample arithmetic of the saugns generator back end,
 * stated once for host (strict IEEE, no contraction) and device (-fmad=false).
 *
 * Every function is a restatement of reference arithmetic AS COMPILED by the
 * canonical oracle build (gcc 13.3, -O3 -ffast-math for generator.c / line.c /
 * wave.c): one rounded IEEE operation per
 * C operator, evaluated exactly in the order written here.  Where that order
 * departs from the reference's source text the comment says so; the orders
 * were read from `objdump -d` of oracle/_ref/exe_generator.o / exe_line.o and
 * are pinned bitwise by tests/test_oracle_port.py against oracle/_ref.
 *
 * Build rules: host -O2 -ffp-contract=off (never -ffast-math); device
 * -fmad=false (default -prec-div=true -prec-sqrt=true -ftz=false).
 */
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define SAU_HD __host__ __device__ __forceinline__
#else
#define SAU_HD static inline
#endif

namespace sau {

/* ---- integer helpers (sau/math.h:89-118,283-303) ------------------------ */

SAU_HD uint32_t ranfast32(uint32_t n) {            /* math.h:297-303 */
	uint32_t s = n * 0x9e3779b9u;
	s ^= s >> 14;
	s = (s | 1u) * s;
	s ^= s >> 13;
	return s;
}
SAU_HD uint32_t mcg32(uint32_t seed) { return seed * 0xe47135u; }   /* math.h:283 */
SAU_HD int32_t sar32(int32_t x, int s) {           /* math.h:94-96 */
	return x < 0 ? ~(~x >> s) : x >> s;
}
SAU_HD int32_t foldhd32(int32_t x) {               /* math.h:112-118 */
	uint32_t s = (uint32_t) x;
	if (s + (1u << 29) > (1u << 31))
		s = (1u << 31) + (1u << 30) - s;
	s = (s - (1u << 29)) * 2u;
	return (int32_t) s;
}
SAU_HD int oddness_as_sign(int n) { return 1 - ((n & 1) * 2); }     /* math.h:89 */
/* sau_divi: C signed division, rounds toward zero (generator.c:20) */
SAU_HD int32_t divi(int32_t i, int32_t d) { return i / d; }

/* sau_ftoi = (int64)lrintf(x), RNE, then wraps to the destination width
 * (generator.c:16-17, math.h:63-64).  x86 cvtss2si returns INT64_MIN for NaN
 * and out-of-range inputs; mirror that instead of CUDA's saturation. */
SAU_HD int64_t ftoi64(float x) {
#if defined(__CUDA_ARCH__)
	int64_t r = __float2ll_rn(x);
	if (!(fabsf(x) < 9223372036854775808.f)) r = (int64_t) 0x8000000000000000ull;
	return r;
#else
	if (!(fabsf(x) < 9223372036854775808.f)) return (int64_t) 0x8000000000000000ull;
	return (int64_t) llrintf(x);
#endif
}
SAU_HD float i2f(int32_t i) { return (float) i; }
SAU_HD float u2f(uint32_t u) { return (float) u; }
SAU_HD uint32_t fbits(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }

SAU_HD float fclampf(float x, float mn, float mx) {   /* math.h:133-137 */
	x = x < mn ? mn : x;
	x = x > mx ? mx : x;
	return x;
}
SAU_HD float minf_(float x, float y) { return x > y ? y : x; }       /* math.h:121 */
SAU_HD float maxf_(float x, float y) { return x < y ? y : x; }       /* math.h:127 */

/* sau_sinpi_d5f, math.h:366-379 (source order) */
SAU_HD float sinpi_d5f(float x) {
	const float s0 = +3.14042741234069229463f;
	const float s1 = -5.13655757476162831091f;
	const float s2 = +2.29939170159543653372f;
	float x2 = x * x;
	return x * (s0 + x2 * (s1 + x2 * s2));
}

/* ---- value lines (sau/line.h:153-266, sau/line.c:27-281) ---------------- */

enum { L_cos = 0, L_lin, L_sah, L_exp, L_log, L_xpe, L_lge, L_sqe, L_cub, L_smo,
       L_ncl, L_nhl, L_uwh, L_NAMED };

/* sau_expramp6 as compiled (line.h:195-200; Appendix B.1 "R6"):
 * source x3 + (x2*x3 - x2)*A was refactored to ((A*(x3-1))*x2) + x3. */
SAU_HD float expramp6(float x) {
	float x2 = x * x;
	float x3 = x * x2;
	float A = x2 * (1163.f / 1792.f) + x * (629.f / 1792.f);
	float B = x3 + (-1.f);
	return ((A * B) * x2) + x3;
}
/* sau_sinramp coefficients, line.h:174-183 */
#define SAU_SR0 (+1.5702137061703461473139223358864f)
#define SAU_SR1 (-2.568278787380814155456160152724f)
#define SAU_SR2 (+1.1496958507977182668618673644367f)
SAU_HD float sinramp(float x) {            /* source order (used by val_cos) */
	float x2 = x * x;
	return x * (SAU_SR0 + x2 * (SAU_SR1 + x2 * SAU_SR2));
}

/* Per-fill constants hoisted out of the sample loop. */
struct LineFill {
	float v0, vt, inv, vm, vd, c;   /* c: shape-specific constant */
	int32_t adj_pos;                /* pos - time/2 (line.c:83) */
	uint32_t pos;
	int type;                       /* resolved: exp/log already dispatched */
};

SAU_HD LineFill line_fill_setup(int type, float v0, float vt, uint32_t pos, uint32_t time) {
	LineFill f;
	if (type == L_exp) type = (v0 > vt) ? L_xpe : L_lge;        /* line.c:125-131 */
	else if (type == L_log) type = (v0 < vt) ? L_xpe : L_lge;   /* line.c:142-148 */
	f.type = type;
	f.v0 = v0; f.vt = vt; f.pos = pos;
	f.adj_pos = (int32_t) (pos - (time / 2));
	f.inv = 1.f / u2f(time);
	f.vm = (v0 + vt) * 0.5f;
	f.vd = vt - v0;
	f.c = 0.f;
	switch (type) {
	case L_lin: f.c = f.vd * f.inv; break;
	case L_xpe: f.c = v0 - vt; break;
	case L_sqe: f.c = v0 - vt; break;
	case L_cub: f.inv = -2.f * f.inv; f.c = (v0 - vt) * 0.5f; break;
	case L_uwh: f.c = f.vd * (0.5f / 2147483648.f); break;      /* line.c:228-230 */
	default: break;
	}
	return f;
}

/* Value number i of a fill (before the optional ratio multiply, which is
 * always the last operation: line.c:35,72,90).  `cub_tail` selects the
 * expression gcc emitted for the scalar tail of sauLine_fill_cub. */
SAU_HD float line_fill_at(const LineFill &f, uint32_t i, bool cub_tail) {
	switch (f.type) {
	default:
	case L_sah: return f.v0;
	case L_lin: {
		float k = i2f((int32_t) i + f.adj_pos);
		return (k * f.c) + f.vm; }
	case L_cos: {
		float x = i2f((int32_t) i + f.adj_pos) * f.inv;
		float x2 = x * x;
		float t = x * f.vd;
		float p = ((x2 * SAU_SR2 + SAU_SR1) * x2 + SAU_SR0);
		return p * t + f.vm; }
	case L_xpe: {
		float x = 1.f - u2f(i + f.pos) * f.inv;
		return expramp6(x) * f.c + f.vt; }
	case L_lge: {
		float x = u2f(i + f.pos) * f.inv;
		return expramp6(x) * f.vd + f.v0; }
	case L_sqe: {
		float x = 0.5f - i2f((int32_t) i + f.adj_pos) * f.inv;
		return ((x * x) * f.c) + f.vt; }
	case L_cub: {
		float x = i2f((int32_t) i + f.adj_pos) * f.inv;   /* inv = -2/time */
		float x3 = (x * x) * x;
		if (cub_tail) return (x3 * f.c + f.c) + f.vt;
		return ((x3 + 1.f) * f.c) + f.vt; }
	case L_smo: {
		float x = u2f(i + f.pos) * f.inv;
		float p = ((x * 6.f + (-15.f)) * x + 10.f);
		float q = (x * x) * (x * f.vd);
		return p * q + f.v0; }
	case L_uwh: {
		int32_t s = (int32_t) ranfast32(f.pos + i);
		return f.vm + f.c * i2f(s); }
	case L_ncl: {
		float x = i2f((int32_t) i + f.adj_pos) * f.inv;
		float xb = x + 0.5f;
		float t = ((xb + xb) + (-3.f)) * xb + 1.f;
		float u = xb * (0.5f / 2147483648.f);
		int32_t s = (int32_t) ranfast32(f.pos + i);
		float r = (i2f(s) * t) * u;
		return ((r + x) * f.vd) + f.vm; }
	case L_nhl: {
		float x = i2f((int32_t) i + f.adj_pos) * f.inv;
		float xb = x + 0.5f;
		float t = 1.f - xb;
		float u = xb * (1.f / 2147483648.f);
		int32_t s = (int32_t) ranfast32(f.pos + i);
		float r = (i2f(s) * t) * u;
		return ((r + x) * f.vd) + f.vm; }
	}
}

/* sauLine_val_* as compiled into line.o (line.h:153-266; Appendix B.2);
 * reached through sauLine_val_funcs / sauLine_map_funcs by the R oscillator. */
SAU_HD float line_val(int type, float x, float a, float b, bool cub_tail) {
	float d = b - a;
	if (type == L_exp) type = (a > b) ? L_xpe : L_lge;
	else if (type == L_log) type = (a < b) ? L_xpe : L_lge;
	switch (type) {
	default:
	case L_sah: return a;
	case L_lin: return a + d * x;
	case L_cos: return a + d * (sinramp(x - 0.5f) + 0.5f);
	case L_xpe: return expramp6(1.f - x) * (a - b) + b;
	case L_lge: return expramp6(x) * d + a;
	case L_sqe: { float y = 1.f - x; return b + (a - b) * (y * y); }
	case L_cub: {
		float y = (0.5f - x) * 2.f;
		float y3 = (y * y) * y;
		if (cub_tail) { float h = (a - b) * 0.5f; return (y3 * h + h) + b; }
		return b + (a - b) * (y3 * 0.5f + 0.5f); }
	case L_smo: {
		float p = ((x * 6.f - 15.f) * x + 10.f);
		return p * ((d * x) * (x * x)) + a; }
	case L_uwh: {
		int32_t s = (int32_t) ranfast32(fbits(x));
		return a + d * (0.5f + (0.5f * (1.f / 2147483648.f)) * i2f(s)); }
	case L_ncl: {
		int32_t s = (int32_t) ranfast32(fbits(x));
		float t = ((x + x) - 3.f) * x + 1.f;
		float r = (i2f(s) * t) * ((0.5f * (1.f / 2147483648.f)) * x);
		return (x + r) * d + a; }
	case L_nhl: {
		int32_t s = (int32_t) ranfast32(fbits(x));
		float r = (i2f(s) * (1.f - x)) * ((1.f / 2147483648.f) * x);
		return (x + r) * d + a; }
	}
}

/* perlin_amp per line type, line.h:18-32 */
SAU_HD float line_perlin_amp(int type) {
	switch (type) {
	case L_sah: case L_uwh: return 1.f;
	case L_exp: case L_log: case L_xpe: case L_lge: return 1.55845810035f;
	case L_sqe: case L_nhl: return 1.89339094650f;
	default: return 2.f;
	}
}

/* ---- wave oscillator (sau/wave.h:127-141, sau/generator/wosc.h) --------- */

constexpr int WAVE_LENBITS = 11, WAVE_LENMASK = 2047;
constexpr int WAVE_SLENBITS = 21;
constexpr uint32_t WAVE_SLEN = 1u << 21, WAVE_SLENMASK = (1u << 21) - 1;

/* sauWave_get_herp as compiled: c2 is associated (s0-2.5*s1)+(2*s2-0.5*s3)
 * (source: left to right), the Horner steps are as written in the source.
 * Returns the polynomial part; the caller adds c0 (needed apart by reset). */
SAU_HD void herp_coefs(float s0, float s1, float s2, float s3, double *c1, double *c2, double *c3) {
	*c1 = 0.5 * (double) (s2 - s0);
	*c2 = ((double) s0 - 2.5 * (double) s1) + ((double) (s2 + s2) - 0.5 * (double) s3);
	*c3 = 0.5 * (double) (s3 - s0) + 1.5 * (double) (s1 - s2);
}
SAU_HD double herp_horner(double c3, double c2, double c1, uint32_t phase) {
	float xf = u2f(phase & WAVE_SLENMASK) * (1.f / 2097152.f);
	double x = (double) xf;
	return ((c3 * x + c2) * x + c1) * x;
}
SAU_HD double herp_poly(float s0, float s1, float s2, float s3, uint32_t phase) {
	double c1, c2, c3;
	herp_coefs(s0, s1, s2, s3, &c1, &c2, &c3);
	return herp_horner(c3, c2, c1, phase);
}
template <typename LutPtr>
SAU_HD double herp(LutPtr lut, uint32_t phase, double *poly_out, double *c0_out) {
	uint32_t ind = phase >> WAVE_SLENBITS;
	float s0 = lut[(ind - 1) & WAVE_LENMASK];
	float s1 = lut[ind];
	float s2 = lut[(ind + 1) & WAVE_LENMASK];
	float s3 = lut[(ind + 2) & WAVE_LENMASK];
	double p = herp_poly(s0, s1, s2, s3, phase);
	if (poly_out) { *poly_out = p; *c0_out = (double) s1; }
	return p + (double) s1;
}
/* One differentiated output sample, wosc.h:254-256 (float divide, then double). */
SAU_HD float wosc_diff(double Is, double prev_Is, int32_t phase_diff,
		float diff_scale, float diff_offset) {
	float xq = diff_scale / i2f(phase_diff);
	return (float) ((Is - prev_Is) * (double) xq + (double) diff_offset);
}
SAU_HD float wave_dvscale(float amp_scale) {       /* wave.h:144-145 */
	return amp_scale * 0.125f * 4294967296.f;
}

/* phase offsets, wosc.h:141-166 / rasg.h:172-215 as compiled (Appendix B.3):
 * fPM alone folds fpm_scale*2^31 into one constant; PM+fPM multiplies
 * (fpm*f) by fpm_scale before adding pm. */
#define SAU_FPM_SCALE ((float) (1.0 / 632.45553203367586639978))
SAU_HD int64_t pofs_pm(float pm, float phase_scale) { return ftoi64(pm * phase_scale); }
SAU_HD int64_t pofs_fpm(float fpm, float f, float phase_scale) {
	return ftoi64((fpm * f) * (SAU_FPM_SCALE * phase_scale));
}
SAU_HD int64_t pofs_pm_fpm(float pm, float fpm, float f, float phase_scale) {
	return ftoi64((((fpm * f) * SAU_FPM_SCALE) + pm) * phase_scale);
}

/* ---- noise (sau/generator/noise.h) -------------------------------------- */

SAU_HD float soft_sqrtm2logp1_2_r01(float x) {     /* noise.h:61-70 */
	const float s0 = -0.80270565422983103084f;
	const float s1 = +5.52274428214641442648f;
	const float s2 = -138.87126103150588693697f;
	float x2 = x * x;
	float x4 = x2 * x2;
	return 0.5f + x * (s0 + x4 * (s1 + x4 * s2));
}
SAU_HD float ssgauss_dist4(float x) {              /* noise.h:77-81 */
	float x2 = x * x;
	float gx = (x + x2) * 0.5f;
	return x * (1.f - gx * (1.f - x2));
}
SAU_HD float franssgauss32(uint32_t n) {           /* noise.h:90-98 */
	int32_t s0 = (int32_t) ranfast32(n);
	int32_t s1 = (int32_t) mcg32((uint32_t) s0);
	float a = (float) ((double) s0 * 0x1p-32);
	float b = (float) ((double) s1 * 0x1p-32);
	float c = ssgauss_dist4(soft_sqrtm2logp1_2_r01(a));
	return c * sinpi_d5f(b);
}

/* ---- random segments oscillator (sau/generator/rasg.h:299-671) ---------- */

enum { RAS_F_URAND = 0, RAS_F_GAUSS, RAS_F_BIN, RAS_F_TERN, RAS_F_FIXED, RAS_F_ADDREC };
enum { RAS_O_PERLIN = 1, RAS_O_HALFSHAPE = 2, RAS_O_ZIGZAG = 4, RAS_O_SQUARE = 8,
       RAS_O_VIOLET = 16 };

SAU_HD float fscalei(uint32_t i, float scale) { return i2f((int32_t) i) * scale; } /* generator.c:19 */

/* Segment end values a (this cycle) and b (next cycle) for one sample.
 * Dispatch mirrors sauRasG_map_* and their _s twins (same RASG_MAP_* bodies). */
SAU_HD void rasg_ends(unsigned func, unsigned flags, int sr, uint32_t alpha,
		uint32_t cycle, float &a, float &b) {
	const float sc = 1.f / 2147483648.f;   /* 0x1p-31f */
	switch (func) {
	default:
	case RAS_F_URAND:
		if (flags & RAS_O_VIOLET) {                        /* rasg.h:307-312 */
			uint32_t s0 = ranfast32(cycle - 1) / 2;
			uint32_t s1 = ranfast32(cycle) / 2;
			uint32_t s2 = ranfast32(cycle + 1) / 2;
			a = fscalei(s1 - s0, sc);
			b = fscalei(s2 - s1, sc);
		} else {                                           /* rasg.h:336-338 */
			a = fscalei(ranfast32(cycle), sc);
			b = fscalei(ranfast32(cycle + 1), sc);
		}
		break;
	case RAS_F_GAUSS:                                      /* rasg.h:376-378 */
		a = franssgauss32(cycle);
		b = franssgauss32(cycle + 1);
		break;
	case RAS_F_BIN:
		if (flags & RAS_O_VIOLET) {                        /* rasg.h:398-415 */
			const float scale_diff = 1.f - (i2f(sar32(INT32_MAX, sr)) / 2147483648.f);
			const float scale = (1.f + scale_diff * scale_diff) / 2147483648.f;
			uint32_t sb = (cycle & 1) << 31;
			uint32_t sb_flip = (1u << 31) - sb;
			uint32_t s0 = (uint32_t) divi((int32_t) ((uint32_t) sar32((int32_t) ranfast32(cycle - 1), sr) + sb), 2);
			uint32_t s1 = (uint32_t) divi((int32_t) ((uint32_t) sar32((int32_t) ranfast32(cycle), sr) + sb_flip), 2);
			uint32_t s2 = (uint32_t) divi((int32_t) ((uint32_t) sar32((int32_t) ranfast32(cycle + 1), sr) + sb), 2);
			a = fscalei(s1 - s0, scale);
			b = fscalei(s2 - s1, scale);
		} else {                                           /* rasg.h:459-464 */
			uint32_t offs = (uint32_t) INT32_MAX + (cycle & 1) * 2;
			uint32_t s1 = (uint32_t) sar32((int32_t) ranfast32(cycle), sr) + offs;
			uint32_t s2 = (uint32_t) sar32((int32_t) ranfast32(cycle + 1), sr) - offs;
			a = fscalei(s1, sc);
			b = fscalei(s2, sc);
		}
		break;
	case RAS_F_TERN: {                                     /* rasg.h:509-516 */
		uint32_t sb = (cycle & 1) << 31;
		uint32_t sb_flip = (1u << 31) - sb;
		uint32_t s1 = (uint32_t) sar32((int32_t) ranfast32(cycle), sr) + sb_flip;
		uint32_t s2 = (uint32_t) sar32((int32_t) ranfast32(cycle + 1), sr) + sb;
		a = fscalei(s1, sc);
		b = fscalei(s2, sc);
		break; }
	case RAS_F_FIXED:
		if (sr >= 27) {                                    /* rasg.h:538-540,595 */
			a = i2f(oddness_as_sign((int) cycle));
			b = -a;
		} else if (flags & RAS_O_VIOLET) {                 /* rasg.h:563-575 */
			uint32_t sign = (uint32_t) oddness_as_sign((int) cycle);
			uint32_t s0 = (uint32_t) divi((int32_t) (sign * ((ranfast32(cycle - 1) >> sr) - (uint32_t) INT32_MAX)), 2);
			uint32_t s1 = (uint32_t) divi((int32_t) ((0u - sign) * ((ranfast32(cycle) >> sr) - (uint32_t) INT32_MAX)), 2);
			uint32_t s2 = (uint32_t) divi((int32_t) (sign * ((ranfast32(cycle + 1) >> sr) - (uint32_t) INT32_MAX)), 2);
			a = fscalei(s1 - s0, sc);
			b = fscalei(s2 - s1, sc);
		} else {                                           /* rasg.h:608-615 */
			uint32_t sign = (uint32_t) oddness_as_sign((int) cycle);
			a = fscalei((0u - sign) * ((ranfast32(cycle) >> sr) - (uint32_t) INT32_MAX), sc);
			b = fscalei(sign * ((ranfast32(cycle + 1) >> sr) - (uint32_t) INT32_MAX), sc);
		}
		break;
	case RAS_F_ADDREC: {                                   /* rasg.h:659-663 */
		uint32_t s0 = cycle * alpha;
		uint32_t s1 = (cycle + 1) * alpha;
		a = fscalei(s0, sc);
		b = fscalei(s1, sc);
		break; }
	}
}

/* franssgauss32 split into its two factors (result = c * sp). */
SAU_HD void franssgauss32_parts(uint32_t n, float &c, float &sp) {
	int32_t s0 = (int32_t) ranfast32(n);
	int32_t s1 = (int32_t) mcg32((uint32_t) s0);
	float a = (float) ((double) s0 * 0x1p-32);
	float b = (float) ((double) s1 * 0x1p-32);
	c = ssgauss_dist4(soft_sqrtm2logp1_2_r01(a));
	sp = sinpi_d5f(b);
}

/* One output sample of the R oscillator: end values, Perlin scaling, half-
 * shape sort, zig-zag swap, squaring, line mapping (rasg.h:242-280,692-734).
 * `self` selects the self-PM loop (RASG_MAP_S_LOOP), whose Perlin products gcc
 * associated differently per function (read from the disassembly of
 * sauRasG_map_*_s; Appendix B.3 covers the generic case only):
 *   generic : a = (a*phase)*pamp,            b = (b*(phase-1))*pamp
 *   gauss   : a = (sp_a*(pamp*phase))*c_a,   b = (sp_b*((phase-1)*pamp))*c_b
 *   v_bin   : a = (i_a*phase)*(pamp*scale),  b = (i_b*(phase-1))*(pamp*scale)
 * and the block path (sauRasG_run): a = (a*phase)*pamp, b = (b*pamp)*(phase-1). */
SAU_HD float rasg_sample(unsigned func, unsigned flags, int sr, uint32_t alpha, int line,
		uint32_t cycle, float phase, bool self, bool cub_tail) {
	float a, b;
	if (flags & RAS_O_PERLIN) {
		const float pamp = (flags & (RAS_O_HALFSHAPE | RAS_O_ZIGZAG)) ?
			1.f : line_perlin_amp(line);
		if (self && func == RAS_F_GAUSS) {
			float ca, spa, cb, spb;
			franssgauss32_parts(cycle, ca, spa);
			franssgauss32_parts(cycle + 1, cb, spb);
			a = (spa * (pamp * phase)) * ca;
			b = (spb * ((phase - 1.f) * pamp)) * cb;
		} else if (self && func == RAS_F_BIN && (flags & RAS_O_VIOLET)) {
			const float scale_diff = 1.f - (i2f(sar32(INT32_MAX, sr)) / 2147483648.f);
			const float scale = (1.f + scale_diff * scale_diff) / 2147483648.f;
			const float K = pamp * scale;
			uint32_t sb = (cycle & 1) << 31;
			uint32_t sb_flip = (1u << 31) - sb;
			uint32_t s0 = (uint32_t) divi((int32_t) ((uint32_t) sar32((int32_t) ranfast32(cycle - 1), sr) + sb), 2);
			uint32_t s1 = (uint32_t) divi((int32_t) ((uint32_t) sar32((int32_t) ranfast32(cycle), sr) + sb_flip), 2);
			uint32_t s2 = (uint32_t) divi((int32_t) ((uint32_t) sar32((int32_t) ranfast32(cycle + 1), sr) + sb), 2);
			a = (i2f((int32_t) (s1 - s0)) * phase) * K;
			b = (i2f((int32_t) (s2 - s1)) * (phase - 1.f)) * K;
		} else {
			rasg_ends(func, flags, sr, alpha, cycle, a, b);
			if (self) {
				a = (a * phase) * pamp;
				b = (b * (phase - 1.f)) * pamp;
			} else {
				a = (a * phase) * pamp;
				b = (b * pamp) * (phase + (-1.f));
			}
		}
	} else {
		rasg_ends(func, flags, sr, alpha, cycle, a, b);
	}
	if (flags & RAS_O_HALFSHAPE) {                         /* rasg.h:259-265,712-720 */
		float mx = maxf_(a, b), mn = minf_(a, b);
		a = mx; b = mn;
	}
	if (flags & RAS_O_ZIGZAG) { float t = a; a = b; b = t; }   /* rasg.h:266-269 */
	if (flags & RAS_O_SQUARE) { a *= fabsf(a); b *= fabsf(b); } /* rasg.h:270-274 */
	return line_val(line, phase, a, b, cub_tail);
}

} // namespace sau
