/* oracle/saugen_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement ("port") of the saugns generator back end
 * (reference sau/generator.c + sau/generator/{wosc,rasg,noise}.h + sau/line.c),
 * block-structured like the reference (1024-sample blocks, recursive operator
 * walk over shared work buffers).  Control flow is stated here independently
 * of the CUDA product (which uses per-voice bytecode and 128-sample chunks);
 * the per-sample arithmetic comes from saugns_b200/csrc/sau_arith.h, the one
 * statement of the as-compiled operation orders shared with the device code.
 *
 * PINNING: tests/test_oracle_port.py checks this port bit-for-bit (PCM and
 * integer/float operator state) against the UNMODIFIED reference built in
 * oracle/_ref by oracle/Makefile, on all of the reference's example scripts
 * and on feature scripts; parity of this oracle is therefore pinned by the
 * live reference, which holds no golden vectors of its own.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load the library built from this file (oracle/_ref/liboracle.so).
 * Build: g++ -O2 -ffp-contract=off (no -ffast-math).
 */
#include "../saugns_b200/csrc/sau_arith.h"
#include "../include/sau_program_abi.h"
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

constexpr uint32_t BUF_LEN = 1024;                  /* generator.c:28 */
typedef float Buf[BUF_LEN];

enum { ON_INIT = 1, ON_VISITED = 2, ON_TIME_INF = 4 };   /* generator.c:39-43 */
enum { VN_INIT = 1 };
enum { OSC_RESET_DIFF = 1 };

struct Line { float v0, vt; uint32_t pos, end; uint8_t type, flags; };

struct Tables {
	float pilut[SAUABI_WAVE_NAMED][SAUABI_WAVE_LEN];
	float amp_scale[SAUABI_WAVE_NAMED], amp_dc[SAUABI_WAVE_NAMED];
	int32_t phase_adj[SAUABI_WAVE_NAMED];
};

struct Op {                                         /* OperatorNode, generator.c:45-88 */
	uint32_t time; uint8_t type, flags;
	Line amp, amp2, pan, freq, freq2, pm_a;
	const sauabi_ProgramIDArr *amods, *ramods, *camods, *fmods, *rfmods,
		*pmods, *fpmods, *apmods;
	/* sauWOsc, wosc.h:32-50 */
	uint32_t phase, prev_phase; uint8_t wave, oscflags;
	double prev_Is; float prev_s, fb_s; float coeff;
	/* sauRasG, rasg.h:29-39 */
	uint64_t cycle_phase; bool rate2x;
	uint8_t ras_line; unsigned ras_flags, ras_func, ras_level; uint32_t ras_alpha;
	/* sauNoiseG, noise.h:23-27 */
	uint32_t n, nprev; uint8_t ntype;
};

struct Voice { uint32_t duration; uint8_t flags, freq_buf_id; uint32_t carr_op_id; };

struct Gen {
	const sauabi_Program *prg;
	Tables tab;
	uint32_t srate;
	bool out_clear;
	uint32_t mix_add_max;
	std::vector<float> gen_bufs_store;
	Buf *gen_bufs; Buf mix_l, mix_r;
	size_t event; std::vector<uint32_t> ev_wait;
	uint32_t event_pos;
	uint32_t voice;
	std::vector<Voice> voices;
	std::vector<Op> ops;
	float amp_scale;
	uint32_t block_pos;   /* offset of the current block in its run_for_time */
};

const sauabi_ProgramIDArr blank_idarr = {0};

uint32_t ms_in_samples(uint64_t ms, uint64_t srate, int *carry) {   /* math.h:35-46 */
	uint64_t t = ms * srate;
	if (carry) { t += *carry; *carry = (int) (t % 1000); }
	return (uint32_t) (t / 1000);
}

/* ---- sauLine state machine, line.c:287-473 ---- */

uint32_t line_get(Line *o, float *buf, uint32_t buf_len, const float *mulbuf) {  /* line.c:349-378 */
	if (!(o->flags & SAUABI_LINEP_GOAL)) return 0;
	if (o->flags & SAUABI_LINEP_GOAL_RATIO) {
		if (!(o->flags & SAUABI_LINEP_STATE_RATIO)) {
			if (mulbuf) o->v0 /= mulbuf[0];
			o->flags |= SAUABI_LINEP_STATE_RATIO;
		}
	} else {
		if (o->flags & SAUABI_LINEP_STATE_RATIO) {
			if (mulbuf) o->v0 *= mulbuf[0];
			o->flags &= ~SAUABI_LINEP_STATE_RATIO;
		}
		mulbuf = NULL;
	}
	if (o->pos >= o->end) return 0;
	uint32_t len = o->end - o->pos;
	if (len > buf_len) len = buf_len;
	sau::LineFill f = sau::line_fill_setup(o->type, o->v0, o->vt, o->pos, o->end);
	for (uint32_t i = 0; i < len; ++i) {
		/* gcc's scalar tail of sauLine_fill_cub: last element of an odd-length fill */
		bool tail = (len & 1) && i == len - 1;
		float v = sau::line_fill_at(f, i, tail);
		buf[i] = mulbuf ? v * mulbuf[i] : v;
	}
	return len;
}
bool line_advance(Line *o, uint32_t buf_len) {      /* line.c:385-398 */
	if (o->pos < o->end) {
		uint32_t len = o->end - o->pos;
		if (len > buf_len) len = buf_len;
		o->pos += len;
	}
	if (o->pos >= o->end) {
		o->pos = 0;
		o->flags &= ~SAUABI_LINEP_TIME;
		return false;
	}
	return true;
}
void line_run(Line *o, float *buf, uint32_t buf_len, const float *mulbuf) {   /* line.c:417-445 */
	uint32_t len = 0;
	bool fill = false;
	if (!(o->flags & SAUABI_LINEP_GOAL)) {
		line_advance(o, buf_len);
		fill = true;
	} else {
		len = line_get(o, buf, buf_len, mulbuf);
		o->pos += len;
		if (o->pos >= o->end) {
			o->v0 = o->vt;
			o->pos = 0;
			o->flags &= ~(SAUABI_LINEP_GOAL | SAUABI_LINEP_GOAL_RATIO | SAUABI_LINEP_TIME);
			fill = true;
		}
	}
	if (fill) {
		if (!(o->flags & SAUABI_LINEP_STATE_RATIO)) mulbuf = NULL;
		else if (mulbuf) mulbuf += len;
		for (uint32_t i = 0; i < buf_len - len; ++i)
			buf[len + i] = mulbuf ? o->v0 * mulbuf[i] : o->v0;
	}
}
void line_skip(Line *o, uint32_t skip_len) {        /* line.c:456-473 */
	if (!line_advance(o, skip_len)) {
		if (!(o->flags & SAUABI_LINEP_GOAL)) return;
		o->v0 = o->vt;
		if (o->flags & SAUABI_LINEP_GOAL_RATIO) o->flags |= SAUABI_LINEP_STATE_RATIO;
		else o->flags &= ~SAUABI_LINEP_STATE_RATIO;
		o->flags &= ~(SAUABI_LINEP_GOAL | SAUABI_LINEP_GOAL_RATIO);
	}
}
void line_copy(Line *o, const sauabi_Line *src, uint32_t srate) {   /* line.c:287-332 */
	if (!src) return;
	uint8_t mask = 0;
	if (src->flags & SAUABI_LINEP_STATE) {
		o->v0 = src->v0;
		mask |= SAUABI_LINEP_STATE | SAUABI_LINEP_STATE_RATIO;
	} else if (o->flags & SAUABI_LINEP_GOAL) {
		if (src->flags & SAUABI_LINEP_GOAL) {
			float f;
			line_get(o, &f, 1, NULL);
			o->v0 = f;
		}
	}
	if (src->flags & SAUABI_LINEP_GOAL) {
		o->vt = src->vt;
		if (src->flags & SAUABI_LINEP_TIME_IF_NEW) o->end -= o->pos;
		o->pos = 0;
		mask |= SAUABI_LINEP_GOAL | SAUABI_LINEP_GOAL_RATIO;
	}
	if (src->flags & SAUABI_LINEP_TYPE) {
		o->type = src->type;
		mask |= SAUABI_LINEP_TYPE;
	}
	if (!(o->flags & SAUABI_LINEP_TIME) || !(src->flags & SAUABI_LINEP_TIME_IF_NEW)) {
		if (src->flags & SAUABI_LINEP_TIME) {
			o->end = ms_in_samples(src->time_ms, srate, NULL);
			mask |= SAUABI_LINEP_TIME;
		}
	}
	o->flags &= ~mask;
	o->flags |= (src->flags & mask);
}

/* ---- R oscillator setters, rasg.h:59-119 ---- */

uint32_t ras_get_cycle(Op *o) { return (uint32_t) (o->cycle_phase >> 32) & ~1u; }
uint32_t ras_get_phase(Op *o) {
	return o->rate2x ? (uint32_t) (o->cycle_phase >> 1) : (uint32_t) o->cycle_phase;
}
void ras_set_cycle(Op *o, uint32_t cycle) {
	uint32_t phase = ras_get_phase(o);
	uint64_t phase64 = o->rate2x ? ((uint64_t) phase) << 1 : phase;
	o->cycle_phase = ((uint64_t) (cycle & ~1u)) << 32 | phase64;
}
void ras_set_phase(Op *o, uint32_t phase) {
	uint32_t cycle = ras_get_cycle(o);
	uint64_t phase64 = o->rate2x ? ((uint64_t) phase) << 1 : phase;
	o->cycle_phase = ((uint64_t) cycle) << 32 | phase64;
}
void ras_set_opt(Op *o, const sauabi_RasOpt *opt) {
	unsigned flags = opt->flags;
	if (opt->flags & SAUABI_RAS_O_LINE_SET) o->ras_line = opt->line;
	if (opt->flags & SAUABI_RAS_O_FUNC_SET) o->ras_func = opt->func;
	else flags |= o->ras_flags;
	if (opt->flags & SAUABI_RAS_O_LEVEL_SET) o->ras_level = opt->level;
	if (opt->flags & SAUABI_RAS_O_ASUBVAL_SET) o->ras_alpha = opt->alpha;
	o->ras_flags = flags & 0x3ff;
	bool rate2x = !(flags & SAUABI_RAS_O_HALFSHAPE);
	if (rate2x != o->rate2x) {
		uint32_t cycle = ras_get_cycle(o);
		uint32_t phase = ras_get_phase(o);
		o->rate2x = rate2x;
		ras_set_cycle(o, cycle);
		ras_set_phase(o, phase);
	}
}

/* ---- events, generator.c:233-377 ---- */

void prepare_op(Gen *g, Op *n, Voice *vn, const sauabi_ProgramOpData *od) {
	if (od->use_type == SAUABI_POP_carr && vn) vn->freq_buf_id = 0;
	memset((void*) n, 0, sizeof(*n));
	float coeff = (float) (4294967296.0 / g->srate);          /* wosc.h:30, math.h:386 */
	switch (od->type) {
	case SAUABI_POPT_wave:                                    /* wosc.h:55-71 */
		n->phase = (uint32_t) g->tab.phase_adj[SAUABI_WAVE_sin];
		n->coeff = coeff;
		n->wave = SAUABI_WAVE_sin;
		n->oscflags = OSC_RESET_DIFF;
		if (od->use_type == SAUABI_POP_carr && vn) vn->freq_buf_id = 2;
		break;
	case SAUABI_POPT_raseg:                                   /* rasg.h:44-57 */
		n->coeff = coeff;
		n->rate2x = true;
		n->ras_line = SAUABI_LINE_lin;
		n->ras_func = SAUABI_RAS_F_URAND;
		n->ras_level = 27;   /* sau_ras_level(9), program.h:146-148 */
		n->ras_alpha = 0x9e3779b9u;
		if (od->use_type == SAUABI_POP_carr && vn) vn->freq_buf_id = 3;
		break;
	}
	n->fmods = n->rfmods = n->pmods = n->fpmods = n->apmods = &blank_idarr;
	n->amods = n->ramods = n->camods = &blank_idarr;
	n->type = od->type;
	n->flags = ON_INIT;
}

void update_op(Gen *g, Op *n, const sauabi_ProgramOpData *od) {
	uint32_t params = od->params;
	bool osc = false;
	switch (od->type) {
	case SAUABI_POPT_noise:
		if (params & SAUABI_POPP_MODE) { n->ntype = od->mode.main; n->nprev = 0; }  /* noise.h:33-36 */
		if (params & SAUABI_POPP_SEED) n->n = od->seed;
		break;
	case SAUABI_POPT_wave:
		if (params & SAUABI_POPP_MODE) {                      /* wosc.h:81-87 */
			uint8_t wave = od->mode.main;
			n->phase += (uint32_t) g->tab.phase_adj[wave] - (uint32_t) g->tab.phase_adj[n->wave];
			n->wave = wave;
			n->oscflags |= OSC_RESET_DIFF;
		}
		if (params & SAUABI_POPP_PHASE)                       /* wosc.h:73-75 */
			n->phase = od->phase + (uint32_t) g->tab.phase_adj[n->wave];
		osc = true;
		break;
	case SAUABI_POPT_raseg:
		if (params & SAUABI_POPP_MODE) ras_set_opt(n, &od->mode.ras);
		if (params & SAUABI_POPP_PHASE) ras_set_phase(n, od->phase);
		if (params & SAUABI_POPP_SEED) ras_set_cycle(n, od->seed);
		osc = true;
		break;
	}
	if (osc) {
		if (od->fmods) n->fmods = od->fmods;
		if (od->rfmods) n->rfmods = od->rfmods;
		if (od->pmods) n->pmods = od->pmods;
		if (od->apmods) n->apmods = od->apmods;
		if (od->fpmods) n->fpmods = od->fpmods;
		line_copy(&n->freq, od->freq, g->srate);
		line_copy(&n->freq2, od->freq2, g->srate);
		line_copy(&n->pm_a, od->pm_a, g->srate);
	}
	if (params & SAUABI_POPP_TIME) {
		if (od->time.flags & SAUABI_TIMEP_IMPLICIT) {
			n->time = 0;
			n->flags |= ON_TIME_INF;
		} else {
			n->time = ms_in_samples(od->time.v_ms, g->srate, NULL);
			n->flags &= ~ON_TIME_INF;
		}
	}
	if (od->camods) n->camods = od->camods;
	if (od->amods) n->amods = od->amods;
	if (od->ramods) n->ramods = od->ramods;
	line_copy(&n->amp, od->amp, g->srate);
	line_copy(&n->amp2, od->amp2, g->srate);
	line_copy(&n->pan, od->pan, g->srate);
}

void handle_event(Gen *g, const sauabi_ProgramEvent *pe) {
	Voice *vn = NULL;
	if (pe->vo_id != SAUABI_PVO_NO_ID) vn = &g->voices[pe->vo_id];
	for (size_t i = 0; i < pe->op_data_count; ++i) {
		const sauabi_ProgramOpData *od = &pe->op_data[i];
		Op *n = &g->ops[od->id];
		if (!(n->flags & ON_INIT)) prepare_op(g, n, vn, od);
		update_op(g, n, od);
	}
	if (vn) {
		vn->carr_op_id = pe->carr_op_id;
		vn->flags |= VN_INIT;
		if (g->voice > pe->vo_id) g->voice = pe->vo_id;
		vn->duration = g->ops[vn->carr_op_id].time;           /* generator.c:233-240 */
	}
}

/* ---- block mixing, generator.c:384-440 ---- */

void block_mix(float *buf, uint32_t len, bool wave_env, bool layer,
		const float *in, const float *amp) {
	if (!wave_env) {
		if (layer) for (uint32_t i = 0; i < len; ++i) buf[i] += in[i] * amp[i];
		else for (uint32_t i = 0; i < len; ++i) buf[i] = in[i] * amp[i];
	} else {
		for (uint32_t i = 0; i < len; ++i) {
			float s = in[i];
			float s_amp = amp[i] * 0.5f;
			s = (s * s_amp) + fabsf(s_amp);
			if (layer) buf[i] *= s; else buf[i] = s;
		}
	}
}

uint32_t run_block(Gen *g, Buf *bufs, uint32_t buf_len, Op *n, float *parent_freq,
		bool wave_env, bool layer);

struct Par { Line *par, *r_par; const sauabi_ProgramIDArr *mods, *r_mods; };

void run_param(Gen *g, Buf *bufs, uint32_t len, Par p, float *param_mulbuf,
		float *reused_freq, bool is_freq) {                   /* generator.c:448-477 */
	float *par_buf = bufs[0];
	float *freq = reused_freq ? reused_freq : is_freq ? par_buf : NULL;
	line_run(p.par, par_buf, len, param_mulbuf);
	if (p.r_mods->count > 0) {
		float *r_par_buf = bufs[1];
		line_run(p.r_par, r_par_buf, len, param_mulbuf);
		for (uint32_t i = 0; i < p.r_mods->count; ++i)
			run_block(g, bufs + 2, len, &g->ops[p.r_mods->ids[i]], freq, true, i);
		float *mod_buf = bufs[2];
		for (uint32_t i = 0; i < len; ++i)
			par_buf[i] += (r_par_buf[i] - par_buf[i]) * mod_buf[i];
	} else {
		line_skip(p.r_par, len);
	}
	for (uint32_t i = 0; i < p.mods->count; ++i)
		run_block(g, bufs, len, &g->ops[p.mods->ids[i]], freq, false, true);
}

bool run_selfmod_param(Gen *g, Buf *bufs, uint32_t len, Op *n, float *freq) {  /* generator.c:479-498 */
	bool filled = false;
	if (n->pm_a.v0 != 0.f || (n->pm_a.flags & SAUABI_LINEP_GOAL)) {
		line_run(&n->pm_a, bufs[0], len, NULL);
		filled = true;
	} else {
		line_skip(&n->pm_a, len);
	}
	for (uint32_t i = 0; i < n->apmods->count; ++i) {
		run_block(g, bufs, len, &g->ops[n->apmods->ids[i]], freq, false, filled);
		filled = true;
	}
	return filled;
}

/* ---- noise, noise.h:41-185 ---- */

void noise_run(Op *o, float *buf, uint32_t len) {
	const float scale = 1.f / 2147483648.f;
	switch (o->ntype) {
	default:
	case SAUABI_NOISE_wh:
		for (uint32_t i = 0; i < len; ++i) buf[i] = sau::fscalei(sau::ranfast32(o->n++), scale);
		break;
	case SAUABI_NOISE_gw:
		for (uint32_t i = 0; i < len; ++i) buf[i] = sau::franssgauss32(o->n++);
		break;
	case SAUABI_NOISE_bw:
		for (uint32_t i = 0; i < len; ++i) {
			uint32_t n = o->n++;
			int32_t s = sau::sar32((int32_t) sau::ranfast32(n), 31) * 2 + 1;
			buf[i] = (float) s;
		}
		break;
	case SAUABI_NOISE_tw:
		for (uint32_t i = 0; i < len; ++i) {
			uint32_t n = o->n++;
			int32_t s = sau::sar32((int32_t) sau::ranfast32(n), 31) * 2 + 1;
			buf[i] = (n & 1) ? (float) s : 0.f;
		}
		break;
	case SAUABI_NOISE_re: {
		uint32_t sum = o->nprev;
		for (uint32_t i = 0; i < len; ++i) {
			int32_t s = (int32_t) sau::ranfast32(o->n++);
			sum += (uint32_t) (s >> 6);
			s = sau::foldhd32((int32_t) sum);
			buf[i] = sau::fscalei((uint32_t) s, scale);
		}
		o->nprev = sum;
		break; }
	case SAUABI_NOISE_vi: {
		uint32_t s0 = o->nprev;
		for (uint32_t i = 0; i < len; ++i) {
			uint32_t s1 = sau::ranfast32(o->n++);
			buf[i] = sau::fscalei((s1 / 2) - (s0 / 2), scale);
			s0 = s1;
		}
		o->nprev = s0;
		break; }
	case SAUABI_NOISE_bv: {
		int32_t s0 = (int32_t) o->nprev;
		for (uint32_t i = 0; i < len; ++i) {
			uint32_t n = o->n++;
			int32_t s1 = sau::sar32((int32_t) sau::ranfast32(n), 31);
			s1 = (n & 1) ? (s1 * 2 + 1) : 0;
			buf[i] = (float) (s1 - s0);
			s0 = s1;
		}
		o->nprev = (uint32_t) s0;
		break; }
	}
}

/* ---- wave oscillator, wosc.h:135-310 ---- */

void phasor_fill(Op *o, uint32_t *phase_buf, uint32_t len, const float *freq,
		const float *pm, const float *fpm) {
	const float ps = 2147483648.f;
	for (uint32_t i = 0; i < len; ++i) {
		float f = freq[i];
		int64_t ofs = 0;
		if (pm && fpm) ofs = sau::pofs_pm_fpm(pm[i], fpm[i], f, ps);
		else if (pm) ofs = sau::pofs_pm(pm[i], ps);
		else if (fpm) ofs = sau::pofs_fpm(fpm[i], f, ps);
		o->phase += (uint32_t) sau::ftoi64(o->coeff * f);         /* pre-increment, wosc.h:129 */
		phase_buf[i] = (uint32_t) ofs + o->phase;
	}
}

void wosc_reset(Gen *g, Op *o, uint32_t phase) {                 /* wosc.h:215-230 */
	if (o->oscflags & OSC_RESET_DIFF) {
		const float *lut = g->tab.pilut[o->wave];
		double poly, c0;
		sau::herp(lut, phase - sau::WAVE_SLEN, &poly, &c0);
		double Is = sau::herp(lut, phase, (double*) 0, (double*) 0);
		/* as compiled: (Is - poly_prev) - c0_prev, x = amp_scale*256 exactly */
		float x = g->tab.amp_scale[o->wave] * 256.f;
		o->prev_s = (float) (((Is - poly) - c0) * (double) x + (double) g->tab.amp_dc[o->wave]);
		o->prev_Is = Is;
		o->prev_phase = phase;
	}
	o->oscflags &= ~OSC_RESET_DIFF;
}

void wosc_run(Gen *g, Op *o, float *buf, uint32_t len, const uint32_t *phase_buf,
		const float *pm_abuf) {
	const float *lut = g->tab.pilut[o->wave];
	const float diff_scale = sau::wave_dvscale(g->tab.amp_scale[o->wave]);
	const float diff_offset = g->tab.amp_dc[o->wave];
	if (len > 0 && (o->oscflags & OSC_RESET_DIFF)) wosc_reset(g, o, phase_buf[0]);
	for (uint32_t i = 0; i < len; ++i) {
		float s;
		uint32_t phase = phase_buf[i];
		if (pm_abuf)
			phase += (uint32_t) sau::ftoi64(o->fb_s * pm_abuf[i] * 2147483648.f);
		int32_t phase_diff = (int32_t) (phase - o->prev_phase);
		if (phase_diff == 0) {
			s = o->prev_s;
		} else {
			double Is = sau::herp(lut, phase, (double*) 0, (double*) 0);
			s = sau::wosc_diff(Is, o->prev_Is, phase_diff, diff_scale, diff_offset);
			o->prev_Is = Is;
			o->prev_s = s;
			o->prev_phase = phase;
		}
		buf[i] = s;
		if (pm_abuf) o->fb_s = (o->fb_s + s) * 0.5f;
	}
}

/* ---- random segments oscillator, rasg.h:165-222,692-772 ---- */

void cyclor_fill(Op *o, uint32_t *cycle_buf, float *phase_f, uint32_t len,
		const float *freq, const float *pm, const float *fpm) {
	float coeff = o->coeff;
	float ps = 2147483648.f;
	if (o->rate2x) { coeff *= 2; ps *= 2; }
	for (uint32_t i = 0; i < len; ++i) {
		float f = freq[i];
		int64_t ofs = 0;
		if (pm && fpm) ofs = sau::pofs_pm_fpm(pm[i], fpm[i], f, ps);
		else if (pm) ofs = sau::pofs_pm(pm[i], ps);
		else if (fpm) ofs = sau::pofs_fpm(fpm[i], f, ps);
		uint64_t cp = (uint64_t) ofs + o->cycle_phase;             /* post-increment, rasg.h:154-155 */
		o->cycle_phase += (uint64_t) sau::ftoi64(coeff * f);
		cycle_buf[i] = (uint32_t) (cp >> 32);
		uint32_t phase = ((uint32_t) cp) >> 1;
		phase_f[i] = sau::i2f((int32_t) phase) * (1.f / 2147483648.f);
	}
}

void rasg_run(Op *o, uint32_t len, float *main_buf, const uint32_t *cycle_buf,
		const float *pm_abuf) {
	const unsigned flags = o->ras_flags, func = o->ras_func;
	const int sr = o->ras_level, line = o->ras_line;
	if (!pm_abuf) {
		for (uint32_t i = 0; i < len; ++i) {
			/* sauLine_map_cub: 4-wide body, scalar tail (line.c:16-24 as compiled) */
			bool tail = i >= (len & ~3u);
			main_buf[i] = sau::rasg_sample(func, flags, sr, o->ras_alpha, line,
					cycle_buf[i], main_buf[i], false, tail);
		}
	} else {                                                      /* rasg.h:242-280 */
		for (uint32_t i = 0; i < len; ++i) {
			float pm_a = o->fb_s * pm_abuf[i] * 0.5f;
			float phase = main_buf[i] + pm_a;
			int32_t cycle_adj = (int32_t) floorf(phase);
			uint32_t cycle = cycle_buf[i] + (uint32_t) cycle_adj;
			phase -= (float) cycle_adj;
			float s = sau::rasg_sample(func, flags, sr, o->ras_alpha, line,
					cycle, phase, true, false);
			main_buf[i] = s;
			o->fb_s = ((o->fb_s + o->prev_s) + s) * 0.5f;         /* as compiled, B.3 */
			o->prev_s = s;
		}
	}
}

/* ---- operator walk, generator.c:505-729 ---- */

void run_block_gen(Gen *g, Buf *bufs, uint32_t len, Op *n, bool wave_env, bool layer) {
	float *mix_buf = *(bufs++);
	Par p = { &n->amp, &n->amp2, n->amods, n->ramods };
	run_param(g, bufs, len, p, NULL, NULL, false);
	float *amp = *(bufs++);
	float *tmp = *bufs;
	if (n->type == SAUABI_POPT_noise) noise_run(n, tmp, len);
	else for (uint32_t i = 0; i < len; ++i) tmp[i] = 1.f;
	block_mix(mix_buf, len, wave_env, layer, tmp, amp);
}

void run_block_wosc(Gen *g, Buf *bufs, uint32_t len, Op *n, float *parent_freq,
		bool wave_env, bool layer) {
	float *mix_buf = *(bufs++), *pm_buf = NULL, *fpm_buf = NULL;
	uint32_t *phase_buf = (uint32_t*) *(bufs++);
	Par pf = { &n->freq, &n->freq2, n->fmods, n->rfmods };
	run_param(g, bufs, len, pf, parent_freq, NULL, true);
	float *freq = *(bufs++);
	if (n->pmods->count > 0) {
		for (uint32_t i = 0; i < n->pmods->count; ++i)
			run_block(g, bufs + 0, len, &g->ops[n->pmods->ids[i]], freq, false, i);
		pm_buf = bufs[0];
	}
	if (n->fpmods->count > 0) {
		for (uint32_t i = 0; i < n->fpmods->count; ++i)
			run_block(g, bufs + 1, len, &g->ops[n->fpmods->ids[i]], freq, false, i);
		fpm_buf = bufs[1];
	}
	phasor_fill(n, phase_buf, len, freq, pm_buf, fpm_buf);
	Par pa = { &n->amp, &n->amp2, n->amods, n->ramods };
	run_param(g, bufs, len, pa, NULL, freq, false);
	float *amp = *(bufs++);
	float *tmp = *(bufs++);
	if (run_selfmod_param(g, bufs, len, n, freq))
		wosc_run(g, n, tmp, len, phase_buf, *bufs);
	else
		wosc_run(g, n, tmp, len, phase_buf, NULL);
	block_mix(mix_buf, len, wave_env, layer, tmp, amp);
}

void run_block_rasg(Gen *g, Buf *bufs, uint32_t len, Op *n, float *parent_freq,
		bool wave_env, bool layer) {
	float *mix_buf = *(bufs++), *pm_buf = NULL, *fpm_buf = NULL;
	uint32_t *cycle_buf = (uint32_t*) *(bufs++);
	float *rasg_buf = *(bufs++);
	Par pf = { &n->freq, &n->freq2, n->fmods, n->rfmods };
	run_param(g, bufs, len, pf, parent_freq, NULL, true);
	float *freq = *(bufs++);
	if (n->pmods->count > 0) {
		for (uint32_t i = 0; i < n->pmods->count; ++i)
			run_block(g, bufs + 0, len, &g->ops[n->pmods->ids[i]], freq, false, i);
		pm_buf = bufs[0];
	}
	if (n->fpmods->count > 0) {
		for (uint32_t i = 0; i < n->fpmods->count; ++i)
			run_block(g, bufs + 1, len, &g->ops[n->fpmods->ids[i]], freq, false, i);
		fpm_buf = bufs[1];
	}
	cyclor_fill(n, cycle_buf, rasg_buf, len, freq, pm_buf, fpm_buf);
	Par pa = { &n->amp, &n->amp2, n->amods, n->ramods };
	run_param(g, bufs, len, pa, NULL, freq, false);
	float *amp = *(bufs++);
	if (run_selfmod_param(g, bufs, len, n, freq))
		rasg_run(n, len, rasg_buf, cycle_buf, *bufs);
	else
		rasg_run(n, len, rasg_buf, cycle_buf, NULL);
	block_mix(mix_buf, len, wave_env, layer, rasg_buf, amp);
}

uint32_t run_block(Gen *g, Buf *bufs, uint32_t buf_len, Op *n, float *parent_freq,
		bool wave_env, bool layer) {                          /* generator.c:675-729 */
	float *mix_buf = *bufs;
	if (n->flags & ON_VISITED) {
		for (uint32_t i = 0; i < buf_len; ++i) mix_buf[i] = 0;
		return buf_len;
	}
	n->flags |= ON_VISITED;
	uint32_t len = buf_len, skip_len = 0;
	if (n->time < len && !(n->flags & ON_TIME_INF)) {
		skip_len = len - n->time;
		len = n->time;
	}
	switch (n->type) {
	case SAUABI_POPT_amp:
	case SAUABI_POPT_noise: run_block_gen(g, bufs, len, n, wave_env, layer); break;
	case SAUABI_POPT_wave: run_block_wosc(g, bufs, len, n, parent_freq, wave_env, layer); break;
	case SAUABI_POPT_raseg: run_block_rasg(g, bufs, len, n, parent_freq, wave_env, layer); break;
	}
	if (!(n->flags & ON_TIME_INF)) {
		if (!layer && skip_len > 0)
			for (uint32_t i = 0; i < skip_len; ++i) mix_buf[len + i] = 0;
		n->time -= len;
	}
	n->flags &= ~ON_VISITED;
	return len;
}

/* ---- voices and output, generator.c:734-878 ---- */

void mix_add(Gen *g, Op *n, Voice *vn, uint32_t len) {
	float *s_buf = g->gen_bufs[0];
	float *pan_buf = NULL;
	if ((n->pan.flags & SAUABI_LINEP_GOAL) || n->camods->count > 0) {
		pan_buf = g->gen_bufs[1 + vn->freq_buf_id];
		line_run(&n->pan, pan_buf, len, NULL);
	} else {
		line_skip(&n->pan, len);
	}
	if (n->camods->count > 0) {
		float *freq_buf = vn->freq_buf_id > 0 ? g->gen_bufs[vn->freq_buf_id] : NULL;
		for (uint32_t i = 0; i < n->camods->count; ++i)
			run_block(g, g->gen_bufs + 1 + vn->freq_buf_id, len,
					&g->ops[n->camods->ids[i]], freq_buf, false, true);
	}
	for (uint32_t i = 0; i < len; ++i) {
		float s = s_buf[i] * g->amp_scale;
		float s_r = s * (pan_buf ? pan_buf[i] : n->pan.v0);
		/* as compiled (Appendix B.3): (L + s) - r, (R + s) + r */
		g->mix_l[i] = (g->mix_l[i] + s) - s_r;
		g->mix_r[i] = (g->mix_r[i] + s) + s_r;
	}
	if (g->mix_add_max < len) g->mix_add_max = len;
}

uint32_t run_for_time(Gen *g, uint32_t time, int16_t *buf, bool stereo) {
	int16_t *sp = buf;
	uint32_t gen_len = 0;
	while (time > 0) {
		uint32_t len = time < BUF_LEN ? time : BUF_LEN;
		time -= len;
		if (g->mix_add_max) {
			memset(g->mix_l, 0, sizeof(float) * g->mix_add_max);
			memset(g->mix_r, 0, sizeof(float) * g->mix_add_max);
			g->mix_add_max = 0;
		}
		uint32_t last_len = 0;
		for (uint32_t i = g->voice; i < g->voices.size(); ++i) {
			Voice *vn = &g->voices[i];
			if (vn->duration == 0) continue;
			Op *n = &g->ops[vn->carr_op_id];                   /* run_voice, generator.c:833-846 */
			uint32_t vtime = vn->duration, out_len = 0;
			if (vtime > len) vtime = len;
			if (n->time > 0)
				out_len = run_block(g, g->gen_bufs, vtime, n, NULL, false, false);
			if (out_len > 0) mix_add(g, n, vn, out_len);
			vn->duration -= vtime;
			if (out_len > last_len) last_len = out_len;
		}
		if (last_len > 0) {
			gen_len += last_len;
			g->out_clear = false;
			for (uint32_t i = 0; i < last_len; ++i) {            /* generator.c:795-825 */
				if (stereo) {
					float l = sau::fclampf(g->mix_l[i], -1.f, 1.f);
					float r = sau::fclampf(g->mix_r[i], -1.f, 1.f);
					*sp++ += (int16_t) lrintf(l * 32767.f);
					*sp++ += (int16_t) lrintf(r * 32767.f);
				} else {
					float m = (g->mix_l[i] + g->mix_r[i]) * 0.5f;
					m = sau::fclampf(m, -1.f, 1.f);
					*sp++ += (int16_t) lrintf(m * 32767.f);
				}
			}
		}
	}
	return gen_len;
}

} // namespace

extern "C" {

typedef struct OracleTables {
	const float *pilut[SAUABI_WAVE_NAMED];
	float amp_scale[SAUABI_WAVE_NAMED];
	float amp_dc[SAUABI_WAVE_NAMED];
	int32_t phase_adj[SAUABI_WAVE_NAMED];
} OracleTables;   /* same layout as saugen_WaveTables */

/* sau_create_Generator, generator.c:172-217 */
void *oracle_create(const sauabi_Program *prg, uint32_t srate, const OracleTables *t) {
	Gen *g = new Gen();
	g->prg = prg;
	g->srate = srate;
	for (int w = 0; w < SAUABI_WAVE_NAMED; ++w) {
		memcpy(g->tab.pilut[w], t->pilut[w], sizeof(float) * SAUABI_WAVE_LEN);
		g->tab.amp_scale[w] = t->amp_scale[w];
		g->tab.amp_dc[w] = t->amp_dc[w];
		g->tab.phase_adj[w] = t->phase_adj[w];
	}
	g->ops.resize(prg->op_count);
	if (prg->op_count) memset((void*) g->ops.data(), 0, sizeof(Op) * prg->op_count);
	g->voices.assign(prg->vo_count, Voice{0, 0, 0, 0});
	size_t nb = (size_t) (1 + prg->op_nest_depth) * 7;          /* generator.c:133 */
	g->gen_bufs_store.assign(nb * BUF_LEN, 0.f);
	g->gen_bufs = (Buf*) g->gen_bufs_store.data();
	memset(g->mix_l, 0, sizeof(Buf)); memset(g->mix_r, 0, sizeof(Buf));
	g->amp_scale = 0.5f * prg->ampmult;
	if (prg->mode & SAUABI_PMODE_AMP_DIV_VOICES) g->amp_scale /= (float) prg->vo_count;
	int carry = 0;
	g->ev_wait.resize(prg->ev_count);
	for (size_t i = 0; i < prg->ev_count; ++i)
		g->ev_wait[i] = ms_in_samples(prg->events[i].wait_ms, srate, &carry);
	g->event = 0; g->event_pos = 0; g->voice = 0;
	g->out_clear = false; g->mix_add_max = 0;
	return g;
}
void oracle_destroy(void *p) { delete (Gen*) p; }

/* sauGenerator_run, generator.c:905-973 */
int oracle_run(void *p, int16_t *buf, size_t buf_len, int stereo, size_t *out_len) {
	Gen *g = (Gen*) p;
	int16_t *sp = buf;
	uint32_t len = (uint32_t) buf_len, skip_len, last_len, gen_len = 0;
	if (!g->out_clear) {
		g->out_clear = true;
		memset(buf, 0, sizeof(int16_t) * (stereo ? len * 2 : len));
	}
	for (;;) {
		skip_len = 0;
		while (g->event < g->ev_wait.size()) {
			uint32_t wait = g->ev_wait[g->event];
			if (g->event_pos < wait) {
				uint32_t waittime = wait - g->event_pos;
				if (waittime < len) { skip_len = len - waittime; len = waittime; }
				g->event_pos += len;
				break;
			}
			handle_event(g, &g->prg->events[g->event]);
			++g->event;
			g->event_pos = 0;
		}
		last_len = run_for_time(g, len, sp, stereo != 0);
		if (skip_len > 0) {
			gen_len += len;
			sp += stereo ? len * 2 : len;
			len = skip_len;
		} else {
			gen_len += last_len;
			break;
		}
	}
	for (;;) {
		if (g->voice == g->voices.size()) {
			if (g->event != g->ev_wait.size()) break;
			if (out_len) *out_len = gen_len;
			return 0;
		}
		if (g->voices[g->voice].duration != 0) break;
		++g->voice;
	}
	if (out_len) *out_len = buf_len;
	return 1;
}

/* State view with the same layout as oracle/ref_harness.c:RefOpState. */
typedef struct OLineView { float v0, vt; uint32_t pos, end, type, flags; } OLineView;
typedef struct OOpView {
	uint32_t inited, type, flags, time;
	OLineView amp, amp2, pan, freq, freq2, pm_a;
	uint32_t i0, i1, mode, oscflags;
	double prev_Is; float prev_s, fb_s; uint32_t alpha, rate2x;
} OOpView;
static void view_line(OLineView *d, const Line *s) {
	d->v0 = s->v0; d->vt = s->vt; d->pos = s->pos; d->end = s->end;
	d->type = s->type; d->flags = s->flags;
}
int oracle_op_state(void *p, uint32_t op_id, OOpView *out) {
	Gen *g = (Gen*) p;
	if (op_id >= g->ops.size()) return -1;
	const Op *n = &g->ops[op_id];
	memset(out, 0, sizeof(*out));
	out->inited = (n->flags & ON_INIT) != 0;
	if (!out->inited) return 0;
	out->type = n->type; out->flags = n->flags; out->time = n->time;
	view_line(&out->amp, &n->amp); view_line(&out->amp2, &n->amp2);
	view_line(&out->pan, &n->pan);
	if (n->type >= SAUABI_POPT_wave) {
		view_line(&out->freq, &n->freq); view_line(&out->freq2, &n->freq2);
		view_line(&out->pm_a, &n->pm_a);
	}
	switch (n->type) {
	case SAUABI_POPT_noise: out->i0 = n->n; out->i1 = n->nprev; out->mode = n->ntype; break;
	case SAUABI_POPT_wave:
		out->i0 = n->phase; out->i1 = n->prev_phase; out->mode = n->wave;
		out->oscflags = n->oscflags; out->prev_Is = n->prev_Is;
		out->prev_s = n->prev_s; out->fb_s = n->fb_s; break;
	case SAUABI_POPT_raseg:
		out->i0 = (uint32_t) n->cycle_phase; out->i1 = (uint32_t) (n->cycle_phase >> 32);
		out->mode = n->ras_line;
		out->oscflags = n->ras_flags | (n->ras_func << 16) | (n->ras_level << 24);
		out->prev_s = n->prev_s; out->fb_s = n->fb_s;
		out->alpha = n->ras_alpha; out->rate2x = n->rate2x; break;
	}
	return 0;
}
int oracle_voice_state(void *p, uint32_t vo_id, uint32_t *out) {
	Gen *g = (Gen*) p;
	if (vo_id >= g->voices.size()) return -1;
	out[0] = g->voices[vo_id].duration; out[1] = g->voices[vo_id].flags;
	out[2] = g->voices[vo_id].carr_op_id; out[3] = g->voices[vo_id].freq_buf_id;
	return 0;
}
const float *oracle_gen_buf(void *p, uint32_t k) { return ((Gen*) p)->gen_bufs[k]; }
const float *oracle_mix_buf(void *p, uint32_t ch) { Gen *g = (Gen*) p; return ch ? g->mix_r : g->mix_l; }

} // extern "C"
