/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the
 * product). White-box access to the UNMODIFIED reference generator, compiled
 * by oracle/Makefile into oracle/_ref/libsauref.so.
 *
 * The reference generator that actually renders is the separately compiled,
 * as-shipped-flags object (pic_generator.o). This TU includes
 * sau/generator.c from the read-only reference tree ONLY to obtain its
 * private struct definitions (sauGenerator, OperatorNode, VoiceNode:
 * sau/generator.c:45-130); the function copies that come along are renamed
 * through the macros below and never called.
 */
#define sau_create_Generator  wbunused_create_Generator
#define sau_destroy_Generator wbunused_destroy_Generator
#define sauGenerator_run      wbunused_Generator_run
#define sauNoise_names        wbunused_Noise_names
#include "sau/generator.c"
#undef sau_create_Generator
#undef sau_destroy_Generator
#undef sauGenerator_run
#undef sauNoise_names
#include "sau/script.h"

/* the real ones, from pic_generator.o */
sauGenerator* sau_create_Generator(const sauProgram *restrict prg, uint32_t srate);
void sau_destroy_Generator(sauGenerator *restrict o);
bool sauGenerator_run(sauGenerator *restrict o, int16_t *restrict buf,
		size_t buf_len, bool stereo, size_t *restrict out_len);

/* ---- front end --------------------------------------------------------- */

/* Parse + build a program from script text (is_path=0) or a file path.
 * Deterministic mode (-d, saugns.c:430) is always on. */
const sauProgram *refwb_build_program(const char *str, int is_path) {
	sauScriptArg arg = {0};
	arg.str = str;
	arg.is_path = is_path != 0;
	arg.no_time = true;
	return sau_build_Program(&arg);
}

void refwb_discard_program(sauProgram *prg) { sau_discard_Program(prg); }

void refwb_program_info(const sauProgram *prg, uint32_t *out) {
	out[0] = (uint32_t) prg->ev_count;
	out[1] = prg->vo_count;
	out[2] = prg->op_count;
	out[3] = prg->op_nest_depth;
	out[4] = prg->duration_ms;
	out[5] = prg->mode;
}

/* ---- generator pass-through -------------------------------------------- */

sauGenerator *refwb_create(const sauProgram *prg, uint32_t srate) {
	return sau_create_Generator(prg, srate);
}
void refwb_destroy(sauGenerator *g) { sau_destroy_Generator(g); }
int refwb_run(sauGenerator *g, int16_t *buf, size_t buf_len, int stereo,
		size_t *out_len) {
	return sauGenerator_run(g, buf, buf_len, stereo != 0, out_len);
}

/* Render a whole program like Player_run (saugns.c:575-623) does, with the
 * given call size, appending every call's out_len frames to `out`.
 * Returns frames written (stops early if out_cap frames would be exceeded). */
size_t refwb_render(const sauProgram *prg, uint32_t srate, int stereo,
		size_t call_len, int16_t *out, size_t out_cap) {
	sauGenerator *g = sau_create_Generator(prg, srate);
	if (!g) return 0;
	size_t ch = stereo ? 2 : 1, total = 0;
	int16_t *buf = calloc(call_len * ch, sizeof(int16_t));
	bool run = true;
	while (run) {
		size_t len = 0;
		run = sauGenerator_run(g, buf, call_len, stereo != 0, &len);
		if (total + len > out_cap) len = out_cap - total;
		memcpy(out + total * ch, buf, len * ch * sizeof(int16_t));
		total += len;
		if (total >= out_cap) break;
	}
	free(buf);
	sau_destroy_Generator(g);
	return total;
}

/* Same, but discards the audio: used for CPU-baseline timing. */
size_t refwb_render_null(const sauProgram *prg, uint32_t srate, int stereo,
		size_t call_len, size_t max_frames) {
	sauGenerator *g = sau_create_Generator(prg, srate);
	if (!g) return 0;
	size_t ch = stereo ? 2 : 1, total = 0;
	int16_t *buf = calloc(call_len * ch, sizeof(int16_t));
	bool run = true;
	while (run) {
		size_t len = 0;
		run = sauGenerator_run(g, buf, call_len, stereo != 0, &len);
		total += len;
		if (max_frames && total >= max_frames) break;
	}
	free(buf);
	sau_destroy_Generator(g);
	return total;
}

/* ---- white-box state --------------------------------------------------- */

typedef struct RefLineState {
	float v0, vt;
	uint32_t pos, end;
	uint32_t type, flags;
} RefLineState;

typedef struct RefOpState {
	uint32_t inited;      /* ON_INIT set */
	uint32_t type;        /* SAU_POPT_N_* */
	uint32_t flags;
	uint32_t time;
	RefLineState amp, amp2, pan, freq, freq2, pm_a;
	/* W: phase, prev_phase; R: lo/hi of cycle_phase; N: n, prev */
	uint32_t i0, i1;
	uint32_t mode;        /* wave / noise type / ras line */
	uint32_t oscflags;    /* W: SAU_OSC_* flags; R: opt.flags | func<<16 | level<<24 */
	double prev_Is;
	float prev_s, fb_s;
	uint32_t alpha, rate2x;
} RefOpState;

static void get_line(RefLineState *d, const sauLine *s) {
	d->v0 = s->v0; d->vt = s->vt; d->pos = s->pos; d->end = s->end;
	d->type = s->type; d->flags = s->flags;
}

int refwb_op_state(const sauGenerator *g, uint32_t op_id, RefOpState *out) {
	if (op_id >= g->op_count) return -1;
	const OperatorNode *n = &g->operators[op_id];
	memset(out, 0, sizeof(*out));
	out->inited = (n->gen.flags & ON_INIT) != 0;
	if (!out->inited) return 0;
	out->type = n->gen.type;
	out->flags = n->gen.flags;
	out->time = n->gen.time;
	get_line(&out->amp, &n->gen.amp.par);
	get_line(&out->amp2, &n->gen.amp.r_par);
	get_line(&out->pan, &n->gen.pan);
	if (n->gen.type >= SAU_POPT_N_wave) {
		get_line(&out->freq, &n->osc.freq.par);
		get_line(&out->freq2, &n->osc.freq.r_par);
		get_line(&out->pm_a, &n->osc.pm_a);
	}
	switch (n->gen.type) {
	case SAU_POPT_N_noise:
		out->i0 = n->ng.noiseg.n;
		out->i1 = n->ng.noiseg.prev;
		out->mode = n->ng.noiseg.type;
		break;
	case SAU_POPT_N_wave:
		out->i0 = n->wo.wosc.phasor.phase;
		out->i1 = n->wo.wosc.prev_phase;
		out->mode = n->wo.wosc.wave;
		out->oscflags = n->wo.wosc.flags;
		out->prev_Is = n->wo.wosc.prev_Is;
		out->prev_s = n->wo.wosc.prev_s;
		out->fb_s = n->wo.wosc.fb_s;
		break;
	case SAU_POPT_N_raseg:
		out->i0 = (uint32_t) n->rg.rasg.cyclor.cycle_phase;
		out->i1 = (uint32_t) (n->rg.rasg.cyclor.cycle_phase >> 32);
		out->mode = n->rg.rasg.opt.line;
		out->oscflags = n->rg.rasg.opt.flags |
			(n->rg.rasg.opt.func << 16) | (n->rg.rasg.opt.level << 24);
		out->prev_s = n->rg.rasg.prev_s;
		out->fb_s = n->rg.rasg.fb_s;
		out->alpha = n->rg.rasg.opt.alpha;
		out->rate2x = n->rg.rasg.cyclor.rate2x;
		break;
	}
	return 0;
}

int refwb_voice_state(const sauGenerator *g, uint32_t vo_id, uint32_t *out) {
	if (vo_id >= g->vo_count) return -1;
	const VoiceNode *vn = &g->voices[vo_id];
	out[0] = vn->duration;
	out[1] = vn->flags;
	out[2] = vn->carr_op_id;
	out[3] = vn->freq_buf_id;
	return 0;
}

void refwb_gen_state(const sauGenerator *g, uint32_t *out) {
	out[0] = (uint32_t) g->event;
	out[1] = g->event_pos;
	out[2] = g->voice;
	out[3] = g->vo_count;
	out[4] = g->op_count;
	out[5] = (uint32_t) g->ev_count;
}
float refwb_amp_scale(const sauGenerator *g) { return g->amp_scale; }

/* Work buffers of the last rendered block (sau/generator.c:120,157-162). */
const float *refwb_gen_buf(const sauGenerator *g, uint32_t k) { return g->gen_bufs[k]; }
const float *refwb_mix_buf(const sauGenerator *g, uint32_t ch) { return g->mix_bufs[ch]; }

/* ---- tables (sau/wave.c:49-66) ------------------------------------------ */

const float *refwb_pilut(uint32_t wave) {
	sau_global_init_Wave();
	return wave < SAU_WAVE_NAMED ? sauWave_piluts[wave] : NULL;
}
const float *refwb_lut(uint32_t wave) {
	sau_global_init_Wave();
	return wave < SAU_WAVE_NAMED ? sauWave_luts[wave] : NULL;
}
void refwb_picoeffs(uint32_t wave, float *amp_scale, float *amp_dc, int32_t *phase_adj) {
	*amp_scale = sauWave_picoeffs[wave].amp_scale;
	*amp_dc = sauWave_picoeffs[wave].amp_dc;
	*phase_adj = sauWave_picoeffs[wave].phase_adj;
}

/* ---- function-level entry points for fuzzing the restatement ------------ */

void refwb_line_fill(uint32_t type, float *buf, uint32_t len, float v0, float vt,
		uint32_t pos, uint32_t time, const float *mulbuf) {
	sauLine_fill_funcs[type](buf, len, v0, vt, pos, time, mulbuf);
}
void refwb_line_map(uint32_t type, float *buf, uint32_t len,
		const float *end0, const float *end1) {
	sauLine_map_funcs[type](buf, len, end0, end1);
}
float refwb_line_val(uint32_t type, float x, float a, float b) {
	return sauLine_val_funcs[type](x, a, b);
}

/* ---- ABI layout of the boundary's input data model ---------------------- */
/* Order must match saugen_abi_layout() in the product (include/sau_program_abi.h). */
size_t refwb_abi_layout(uint32_t *out, size_t cap) {
	const uint32_t v[] = {
		sizeof(sauLine), offsetof(sauLine, v0), offsetof(sauLine, vt),
		offsetof(sauLine, pos), offsetof(sauLine, end),
		offsetof(sauLine, time_ms), offsetof(sauLine, type),
		offsetof(sauLine, flags),
		sizeof(sauTime), offsetof(sauTime, v_ms), offsetof(sauTime, flags),
		sizeof(sauRasOpt), offsetof(sauRasOpt, line), offsetof(sauRasOpt, alpha),
		sizeof(sauProgramIDArr), offsetof(sauProgramIDArr, ids),
		sizeof(sauProgramOpData),
		offsetof(sauProgramOpData, id), offsetof(sauProgramOpData, params),
		offsetof(sauProgramOpData, time), offsetof(sauProgramOpData, pan),
		offsetof(sauProgramOpData, amp), offsetof(sauProgramOpData, amp2),
		offsetof(sauProgramOpData, freq), offsetof(sauProgramOpData, freq2),
		offsetof(sauProgramOpData, pm_a), offsetof(sauProgramOpData, phase),
		offsetof(sauProgramOpData, seed), offsetof(sauProgramOpData, use_type),
		offsetof(sauProgramOpData, type), offsetof(sauProgramOpData, mode),
		offsetof(sauProgramOpData, camods), offsetof(sauProgramOpData, amods),
		offsetof(sauProgramOpData, ramods), offsetof(sauProgramOpData, fmods),
		offsetof(sauProgramOpData, rfmods), offsetof(sauProgramOpData, pmods),
		offsetof(sauProgramOpData, apmods), offsetof(sauProgramOpData, fpmods),
		sizeof(sauProgramEvent),
		offsetof(sauProgramEvent, wait_ms), offsetof(sauProgramEvent, vo_id),
		offsetof(sauProgramEvent, carr_op_id), offsetof(sauProgramEvent, op_count),
		offsetof(sauProgramEvent, op_data_count), offsetof(sauProgramEvent, op_list),
		offsetof(sauProgramEvent, op_data),
		sizeof(sauProgram),
		offsetof(sauProgram, events), offsetof(sauProgram, ev_count),
		offsetof(sauProgram, mode), offsetof(sauProgram, vo_count),
		offsetof(sauProgram, op_count), offsetof(sauProgram, op_nest_depth),
		offsetof(sauProgram, duration_ms), offsetof(sauProgram, ampmult),
		offsetof(sauProgram, name),
		SAU_WAVE_NAMED, SAU_LINE_NAMED, SAU_NOISE_NAMED, SAU_RAS_FUNCTIONS,
	};
	size_t n = sizeof(v) / sizeof(v[0]);
	for (size_t i = 0; i < n && i < cap; ++i) out[i] = v[i];
	return n;
}
/* RasOpt bitfields cannot be offsetof'd: decode one for the layout test. */
void refwb_rasopt_decode(const sauRasOpt *o, uint32_t *out) {
	out[0] = o->line; out[1] = o->flags; out[2] = o->func;
	out[3] = o->level; out[4] = o->alpha;
}
