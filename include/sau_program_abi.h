/* include/sau_program_abi.h
 *
 * Binary layout of the data model the saugns front end (scanner + parser +
 * parseconv, kept as the reference's own host C code) hands to the generator
 * back end.  This is the INPUT FORMAT of the drop-in boundary: a
 * `const sauProgram*` (reference sau/program.h:253-265) arrives at
 * sau_create_Generator (reference sau/generator.h:20-21) and is only read.
 *
 * Nothing here is code: it is a field-for-field declaration of the structs in
 * the reference headers, written against the x86-64 SysV ABI, so that the
 * back end can be compiled without the reference tree present.  Each struct
 * cites the declaration it mirrors; tests/test_abi.py checks every size and
 * offset against the real headers through oracle/_ref (refwb_abi_layout).
 */
#ifndef SAU_PROGRAM_ABI_H
#define SAU_PROGRAM_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sau/line.h:99-107 */
enum {
	SAUABI_LINEP_STATE       = 1<<0,
	SAUABI_LINEP_STATE_RATIO = 1<<1,
	SAUABI_LINEP_GOAL        = 1<<2,
	SAUABI_LINEP_GOAL_RATIO  = 1<<3,
	SAUABI_LINEP_TYPE        = 1<<4,
	SAUABI_LINEP_TIME        = 1<<5,
	SAUABI_LINEP_TIME_IF_NEW = 1<<6,
};

/* sau/line.h:18-32 (enum order = SAU_LINE__ITEMS order) */
enum {
	SAUABI_LINE_cos = 0, SAUABI_LINE_lin, SAUABI_LINE_sah, SAUABI_LINE_exp,
	SAUABI_LINE_log, SAUABI_LINE_xpe, SAUABI_LINE_lge, SAUABI_LINE_sqe,
	SAUABI_LINE_cub, SAUABI_LINE_smo, SAUABI_LINE_ncl, SAUABI_LINE_nhl,
	SAUABI_LINE_uwh, SAUABI_LINE_NAMED
};

/* sau/wave.h:33-81 */
enum {
	SAUABI_WAVE_sin = 0, SAUABI_WAVE_tri, SAUABI_WAVE_srs, SAUABI_WAVE_sqr,
	SAUABI_WAVE_ean, SAUABI_WAVE_cat, SAUABI_WAVE_eto, SAUABI_WAVE_par,
	SAUABI_WAVE_mto, SAUABI_WAVE_saw, SAUABI_WAVE_hsi, SAUABI_WAVE_spa,
	SAUABI_WAVE_NAMED
};
#define SAUABI_WAVE_LEN 2048 /* sau/wave.h:18-19 */

/* sau/program.h:102-120 */
enum {
	SAUABI_NOISE_wh = 0, SAUABI_NOISE_gw, SAUABI_NOISE_bw, SAUABI_NOISE_tw,
	SAUABI_NOISE_re, SAUABI_NOISE_vi, SAUABI_NOISE_bv, SAUABI_NOISE_NAMED
};

/* sau/program.h:69-80 */
enum {
	SAUABI_POPT_amp = 0, SAUABI_POPT_noise, SAUABI_POPT_wave, SAUABI_POPT_raseg,
	SAUABI_POPT_TYPES
};

/* sau/program.h:93-99 */
enum {
	SAUABI_POPP_TIME  = 1<<0,
	SAUABI_POPP_MODE  = 1<<1,
	SAUABI_POPP_PHASE = 1<<2,
	SAUABI_POPP_SEED  = 1<<3,
};

/* sau/program.h:25-29 */
enum {
	SAUABI_TIMEP_SET      = 1<<0,
	SAUABI_TIMEP_DEFAULT  = 1<<1,
	SAUABI_TIMEP_IMPLICIT = 1<<2,
};

/* sau/program.h:134-163 */
enum {
	SAUABI_RAS_F_URAND = 0, SAUABI_RAS_F_GAUSS, SAUABI_RAS_F_BIN,
	SAUABI_RAS_F_TERN, SAUABI_RAS_F_FIXED, SAUABI_RAS_F_ADDREC,
	SAUABI_RAS_FUNCTIONS
};
enum {
	SAUABI_RAS_O_PERLIN      = 1U<<0,
	SAUABI_RAS_O_HALFSHAPE   = 1U<<1,
	SAUABI_RAS_O_ZIGZAG      = 1U<<2,
	SAUABI_RAS_O_SQUARE      = 1U<<3,
	SAUABI_RAS_O_VIOLET      = 1U<<4,
	SAUABI_RAS_O_LINE_SET    = 1U<<6,
	SAUABI_RAS_O_FUNC_SET    = 1U<<7,
	SAUABI_RAS_O_LEVEL_SET   = 1U<<8,
	SAUABI_RAS_O_ASUBVAL_SET = 1U<<9,
};

/* sau/program.h:183-204: operator use types (index of the mod-list members) */
enum {
	SAUABI_POP_carr = 0, SAUABI_POP_camod, SAUABI_POP_amod, SAUABI_POP_ramod,
	SAUABI_POP_fmod, SAUABI_POP_rfmod, SAUABI_POP_pmod, SAUABI_POP_apmod,
	SAUABI_POP_fpmod, SAUABI_POP_NAMED
};

#define SAUABI_PVO_NO_ID  UINT16_MAX /* sau/program.h:168 */
#define SAUABI_PMODE_AMP_DIV_VOICES (1<<0) /* sau/program.h:246-248 */

/* sau/line.h:115-121 */
typedef struct sauabi_Line {
	float v0, vt;
	uint32_t pos, end;
	uint32_t time_ms;
	uint8_t type;
	uint8_t flags;
} sauabi_Line;

/* sau/program.h:36-39 */
typedef struct sauabi_Time {
	uint32_t v_ms;
	uint8_t flags;
} sauabi_Time;

/* sau/program.h:126-132 */
typedef struct sauabi_RasOpt {
	uint8_t line;
	unsigned flags: 10;
	unsigned func:  6;
	unsigned level: 8;
	uint32_t alpha;
} sauabi_RasOpt;

/* sau/program.h:177-180 */
typedef struct sauabi_ProgramIDArr {
	uint32_t count;
	uint32_t ids[];
} sauabi_ProgramIDArr;

/* sau/program.h:212-231 */
typedef struct sauabi_ProgramOpData {
	uint32_t id;
	uint32_t params;
	sauabi_Time time;
	sauabi_Line *pan;
	sauabi_Line *amp, *amp2;
	sauabi_Line *freq, *freq2;
	sauabi_Line *pm_a;
	uint32_t phase;
	uint32_t seed;
	uint8_t use_type;
	uint8_t type;
	union {
		uint8_t main;
		sauabi_RasOpt ras;
	} mode;
	/* SAU_POP__ITEMS order, modulator uses only (sau/program.h:183-193,228-230) */
	const sauabi_ProgramIDArr *camods, *amods, *ramods, *fmods, *rfmods,
	                          *pmods, *apmods, *fpmods;
} sauabi_ProgramOpData;

/* sau/program.h:233-241 */
typedef struct sauabi_ProgramEvent {
	uint32_t wait_ms;
	uint16_t vo_id;
	uint32_t carr_op_id;
	uint32_t op_count;
	uint32_t op_data_count;
	const void *op_list;
	const sauabi_ProgramOpData *op_data;
} sauabi_ProgramEvent;

/* sau/program.h:253-265 */
typedef struct sauabi_Program {
	const sauabi_ProgramEvent *events;
	size_t ev_count;
	uint16_t mode;
	uint16_t vo_count;
	uint32_t op_count;
	uint8_t op_nest_depth;
	uint32_t duration_ms;
	float ampmult;
	const char *name;
	void *mp;
	void *parse;
} sauabi_Program;

/* Fills `out` with sizeof/offsetof values in the same fixed order as
 * oracle/ref_harness.c:refwb_abi_layout(); returns the count. */
size_t saugen_abi_layout(uint32_t *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
