/* device_types.h -- data layout in HBM shared by the host runtime and the
 * sm_100a kernels.  Everything the generator needs at run time lives in
 * device memory: per-operator state (mirrors OperatorNode, reference
 * sau/generator.c:45-88), per-voice state (VoiceNode, :97-102), the flattened
 * event / op-data records (sau/program.h:212-241) and the per-voice bytecode
 * the host compiles from the modulator lists.
 */
#pragma once
#include <stdint.h>

namespace saugen {

constexpr int CHUNK = 128;            // samples per interpreter pass (4 per lane)
constexpr int SPL = CHUNK / 32;       // consecutive samples owned by one lane
constexpr int REF_BLOCK = 1024;       // BUF_LEN, sau/generator.c:28
constexpr int WAVE_LEN = 2048;        // sau/wave.h:18-19
constexpr int NUM_WAVES = 12;
constexpr int MAX_NEST = 24;          // len-stack depth per warp
constexpr uint32_t NO_BUF = 0xFF;
/* Carrier rows of a call are stored FRAME-TILE major: tile t (frames 128 t .. 128 t + 127)
 * holds one 512-byte piece per voice, voices side by side:
 *     rows[(t * n_local_voices + voice) * 128 + (frame & 127)]
 * The render kernel writes whole pieces (a chunk is 128 frames), the mix kernel reads
 * a tile as ONE contiguous stream in voice order (it must add in voice order). */
constexpr int ROW_TILE = 128;
#if defined(__CUDACC__)
__host__ __device__
#endif
inline size_t row_index(uint32_t frame, uint32_t tile_stride) {
	return (size_t) (frame / ROW_TILE) * tile_stride + (frame % ROW_TILE);
}

/* sauLine run-time state (sau/line.h:115-121 minus time_ms), split so that the
 * four words a fill needs come in with one 128-bit shared-memory load. */
struct LineState {
	float v0, vt;
	uint32_t pos, end;
};
/* lmeta word of a line: type | flags << 8 | blk_done << 16.  blk_done = the
 * position already wrapped inside the current REF_BLOCK (see line_eval). */
#define LM_TYPE(m)  ((m) & 0xffu)
#define LM_FLAGS(m) (((m) >> 8) & 0xffu)
#define LM_BLK(m)   (((m) >> 16) & 1u)
#define LM_PACK(type, flags, blk) (((type) & 0xffu) | (((flags) & 0xffu) << 8) | (((blk) & 1u) << 16))

enum { LINE_AMP = 0, LINE_AMP2, LINE_PAN, LINE_FREQ, LINE_FREQ2, LINE_PMA, LINE_COUNT };

/* Operator node flags (sau/generator.c:39-43) + ours. */
enum {
	ON_INIT      = 1 << 0,
	ON_TIME_INF  = 1 << 2,
	ON_PMA_RUN   = 1 << 4,   // pm_a line is being run for the current REF_BLOCK
};
enum { OSC_RESET_DIFF = 1 << 0 };   // sau/generator/wosc.h:37-38

/* 192 bytes, every hot group 16-byte aligned: line[i] at 16*i, the
 * {time, type/flags/mode/oscflags, prev_s, fb_s} group at 144 and everything an
 * oscillator carries from chunk to chunk, {i0, i1, prev_Is}, in ONE 128-bit
 * group at 160 (one load, one store by the lane that holds the chunk's last sample). */
struct __align__(16) OpState {
	LineState line[LINE_COUNT];
	uint32_t lmeta[LINE_COUNT];
	float linv[LINE_COUNT];   // 1.f / (float) end, kept current by apply_event
	uint32_t time;
	uint8_t type;          // SAUABI_POPT_*
	uint8_t flags;         // ON_*
	uint8_t mode;          // W: wave, N: noise type, R: line type
	uint8_t oscflags;      // W: OSC_RESET_DIFF, R: bit0 = rate2x
	float prev_s, fb_s;
	/* W: phasor.phase / prev_phase; N: n / prev; R: cycle_phase lo / hi */
	uint32_t i0, i1;
	double prev_Is;
	/* R options (sau/program.h:126-132) */
	uint16_t ras_flags;
	uint8_t ras_func, ras_level;
	uint32_t ras_alpha;
	uint32_t _pad[2];
};

enum { VN_INIT = 1 << 0 };
struct VoiceState {
	uint32_t duration;
	uint32_t carr_op;
	uint32_t flags;
	uint32_t ev_cursor;    // next entry of this voice's event list
	uint32_t code_off;     // current program (offset into GenDesc::code)
	uint32_t code_len;
	uint32_t ops_off;      // the program's operator list (offset into GenDesc::prog_ops)
	uint32_t ops_cnt;
	uint32_t carr_slot;    // carrier's slot in that list
	/* the voice's lowered steady plan as kept in GenDesc::plan_cache (render_kernel.cuh): generation
	 * (0 = none valid), the shared-memory layout its addresses are for */
	uint32_t plan_gen, plan_so, plan_st;
	uint32_t plan_left;    // whole blocks the kept plan still holds for (steady_plan's spans)
};

/* Flattened sauLine delta carried by an event (NULL pointer => present = 0). */
struct LineDelta {
	float v0, vt;
	uint32_t end_samples;  // sau_ms_in_samples(time_ms, srate, NULL)
	uint8_t type, flags, present, _pad;
};

struct OpDataRec {
	uint32_t id;
	uint32_t params;
	uint32_t time_samples;
	uint8_t time_flags;
	uint8_t type;
	uint8_t use_type;
	uint8_t mode_main;
	LineDelta line[LINE_COUNT];
	uint32_t phase, seed;
	uint16_t ras_flags;
	uint8_t ras_func, ras_level;
	uint32_t ras_alpha;
};

struct EventRec {
	uint32_t vo_id;
	uint32_t carr_op_id;
	uint32_t opdata_off, opdata_count;
	uint32_t code_off, code_len;    // voice program after this event
	uint32_t ops_off, ops_cnt;      // its operator list (Instr::op indexes it)
	uint32_t carr_slot;
};

/* Bytecode: the reference's recursive run_block walk (sau/generator.c:448-729)
 * flattened per voice.  a..e are work-buffer slots (gen_bufs indices). */
enum Opcode : uint8_t {
	I_ENTER = 1,   // op; a=out; flags: layer bits
	I_LEAVE,       // op; a=out
	I_ZERO,        // a=out (circular reference guard, generator.c:685-689)
	I_LINE,        // op; a=dst; b=mulbuf|NO_BUF; c=which line; d=1 run / 0 skip
	I_RANGE,       // a=par; b=r_par; c=mod   (generator.c:465-467)
	I_PHASOR,      // op; a=phase dst; b=freq; c=pm|NO_BUF; d=fpm|NO_BUF
	I_PMA,         // op; a=dst  (generator.c:485-490)
	I_WOSC,        // op; a=dst; b=phase; c=pm_a buf; flags HAS_APMODS
	I_CYCLOR,      // op; a=cycle dst; b=phase(float) dst; c=freq; d=pm; e=fpm
	I_RASG,        // op; a=in/out; b=cycle; c=pm_a buf; d,e=tmp; flags HAS_APMODS
	I_NOISE,       // op; a=dst
	I_MIX,         // a=out; b=in|NO_BUF(=1.0); c=amp; flags WAVEENV
	I_VPAN,        // op=carrier; a=pan dst; d=1 run (camods) / 0 decide at run time
	I_VOUT,        // op=carrier; a=carrier out; b=pan buf
	I_END,
	/* fused wave operator (run_block_wosc): a=out; b=freq buf (b, b+1 double as
	 * self-PM scratch); c=pm|NO_BUF; d=fpm|NO_BUF; e=parent freq|NO_BUF (HEAD) */
	I_WHEAD,       // ENTER + freq line -> b
	I_WTAIL,       // phase fill + amp line + oscillator + mix + LEAVE
	I_WLEAF,       // both, nothing leaves the registers
};
enum {
	F_LAYER       = 1 << 0,   // accumulate into out
	F_LAYER_PMA   = 1 << 1,   // layer = "pm_a buffer already filled" (run time)
	F_WAVEENV     = 1 << 2,
	F_HAS_APMODS  = 1 << 3,
	F_HAS_CAMODS  = 1 << 4,
	F_SKIP_FREQ2  = 1 << 5,   // the freq2 / amp2 / pm_a line was set at some point:
	F_SKIP_AMP2   = 1 << 6,   // keep its position bookkeeping (sauLine_skip)
	F_MAY_SELFMOD = 1 << 7,
	F_KEEP_FREQ   = 1 << 8,   // leaf carrier with pan modulators: they read its freq buffer
};
struct Instr {
	uint8_t opcode, a, b, c, d, e;
	uint16_t flags;
	uint32_t op;
	uint32_t aux;          // I_ENTER: index of the matching I_LEAVE
};

/* What the mix kernel needs to know about one voice in one inter-event segment
 * of a call: frames rendered, and its pan.  A pan that stands still during the
 * segment (no sweep, no pan modulators: only an event can change that, and
 * events start new segments) is a constant: the render kernel stores it here
 * and skips the r = s * pan row, the mix kernel redoes that one multiply
 * (mix_add, sau/generator.c:772-786; same operation, same bits).  A moving pan
 * is flagged PAN_DYNAMIC and read from rows_r. */
struct VoiceSeg {
	uint32_t len;
	uint32_t pan;                 // float bits, or PAN_DYNAMIC / PAN_UNSET
};
constexpr uint32_t PAN_UNSET = 0u - 1u;          // not decided yet (both are NaN patterns
constexpr uint32_t PAN_DYNAMIC = 0u - 2u;        // no sauLine value can take: v0 comes from a parser)

/* One generator's device-resident description. */
struct GenDesc {
	OpState *ops;
	VoiceState *voices;
	const EventRec *events;
	const OpDataRec *opdata;
	const Instr *code;
	const uint32_t *prog_ops;     // operator ids of every compiled voice program
	const uint32_t *vev_off;      // [vo_count+1] CSR into vev_idx
	const uint32_t *vev_idx;      // global event indices per voice, in order
	float *rows_s, *rows_r;       // carrier rows of a call, frame-tile major (ROW_TILE above; rows_r:
	                              // only for voices whose pan moves in the segment, see VoiceSeg)
	VoiceSeg *vlen;               // [seg][n_local_voices] per segment of a call
	uint32_t *status;             // (host bookkeeping; the kernels use CallDesc::status)
	uint32_t vlen_cap;            // segments the vlen/status arrays can hold
	uint32_t *progress;           // [n_local_voices] units done in the current call (ticketed launches)
	uint32_t *ticket;             // next (unit, voice) ticket of the current call
	float *mix;                   // [2][row_len] float mix planes (L, R)
	int16_t *pcm;                 // [row_len*2]
	uint32_t vo_count, op_count;
	uint32_t voice_begin, voice_end;
	uint32_t row_len;             // frames per row (max call length)
	uint32_t row_stride;          // floats between consecutive frame tiles = n_local_voices * ROW_TILE
	uint32_t nbufs;               // work buffers per voice warp
	uint32_t srate;
	float coeff;                  // (float)(2^32 / srate), wosc.h:30, rasg.h:27
	float amp_scale;
	uint32_t wave_mask;           // waves referenced by any op-data
	const float *tables;          // 12 x 2048 floats, then a WaveCoeffs
	float *tap;                   // debug (saugen_debug_tap): [op_count][row_len], every operator's output buffer
	uint4 *plan_cache;            // [n_local_voices][1 + 2 * plan_cache_recs] 16-byte words: {generation, records,
	uint32_t plan_cache_recs;     // fused shape, -} and the voice's last stable lowered plan (render_kernel.cuh)
	/* a team that spans several CTAs (render_team.cuh): per voice 16 words {arrivals, epoch, ...} zeroed before
	 * every launch, and a mailbox (TEAM_MAIL_* below) the leader's CTA and the others exchange through */
	uint32_t *team_hdr;
	unsigned char *team_mail;
	uint32_t team_mail_stride;
	float *team_cache;            // teams (render_team.cuh): [n_local_voices][TEAM_SLOTS][team_cache_stride] floats, the
	uint32_t team_cache_stride;   // values that pass from one phase of a stretch to a later one
};

/* Per-wave constants derived from sauWave_picoeffs (sau/wave.h:33-70,144-149),
 * stored after the 12 tables in the device table block. */
struct WaveCoeffs {
	float diff_scale[NUM_WAVES];   // amp_scale * 0.125f * 2^32
	float diff_offset[NUM_WAVES];  // amp_dc
	float amp256[NUM_WAVES];       // amp_scale * 256 (reset path, wosc.h:224)
	int32_t phase_adj[NUM_WAVES];
};

/* One inter-event stretch of a call (sau/generator.c:915-949). */
struct SegDesc {
	uint32_t start;    // frame offset inside the call
	uint32_t len;
	uint32_t ev_end;   // events with index < ev_end are due at the segment start
};

/* A schedulable stretch of a segment: starts at a multiple of REF_BLOCK inside
 * the segment. */
struct UnitDesc {
	uint32_t seg;      // segment index inside the call
	uint32_t off;      // frame offset inside the segment
	uint32_t len;
};

/* One sauGenerator_run call of one generator. */
struct CallDesc {
	const GenDesc *gen;
	uint32_t call_len;
	uint32_t nseg;
	uint32_t seg_off;      // first SegDesc of this call
	uint32_t task_base;    // index of this call's first voice task in the launch
	uint32_t stereo;
	uint32_t unit_off;     // first UnitDesc of this call
	uint32_t nunits;
	uint32_t more_launches;   // 1 = another render launch of this call follows (hand-over cut, runtime.cpp:
	                          // plan_call): the voices' alive flag is set by the last launch only
	int16_t *pcm;             // where this call's PCM goes (two alternate: the call after this one may
	                          // already be rendering while the caller reads this one's, runtime.cpp run-ahead)
	uint32_t *status;         // [0]=any voice still alive, [1+seg]=per-segment max len; one per call slot (the
	                          // read-back of call k runs on the copy stream while call k+1 renders)
};

/* teams (render_team.cuh): outputs of one plan that cross a phase = cache slots per voice; the most
 * lead-in chunks a member renders before its range */
constexpr uint32_t TEAM_SLOTS = 12, TEAM_LEAD_CHUNKS = 7;

/* the mailbox of a team over several CTAs: the leader's command block, master plan and operator states (what
 * the other CTAs mirror in their own shared memory), every member's counts of a phase, and the last member's
 * oscillator state on its way back to the voice */
constexpr uint32_t TEAM_MAX_CTAS = 8, TEAM_MAX_MEMBERS = TEAM_MAX_CTAS * 28, TEAM_MAIL_OPS = 32;
constexpr uint32_t TEAM_CMD_BYTES = 352;       /* 64 + a word per record of the master plan (render_team.cuh:TC_INFO) + 16 */
__host__ __device__ inline uint32_t team_mail_plan_off() { return TEAM_CMD_BYTES; }
__host__ __device__ inline uint32_t team_mail_ops_off(uint32_t plan_bytes) { return TEAM_CMD_BYTES + plan_bytes; }
__host__ __device__ inline uint32_t team_mail_counts_off(uint32_t plan_bytes, uint32_t max_ops) {
	return team_mail_ops_off(plan_bytes) + max_ops * 192u;
}
__host__ __device__ inline uint32_t team_mail_final_off(uint32_t plan_bytes, uint32_t max_ops) {
	return team_mail_counts_off(plan_bytes, max_ops) + TEAM_MAX_MEMBERS * TEAM_MAIL_OPS * 4u;
}
__host__ __device__ inline uint32_t team_mail_bytes(uint32_t plan_bytes, uint32_t max_ops) {
	return (team_mail_final_off(plan_bytes, max_ops) + TEAM_MAIL_OPS * 5u * 4u + 255u) & ~255u;
}

/* A call's descriptors small enough to travel as kernel parameters (prologue_kernel): no
 * host-to-device copy in front of the render launch. */
constexpr uint32_t INLINE_SEGS = 16, INLINE_UNITS = 64;
struct InlineCall {
	CallDesc cd;
	uint32_t nseg, nunits;
	SegDesc segs[INLINE_SEGS];
	UnitDesc units[INLINE_UNITS];
};
/* What the prologue does before a call's render launch (one small kernel instead of a copy, a
 * memset and the run-ahead snapshot copy): descriptors -> device, [vlen][progress] and the slot's
 * status zeroed, operator / voice state -> snapshot. */
struct PrologueArgs {
	CallDesc *d_call;
	SegDesc *d_segs;
	UnitDesc *d_units;
	uint32_t *zero_a, *zero_b, *zero_c;
	uint32_t zero_a_words, zero_b_words, zero_c_words;
	const uint4 *snap_src;
	uint4 *snap_dst;
	uint32_t snap_n16;        // 0 = no snapshot
};

} // namespace saugen
