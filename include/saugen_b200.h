/* include/saugen_b200.h -- C ABI of the B200-native saugns generator back end.
 *
 * Plain pointers and sizes only; no torch / C++ types.  The three core entry
 * points replace, one for one, the reference generator interface
 *   sau_create_Generator   sau/generator.h:20-21   (sau/generator.c:200-217)
 *   sauGenerator_run       sau/generator.h:24-26   (sau/generator.c:905-973)
 *   sau_destroy_Generator  sau/generator.h:22      (sau/generator.c:222-228)
 * and saugns_b200/csrc/dropin.c re-exports them under the reference's own
 * symbol names so that the unmodified CLI (saugns.c:575-623) links against
 * this library instead of generator.o.
 *
 * All rendering happens in hand-written sm_100a CUDA kernels.  There is no
 * CPU fallback: every entry point fails (NULL / negative return) when no CUDA
 * device is usable.
 */
#ifndef SAUGEN_B200_H
#define SAUGEN_B200_H

#include <stddef.h>
#include <stdint.h>
#include "sau_program_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct saugen_Generator saugen_Generator;

/* The 12 pre-integrated wave tables + per-wave coefficients, as built on the
 * host by the front-end library (sau/wave.c:49-66,105-221; sau/wave.h:33-70).
 * They are INPUT DATA of the generator path: the drop-in passes libsau's own arrays
 * (sauWave_piluts / sauWave_picoeffs after sau_global_init_Wave, sau/generator.c:215);
 * a caller without libsau in its process loads the file those arrays were written to
 * (saugen_wave_tables_load; saugns_b200/data/sau_wave_tables.bin, tools/make_wave_tables.py).
 * The back end never regenerates them; saugen_create fails on NULL. */
typedef struct saugen_WaveTables {
	const float *pilut[SAUABI_WAVE_NAMED];   /* 2048 floats each */
	float amp_scale[SAUABI_WAVE_NAMED];
	float amp_dc[SAUABI_WAVE_NAMED];
	int32_t phase_adj[SAUABI_WAVE_NAMED];
} saugen_WaveTables;

/* Reads a table file into one malloc'd block (NULL on error); release with _free. */
saugen_WaveTables *saugen_wave_tables_load(const char *path);
void saugen_wave_tables_free(saugen_WaveTables *t);

/* Creation options (zero-initialise for defaults). */
typedef struct saugen_Options {
	int device;              /* CUDA device ordinal */
	void *stream;            /* cudaStream_t to launch on; NULL = own stream */
	uint32_t voice_begin;    /* render only voices [voice_begin, voice_end) ... */
	uint32_t voice_end;      /* ... 0,0 = all (multi-GPU voice sharding) */
	uint32_t max_call_len;   /* largest buf_len that will be passed; 0 = 256 ms */
	uint32_t sched;          /* 0 = auto; 1 = one warp per voice; 2 = persistent grid taking
	                          * (time unit, voice) tickets; 3 = balanced: one contiguous
	                          * range of (voice, block) items per resident warp (auto picks
	                          * it when the voices need more than one wave of warps) */
	uint32_t pcm_big_endian; /* 1 = big-endian int16 samples: the AU stream `saugns -o -` writes
	                          * (saugns.c:508-511; the swap of player/sndfile.c:160-168 is folded
	                          * into the mix epilogue); 0 = host order, as sauGenerator_run */
} saugen_Options;

/* == sau_create_Generator(prg, srate).  Borrows `prg` until destroy. */
saugen_Generator *saugen_create(const sauabi_Program *prg, uint32_t srate,
		const saugen_WaveTables *tables, const saugen_Options *opt);

/* The flat program (the "instruction form" of
 * sau/parser/parseconv.h:282-331,544-571 taken one step further): everything
 * saugen_create derives from a sauProgram -- events and op-data with pointers turned
 * into indices and ms into samples, the per-voice bytecode, the event timeline -- as
 * one relocatable blob.  saugen_flatten needs no GPU and returns the blob's size
 * (call with blob = NULL to ask); saugen_create_flat instantiates it on any device
 * without the sauProgram: parse and flatten once, render anywhere (other
 * processes, other GPUs).  0 / NULL on error. */
size_t saugen_flatten(const sauabi_Program *prg, uint32_t srate, void *blob, size_t cap);
saugen_Generator *saugen_create_flat(const void *blob, size_t size,
		const saugen_WaveTables *tables, const saugen_Options *opt);

/* Which voices must stay on one generator when a script is voice-sharded: parseconv re-homes an
 * operator whose voice slot was recycled (a later `@label` update gives the same op id a new
 * voice); voices linked by such hand-overs share operator state.  group_of_voice[v] (vo_count
 * entries, may be NULL) = smallest voice index of v's group; returns the number of hand-over
 * events, <0 on error.  No GPU needed.  (Within one generator a hand-over cuts the call into
 * separate render launches, so the state passes through a kernel boundary.) */
int saugen_voice_groups(const sauabi_Program *prg, uint32_t srate, uint32_t *group_of_voice);

/* == sau_destroy_Generator(o); NULL-safe. */
void saugen_destroy(saugen_Generator *o);

/* == sauGenerator_run(o, buf, buf_len, stereo, out_len) with HOST buffers:
 * launches the kernels, copies the PCM back, returns 1 while more signal
 * follows, 0 on the final call, <0 on CUDA error (out_len = 0). */
int saugen_run(saugen_Generator *o, int16_t *buf, size_t buf_len, int stereo,
		size_t *out_len);

/* Same call with the PCM left in device memory (*dev_pcm, valid until the
 * next call on this generator).  Only a 32-byte status record crosses PCIe. */
int saugen_run_device(saugen_Generator *o, size_t buf_len, int stereo,
		int16_t **dev_pcm, size_t *out_len);

/* Batched form (SURVEY.md section 8b "batch gap"): advance n independent
 * generators by one call each with ONE pair of kernel launches.  bufs[i] may
 * be NULL (device-resident).  more[i] receives each generator's return. */
int saugen_run_many(saugen_Generator *const *gens, size_t n, int16_t *const *bufs,
		size_t buf_len, int stereo, size_t *out_lens, int *more);

/* The same call in two halves (a batch owns a stream and the launch scratch):
 * begin() plans, uploads, launches and queues the read-backs without waiting,
 * end() waits, hands out the PCM and returns what saugen_run_many returns.  A
 * driver alternates two batches so that the host side of one call (planning,
 * admitting / retiring generators, consuming PCM) overlaps the kernels of the
 * other.  dest_pinned: every bufs[i] is page-locked memory (saugen_pinned_alloc)
 * and receives the device-to-host copy directly, with no staging copy. */
typedef struct saugen_Batch saugen_Batch;
saugen_Batch *saugen_batch_create(int device);
void saugen_batch_destroy(saugen_Batch *b);
int saugen_batch_begin(saugen_Batch *b, saugen_Generator *const *gens, size_t n,
		int16_t *const *bufs, size_t buf_len, int stereo, int dest_pinned);
int saugen_batch_end(saugen_Batch *b, size_t *out_lens, int *more);
/* Page-locked host memory from the library's pool (recycled, not returned to CUDA). */
void *saugen_pinned_alloc(size_t bytes);
void saugen_pinned_free(void *p);

/* ---- the native batched front end (saugns_b200/csrc/batch_driver.cpp) ----
 * What Player_run (saugns.c:575-623) + the script loop (saugns.c:648-659) + player/sndfile.c do
 * for one script at a time, for n independent programs on one GPU: live sets of generators
 * advanced with one render + one mix launch per call, each program's PCM rendered into ONE
 * page-locked array that takes the device-to-host copies directly. */
typedef struct saugen_BatchOptions {
	int device;              /* CUDA device ordinal */
	uint32_t call_len;       /* frames per generator call; 0 = 4 x 256 ms (results do not depend on it) */
	uint32_t group;          /* generators per live set; 0 = 128 */
	uint32_t depth;          /* alternating live sets; 0 = 2 */
	uint32_t mono;           /* 1 = mono downmix (saugns --mono) */
	uint32_t io_threads;     /* saugen_render_batch_wav: file writer threads; 0 = 4 */
} saugen_BatchOptions;
/* Receives a finished program's whole PCM (frames x channels int16, interleaved) on the driver
 * thread.  The array is page-locked memory of the library's pool and now belongs to the sink:
 * release it with saugen_pinned_free (from any thread) when done with it. */
typedef void (*saugen_pcm_sink)(void *user, size_t index, int16_t *pcm, size_t frames, int channels);
/* 0 = every program rendered; <0 = error (saugen_batch_last_error).  sink NULL: the PCM is
 * delivered to host memory and dropped. */
int saugen_render_batch(const sauabi_Program *const *prgs, size_t n, uint32_t srate,
		const saugen_WaveTables *tables, const saugen_BatchOptions *opt, saugen_pcm_sink sink, void *user);
/* The same with the reference's WAV files as the sink (player/sndfile.c:63-109: 44-byte header,
 * little-endian int16): program i goes to paths[i]. */
int saugen_render_batch_wav(const sauabi_Program *const *prgs, size_t n, uint32_t srate,
		const saugen_WaveTables *tables, const saugen_BatchOptions *opt, const char *const *paths);
const char *saugen_batch_last_error(void);

/* Voice-sharded rendering across GPUs: produce this rank's partial float mix in device
 * memory -- L plane at [0, buf_len), R plane at [row_len, row_len + buf_len), row_len =
 * max_call_len rounded up to 4 -- valid until the call after the next (two blocks alternate);
 * the caller reduces the planes over ranks (NCCL sum) and converts on the root.  The block is
 * followed by SAUGEN_MIX_TAIL floats of the caller's own (saugns_b200/multigpu.py puts every
 * rank's `more` / out_len there so that ONE collective carries data and control). */
#define SAUGEN_MIX_TAIL 64
int saugen_run_mix(saugen_Generator *o, size_t buf_len, float **dev_mix,
		size_t *out_len);
int saugen_mix_to_pcm(saugen_Generator *o, const float *dev_mix, size_t buf_len,
		int stereo, int16_t *host_buf);

/* ---- introspection for parity tests ------------------------------------ */

typedef struct saugen_LineView {
	float v0, vt;
	uint32_t pos, end;
	uint32_t type, flags;
} saugen_LineView;

/* Same field meaning as oracle/ref_harness.c:RefOpState. */
typedef struct saugen_OpView {
	uint32_t inited, type, flags, time;
	saugen_LineView amp, amp2, pan, freq, freq2, pm_a;
	uint32_t i0, i1;
	uint32_t mode, oscflags;
	double prev_Is;
	float prev_s, fb_s;
	uint32_t alpha, rate2x;
} saugen_OpView;

int saugen_read_op(saugen_Generator *o, uint32_t op_id, saugen_OpView *out);
int saugen_read_voice(saugen_Generator *o, uint32_t vo_id, uint32_t out[4]);
/* Float carrier rows (s = carrier*amp_scale, r = s*pan) written by the last
 * call for one voice: n = frames of that call.  The r row exists only where the
 * voice's pan moves (sweep or pan modulators); a constant pan is applied in the
 * mix kernel and leaves r untouched. */
int saugen_read_voice_rows(saugen_Generator *o, uint32_t vo_id, float *s, float *r,
		size_t n);
/* Counters: [0] render launches, [1] launches of the call's other kernels (mix, prologue), [2] work buffers
 * per voice warp, [3] wave mask */
int saugen_counters(saugen_Generator *o, uint64_t out[4]);
/* Device time of render_kernel / mix_kernel accumulated since timing was
 * switched on (CUDA events on the launch stream; for bench.py's roofline). */
int saugen_set_timing(saugen_Generator *o, int on);
int saugen_kernel_ms(saugen_Generator *o, double out[2]);
/* Device arithmetic self-test: how many inputs a hand-expanded fast-path
 * primitive (exact int-divisor division, 32-bit lrintf) gets differently from
 * the plain statement it replaces, over all 2^32 divisors / float patterns. */
long long saugen_selftest(int device, const saugen_WaveTables *tables);
float saugen_amp_scale(saugen_Generator *o);
const char *saugen_last_error(void);
int saugen_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
