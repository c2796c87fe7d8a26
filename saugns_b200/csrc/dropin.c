/* dropin.c -- re-exports the B200 back end under the reference's own symbols
 * so that the UNMODIFIED saugns CLI and libsau front end link against it in
 * place of sau/generator.o:
 *
 *   sau_create_Generator   sau/generator.h:20-21
 *   sau_destroy_Generator  sau/generator.h:22
 *   sauGenerator_run       sau/generator.h:24-26
 *   sauNoise_names         defined inside the generator's private header
 *                          (sau/generator/noise.h:18-21) but used by the
 *                          parser (sau/parser.c:102) -- must come from here.
 *
 * The wave tables are the ones libsau builds on the host
 * (sau_global_init_Wave, sau/wave.c:105; called as in sau/generator.c:215).
 */
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/saugen_b200.h"

/* provided by libsau (sau/wave.h:84-97,151; sau/error.c) */
extern float *const sauWave_piluts[SAUABI_WAVE_NAMED];
struct sauWaveCoeffs { float amp_scale; float amp_dc; int32_t phase_adj; };
extern const struct sauWaveCoeffs sauWave_picoeffs[SAUABI_WAVE_NAMED];
void sau_global_init_Wave(void);
void sau_error(const char *label, const char *fmt, ...);

const char *const sauNoise_names[SAUABI_NOISE_NAMED + 1] = {
	"wh", "gw", "bw", "tw", "re", "vi", "bv", NULL
};

typedef struct sauGenerator sauGenerator;   /* opaque to callers */

sauGenerator *sau_create_Generator(const sauabi_Program *prg, uint32_t srate) {
	saugen_WaveTables t;
	sau_global_init_Wave();
	for (int w = 0; w < SAUABI_WAVE_NAMED; ++w) {
		t.pilut[w] = sauWave_piluts[w];
		t.amp_scale[w] = sauWave_picoeffs[w].amp_scale;
		t.amp_dc[w] = sauWave_picoeffs[w].amp_dc;
		t.phase_adj[w] = sauWave_picoeffs[w].phase_adj;
	}
	saugen_Generator *g = saugen_create(prg, srate, &t, NULL);
	if (!g) sau_error("generator", "B200 back end: %s", saugen_last_error());
	return (sauGenerator*) g;
}

void sau_destroy_Generator(sauGenerator *o) {
	saugen_destroy((saugen_Generator*) o);
}

bool sauGenerator_run(sauGenerator *o, int16_t *buf, size_t buf_len, bool stereo,
		size_t *out_len) {
	int r = saugen_run((saugen_Generator*) o, buf, buf_len, stereo, out_len);
	if (r < 0) {
		sau_error("generator", "B200 back end: %s", saugen_last_error());
		if (out_len) *out_len = 0;
		return false;
	}
	return r > 0;
}
