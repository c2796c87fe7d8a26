/* batch_cli.c -- `saugns_b200_batch`: render many SAU scripts to WAV files on one GPU.
 *
 *     saugns_b200_batch [-r srate] [--mono] [-d device] -o <outdir> script.sau ...
 *     saugns_b200_batch ... -l <file listing one script path per line>
 *
 * The scripts go through the reference's OWN front end (sau_build_Program: lexer, parser,
 * program builder -- unmodified host C code, as the north star keeps it), the resulting
 * sauProgram objects through saugen_render_batch_wav.  It is the batched counterpart of the
 * reference's script loop (saugns.c:648-659: one Player_run per script): outdir/<name>.wav
 * holds exactly what `saugns -m -d -o <name>.wav <name>.sau` writes.  Linked like the drop-in
 * (INTEGRATION.md): libsau's objects without generator.o + libsaugen_b200.so; built by
 * oracle/Makefile `batchcli` because it contains reference objects. */
#define _POSIX_C_SOURCE 200809L
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/saugen_b200.h"

/* the reference front end (sau/program.h:268-270, sau/script.h:135-141, sau/wave.h) */
typedef struct sauScriptArg {
	const char *str;
	bool is_path : 1;
	bool no_time : 1;
	void *predef;
	size_t predef_count;
} sauScriptArg;
sauabi_Program *sau_build_Program(const sauScriptArg *arg);
void sau_discard_Program(sauabi_Program *o);
extern float *const sauWave_piluts[SAUABI_WAVE_NAMED];
struct sauWaveCoeffs { float amp_scale; float amp_dc; int32_t phase_adj; };
extern const struct sauWaveCoeffs sauWave_picoeffs[SAUABI_WAVE_NAMED];
void sau_global_init_Wave(void);
/* defined by the generator in the reference, used by its parser (csrc/dropin.c) */
const char *const sauNoise_names[SAUABI_NOISE_NAMED + 1] = {"wh", "gw", "bw", "tw", "re", "vi", "bv", NULL};

static char *base_name(const char *path) {
	const char *b = strrchr(path, '/');
	b = b ? b + 1 : path;
	char *out = strdup(b);
	char *dot = strrchr(out, '.');
	if (dot) *dot = 0;
	return out;
}

int main(int argc, char **argv) {
	uint32_t srate = 96000;
	saugen_BatchOptions opt;
	memset(&opt, 0, sizeof opt);
	const char *outdir = NULL, *list = NULL;
	char **files = calloc((size_t) argc + 1, sizeof(char*));
	size_t nf = 0, cap = (size_t) argc + 1;
	for (int i = 1; i < argc; ++i) {
		if (!strcmp(argv[i], "-r") && i + 1 < argc) srate = (uint32_t) atoi(argv[++i]);
		else if (!strcmp(argv[i], "-o") && i + 1 < argc) outdir = argv[++i];
		else if (!strcmp(argv[i], "-d") && i + 1 < argc) opt.device = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-l") && i + 1 < argc) list = argv[++i];
		else if (!strcmp(argv[i], "--mono")) opt.mono = 1;
		else files[nf++] = argv[i];
	}
	if (list) {
		FILE *f = fopen(list, "r");
		char line[4096];
		if (!f) { fprintf(stderr, "saugns_b200_batch: cannot read %s\n", list); return 1; }
		while (fgets(line, sizeof line, f)) {
			line[strcspn(line, "\r\n")] = 0;
			if (!line[0]) continue;
			if (nf + 1 >= cap) { cap *= 2; files = realloc(files, cap * sizeof(char*)); }
			files[nf++] = strdup(line);
		}
		fclose(f);
	}
	if (!outdir || !nf || !srate) {
		fprintf(stderr, "usage: saugns_b200_batch [-r srate] [--mono] [-d device] -o <outdir> (script.sau ... | -l list)\n");
		return 1;
	}
	sauabi_Program **prgs = calloc(nf, sizeof *prgs);
	char **paths = calloc(nf, sizeof *paths);
	size_t n = 0;
	for (size_t i = 0; i < nf; ++i) {
		sauScriptArg arg;
		memset(&arg, 0, sizeof arg);
		arg.str = files[i]; arg.is_path = true; arg.no_time = true;     /* as `saugns -d` */
		sauabi_Program *p = sau_build_Program(&arg);
		if (!p) { fprintf(stderr, "saugns_b200_batch: skipping %s\n", files[i]); continue; }
		char *b = base_name(files[i]);
		paths[n] = malloc(strlen(outdir) + strlen(b) + 8);
		sprintf(paths[n], "%s/%s.wav", outdir, b);
		free(b);
		prgs[n++] = p;
	}
	saugen_WaveTables t;
	sau_global_init_Wave();                         /* as sau_create_Generator does, generator.c:215 */
	for (int w = 0; w < SAUABI_WAVE_NAMED; ++w) {
		t.pilut[w] = sauWave_piluts[w];
		t.amp_scale[w] = sauWave_picoeffs[w].amp_scale;
		t.amp_dc[w] = sauWave_picoeffs[w].amp_dc;
		t.phase_adj[w] = sauWave_picoeffs[w].phase_adj;
	}
	int r = saugen_render_batch_wav((const sauabi_Program *const*) prgs, n, srate, &t, &opt,
			(const char *const*) paths);
	if (r < 0) fprintf(stderr, "saugns_b200_batch: error: %s\n", saugen_batch_last_error());
	for (size_t i = 0; i < n; ++i) { sau_discard_Program(prgs[i]); free(paths[i]); }
	return r < 0 ? 1 : 0;
}
