/* wavetab.cpp -- built-in pre-integrated wave tables for stand-alone use of
 * the back end (bench, tests, batch rendering without libsau).
 *
 * In the drop-in configuration the tables are NOT built here: the shim passes
 * the arrays the front-end library built on the host (sauWave_piluts,
 * sau/wave.c:49-62), because the reference's -ffast-math build of wave.c
 * (vectorised libm sin, fused scaling) is what defines them bit for bit
 * (SURVEY.md Appendix B.4).  This file restates the same construction
 * (sau/wave.c:105-214, fill_It :77-98) in strict IEEE arithmetic;
 * tests/test_wavetab.py reports how close the two are (a few last-place
 * differences on the sqrt/integrated tables).
 */
#include <math.h>
#include <string.h>
#include <mutex>
#include "../../include/saugen_b200.h"

namespace saugen {

namespace {
constexpr int LEN = SAUABI_WAVE_LEN, HALF = LEN / 2, QUARTER = LEN / 4;

struct Tabs {
	float sin_[LEN], sqr[LEN], tri[LEN], pitri[LEN], ean[LEN], piean[LEN];
	float saw[LEN], par[LEN], pipar[LEN];
	float srs[LEN], pisrs[LEN], cat[LEN], picat[LEN], mto[LEN], pimto[LEN];
	float hsi[LEN], pihsi[LEN], spa[LEN], pispa[LEN];
};
Tabs T;
saugen_WaveTables W;
std::once_flag once;

/* running integral of (in - mean), rescaled to +/-1 peak (wave.c:77-98) */
void integrate(float *out, const float *in) {
	double dc = 0.0;
	for (int i = 0; i < LEN; ++i) dc += in[i];
	dc /= LEN;
	double sum = 0.0;
	float lb = 0.f, ub = 0.f;
	const float ivscale = 1.f / (LEN * 0.125f);
	for (int i = 0; i < LEN; ++i) {
		sum += in[i] - dc;
		float x = (float) (sum * ivscale);
		if (x < lb) lb = x;
		if (x > ub) ub = x;
		out[i] = x;
	}
	const float out_scale = 1.f / ((ub - lb) * 0.5f);
	const float out_dc = -(ub + lb) * 0.5f;
	for (int i = 0; i < LEN; ++i) out[i] = (out[i] + out_dc) * out_scale;
}

void build() {
	const double PI = 3.14159265358979323846;
	for (int i = 0; i < HALF; ++i) {
		const double x = (double) (i * (1.f / HALF));
		const float sin_x = (float) sin(PI * x);
		T.sin_[i] = sin_x;
		T.sin_[i + HALF] = -sin_x;
		T.sqr[i] = 1.f;
		const float srs_x = sqrtf(sin_x);
		T.srs[i] = srs_x;
		T.hsi[i] = sin_x * 2 - 1.f;
		T.mto[i] = srs_x * 2 - 1.f;
		const float spa_x = (float) sin(PI * 0.5f * (1 + x));
		T.spa[i + QUARTER] = spa_x * 2 - 1.f;
	}
	for (int i = 0; i < HALF; ++i) {
		const double x = (double) (i * (1.f / (HALF - 1)));
		const double x_rev = (double) ((HALF - i) * (1.f / HALF));
		T.par[i + QUARTER] = (float) ((x_rev * x_rev) * 2.f - 1.f);
		T.saw[i] = (float) (1.f - x);
	}
	T.par[HALF + QUARTER] = -1.f;
	T.spa[HALF + QUARTER] = -1.f;
	for (int i = 0; i < QUARTER; ++i) {
		const double x = (double) (i * (1.f / QUARTER));
		const double x_rev = (double) ((QUARTER - i) * (1.f / QUARTER));
		T.pitri[i] = (float) ((x * x) - 1.f);
		T.pitri[i + QUARTER] = (float) (1.f - (x_rev * x_rev));
		T.tri[i] = (float) x;
		T.tri[i + QUARTER] = (float) x_rev;
		T.par[i] = T.par[HALF - i];
		T.par[i + HALF + QUARTER] = T.par[HALF + QUARTER - i];
		T.spa[i] = T.spa[HALF - i];
		T.spa[i + HALF + QUARTER] = T.spa[HALF + QUARTER - i];
	}
	for (int i = HALF; i < LEN; ++i) {
		T.pitri[i] = -T.pitri[i - HALF];
		T.tri[i] = -T.tri[i - HALF];
		T.sqr[i] = -1.f;
		T.saw[i] = -T.saw[(LEN - 1) - i];
		T.hsi[i] = -1.f;
		T.mto[i] = -1.f;
		T.srs[i] = -T.srs[i - HALF];
	}
	const float ean_dc_adj = (float) ((1.14603185654 - 1.f) / 2.f);
	const float ean_scale_adj = (float) (1.f / 1.07301592827);
	for (int i = 0; i < LEN; ++i) {
		T.ean[i] = (T.sin_[i] + T.par[i] - T.tri[i] + ean_dc_adj) * ean_scale_adj;
		T.cat[i] = T.sin_[i] + T.mto[i] - T.srs[i];
	}
	integrate(T.piean, T.ean);
	integrate(T.picat, T.cat);
	integrate(T.pipar, T.par);
	integrate(T.pisrs, T.srs);
	integrate(T.pimto, T.mto);
	integrate(T.pihsi, T.hsi);
	integrate(T.pispa, T.spa);

	/* wave -> pre-integrated table (wave.c:49-62) and coefficients (wave.h:33-70) */
	const float *pil[SAUABI_WAVE_NAMED] = {
		T.sin_, T.pitri, T.pisrs, T.tri, T.piean, T.picat, T.ean, T.pipar, T.pimto, T.par,
		T.pihsi, T.pispa,
	};
	static const double amp_scale[SAUABI_WAVE_NAMED] = {
		1.27324153848, 1.00097751711, 1.52547437578, 2.00000000000, 1.20275515347,
		1.37070880305, 1.26113986272 * -1, 1.02639326795, 1.57268451738,
		1.00048851979 * -1, 1.40333871035, 1.07213756312,
	};
	static const double amp_dc[SAUABI_WAVE_NAMED] = {
		0.0, 0.0, 0.0, 0.0, -0.24257955076, -0.23725526633, 0.0, -0.33333333333,
		-0.23724704918, 0.0, -0.36334126990, 0.27322393756,
	};
	static const int32_t phase_adj[SAUABI_WAVE_NAMED] = {
		INT32_MIN / 2, 0, 0, INT32_MIN / 2, 0, 0, -(INT32_MIN / 2), 0, 0, -(INT32_MIN / 2), 0, 0,
	};
	for (int w = 0; w < SAUABI_WAVE_NAMED; ++w) {
		W.pilut[w] = pil[w];
		W.amp_scale[w] = (float) amp_scale[w];
		W.amp_dc[w] = (float) amp_dc[w];
		W.phase_adj[w] = phase_adj[w];
	}
}
} // namespace

const saugen_WaveTables *builtin_wave_tables() {
	std::call_once(once, build);
	return &W;
}

} // namespace saugen
