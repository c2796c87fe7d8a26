/* batch_driver.cpp -- the native batched front end of the generator back end (SURVEY.md section 8f
 * rank 1): what Player_run does for one script at a time (saugns.c:575-623, looped over the
 * scripts at saugns.c:648-659) and what player/sndfile.c writes (63-109: 44-byte RIFF/WAVE header,
 * little-endian int16 data), for thousands of independent programs on one GPU.
 *
 * saugen_render_batch keeps `depth` live sets of up to `group` generators and advances each set
 * with ONE render + ONE mix launch per call (saugen_batch_begin / _end): while one set's kernels
 * run, the driver thread finishes the other set's previous call, retires finished programs,
 * admits new ones and plans the next call.  Every program renders into ONE page-locked array
 * sized from its duration, which receives the device-to-host copies directly; a finished
 * program's array goes to the sink -- the caller's callback, or the WAV writer threads of
 * saugen_render_batch_wav -- and back to the pool.  No audio is computed here.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/saugen_b200.h"

namespace {

struct Live {
	size_t index = 0;
	saugen_Generator *gen = nullptr;
	int16_t *buf = nullptr;
	size_t cap = 0, pos = 0;           /* frames */
};

struct LiveSet {
	saugen_Batch *batch = nullptr;
	std::vector<Live> v;
	std::vector<saugen_Generator*> gens;
	std::vector<int16_t*> ptrs;
	std::vector<size_t> lens;
	std::vector<int> more;
	bool in_flight = false;
};

size_t capacity_frames(const sauabi_Program *p, uint32_t srate, size_t call_len) {
	const size_t frames = ((size_t) p->duration_ms * srate + 999) / 1000;
	return (frames / call_len + 2) * call_len;
}

struct Driver {
	const sauabi_Program *const *prgs;
	size_t n;
	uint32_t srate;
	const saugen_WaveTables *tables;
	saugen_BatchOptions opt;
	saugen_pcm_sink sink;
	void *user;
	size_t next = 0;
	std::string err;

	bool admit(LiveSet &s) {
		const int ch = opt.mono ? 1 : 2;
		while (s.v.size() < opt.group && next < n) {
			const size_t i = next++;
			saugen_Options go;
			memset(&go, 0, sizeof go);
			go.device = opt.device;
			go.max_call_len = opt.call_len;
			Live l;
			l.index = i;
			l.gen = saugen_create(prgs[i], srate, tables, &go);
			if (!l.gen) { err = std::string("saugen_create: ") + saugen_last_error(); return false; }
			l.cap = capacity_frames(prgs[i], srate, opt.call_len);
			l.buf = (int16_t*) saugen_pinned_alloc(l.cap * ch * sizeof(int16_t));
			if (!l.buf) { saugen_destroy(l.gen); err = "saugen_pinned_alloc failed"; return false; }
			s.v.push_back(l);
		}
		return true;
	}
	bool begin(LiveSet &s) {
		const int ch = opt.mono ? 1 : 2;
		const size_t m = s.v.size();
		s.gens.resize(m); s.ptrs.resize(m);
		for (size_t k = 0; k < m; ++k) {
			Live &l = s.v[k];
			if (l.pos + opt.call_len > l.cap) {            /* rare: longer than announced */
				const size_t ncap = l.cap + 4 * (size_t) opt.call_len;
				int16_t *nb = (int16_t*) saugen_pinned_alloc(ncap * ch * sizeof(int16_t));
				if (!nb) { err = "saugen_pinned_alloc failed"; return false; }
				memcpy(nb, l.buf, l.pos * ch * sizeof(int16_t));
				saugen_pinned_free(l.buf);
				l.buf = nb; l.cap = ncap;
			}
			s.gens[k] = l.gen;
			s.ptrs[k] = l.buf + l.pos * ch;
		}
		const int r = saugen_batch_begin(s.batch, s.gens.data(), m, s.ptrs.data(), opt.call_len, !opt.mono, 1);
		if (r < 0) { err = std::string("saugen_batch_begin: ") + saugen_last_error(); return false; }
		s.in_flight = true;
		return true;
	}
	bool end(LiveSet &s) {
		const size_t m = s.v.size();
		s.lens.assign(m, 0); s.more.assign(m, 0);
		const int r = saugen_batch_end(s.batch, s.lens.data(), s.more.data());
		s.in_flight = false;
		if (r < 0) { err = std::string("saugen_batch_end: ") + saugen_last_error(); return false; }
		size_t w = 0;
		for (size_t k = 0; k < m; ++k) {
			Live &l = s.v[k];
			l.pos += s.lens[k];
			if (s.more[k]) { if (w != k) s.v[w] = l; ++w; continue; }
			saugen_destroy(l.gen);
			sink(user, l.index, l.buf, l.pos, opt.mono ? 1 : 2);      /* the sink owns the array now */
		}
		s.v.resize(w);
		return true;
	}
	int run() {
		std::vector<LiveSet> sets(opt.depth);
		for (LiveSet &s : sets) {
			s.batch = saugen_batch_create(opt.device);
			if (!s.batch) { err = std::string("saugen_batch_create: ") + saugen_last_error(); break; }
		}
		bool ok = err.empty();
		for (size_t k = 0; ok; ++k) {
			LiveSet &s = sets[k % sets.size()];
			if (s.in_flight) ok = end(s);
			if (ok) ok = admit(s);
			if (ok && !s.v.empty()) ok = begin(s);
			if (!ok) break;
			bool any = next < n;
			for (LiveSet &x : sets) any = any || x.in_flight;
			if (!any) break;
		}
		for (LiveSet &s : sets) {
			if (s.in_flight) saugen_batch_end(s.batch, nullptr, nullptr);
			for (Live &l : s.v) { saugen_destroy(l.gen); saugen_pinned_free(l.buf); }
			if (s.batch) saugen_batch_destroy(s.batch);
		}
		return ok ? 0 : -1;
	}
};

/* ---- WAV writer threads (player/sndfile.c:63-109) ---------------------------- */

struct WavJob { size_t index; int16_t *pcm; size_t frames; int ch; };
struct WavWriter {
	const char *const *paths;
	uint32_t srate;
	std::mutex mu;
	std::condition_variable cv;
	std::deque<WavJob> q;
	bool done = false;
	int failed = 0;
	std::vector<std::thread> th;

	static void put32(unsigned char *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
	static void put16(unsigned char *p, uint32_t v) { p[0] = v; p[1] = v >> 8; }
	bool write(const WavJob &j) {
		const uint32_t bytes = (uint32_t) (j.frames * j.ch * 2);
		unsigned char h[44];
		memcpy(h, "RIFF", 4); put32(h + 4, 36 + bytes); memcpy(h + 8, "WAVEfmt ", 8);
		put32(h + 16, 16); put16(h + 20, 1); put16(h + 22, j.ch); put32(h + 24, srate);
		put32(h + 28, j.ch * srate * 2); put16(h + 32, j.ch * 2); put16(h + 34, 16);
		memcpy(h + 36, "data", 4); put32(h + 40, bytes);
		FILE *f = fopen(paths[j.index], "wb");
		if (!f) return false;
		bool ok = fwrite(h, 1, 44, f) == 44 && (bytes == 0 || fwrite(j.pcm, 1, bytes, f) == bytes);
		ok = fclose(f) == 0 && ok;
		return ok;
	}
	void work() {
		for (;;) {
			WavJob j;
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [this] { return done || !q.empty(); });
				if (q.empty()) return;
				j = q.front();
				q.pop_front();
			}
			const bool ok = write(j);
			saugen_pinned_free(j.pcm);
			if (!ok) { std::lock_guard<std::mutex> lk(mu); ++failed; }
		}
	}
	static void sink(void *user, size_t index, int16_t *pcm, size_t frames, int ch) {
		WavWriter *w = (WavWriter*) user;
		{
			std::lock_guard<std::mutex> lk(w->mu);
			w->q.push_back(WavJob{index, pcm, frames, ch});
		}
		w->cv.notify_one();
	}
};

void fill_defaults(saugen_BatchOptions &o, uint32_t srate, const saugen_BatchOptions *in) {
	memset(&o, 0, sizeof o);
	if (in) o = *in;
	if (!o.call_len) o.call_len = 4u * (uint32_t) (((uint64_t) 256 * srate) / 1000);      /* 4 x saugns.c:471 */
	o.call_len = (o.call_len + 3u) & ~3u;
	if (!o.group) o.group = 128;
	if (!o.depth) o.depth = 2;
	if (!o.io_threads) o.io_threads = 4;
}

thread_local std::string g_batch_err;

} // namespace

extern "C" int saugen_render_batch(const sauabi_Program *const *prgs, size_t n, uint32_t srate,
		const saugen_WaveTables *tables, const saugen_BatchOptions *opt, saugen_pcm_sink sink, void *user) {
	if (!prgs || !srate || !tables) { g_batch_err = "saugen_render_batch: NULL argument"; return -1; }
	Driver d;
	d.prgs = prgs; d.n = n; d.srate = srate; d.tables = tables; d.sink = sink; d.user = user;
	if (!sink)             /* no sink: the PCM reaches host memory and is dropped (throughput measurements) */
		d.sink = [](void*, size_t, int16_t *pcm, size_t, int) { saugen_pinned_free(pcm); };
	fill_defaults(d.opt, srate, opt);
	const int r = d.run();
	if (r < 0) g_batch_err = d.err;
	return r;
}

extern "C" int saugen_render_batch_wav(const sauabi_Program *const *prgs, size_t n, uint32_t srate,
		const saugen_WaveTables *tables, const saugen_BatchOptions *opt, const char *const *paths) {
	if (!paths) { g_batch_err = "saugen_render_batch_wav: NULL paths"; return -1; }
	saugen_BatchOptions o;
	fill_defaults(o, srate, opt);
	WavWriter w;
	w.paths = paths; w.srate = srate;
	for (uint32_t t = 0; t < o.io_threads; ++t) w.th.emplace_back(&WavWriter::work, &w);
	const int r = saugen_render_batch(prgs, n, srate, tables, &o, &WavWriter::sink, &w);
	{
		std::lock_guard<std::mutex> lk(w.mu);
		w.done = true;
	}
	w.cv.notify_all();
	for (std::thread &t : w.th) t.join();
	if (r == 0 && w.failed) { g_batch_err = "saugen_render_batch_wav: could not write every file"; return -1; }
	return r;
}

extern "C" const char *saugen_batch_last_error(void) { return g_batch_err.c_str(); }
