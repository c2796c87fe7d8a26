/* runtime.cpp -- host side of the B200 generator back end and its C ABI
 * (include/saugen_b200.h).
 *
 * create : flattens the pointer-rich sauProgram (sau/program.h:212-265) into
 *          index-based records, walks the event list once to mirror the
 *          modulator-graph topology (the mod-list pointer updates of
 *          update_op, sau/generator.c:316-339) and compiles, per event, the
 *          voice's operator walk (run_voice/run_block/mix_add,
 *          generator.c:448-788) into bytecode; uploads everything.
 * run    : replays sauGenerator_run's event/time interleaving
 *          (generator.c:915-949) as a list of segments for this call, then
 *          launches render_kernel + mix_kernel once.  All sample-rate work and
 *          all generator state stay on the device; the host never computes
 *          audio.  There is no CPU fallback.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <mutex>
#include <thread>
#include <atomic>
#include <chrono>

#include "../../include/saugen_b200.h"
#include "device_types.h"

namespace saugen {
size_t render_smem_bytes(uint32_t wave_mask, uint32_t nbufs, uint32_t nslots_ops, uint32_t nplan,
		uint32_t warps, uint32_t team);
cudaError_t launch_render(const CallDesc *d_calls, uint32_t ncalls, const SegDesc *d_segs,
		const UnitDesc *d_units, uint32_t ntasks, const float *d_tables, const double *d_coefs,
		uint32_t wave_mask, uint32_t nbufs, uint32_t nslots_ops, uint32_t nplan, uint32_t warps,
		uint32_t ticketed_ctas, uint32_t sched_mode, uint32_t team, cudaStream_t stream, uint32_t multi = 1);
int render_ctas_per_sm(size_t smem, uint32_t warps);
uint32_t plan_area_bytes(uint32_t nplan);
size_t coef_table_bytes();
cudaError_t launch_coefs(const float *d_tables, double *d_coefs, uint32_t *d_inexact, cudaStream_t stream);
cudaError_t launch_mix(const CallDesc *d_calls, uint32_t ncalls, const SegDesc *d_segs,
		uint32_t max_call_len, uint32_t mode, cudaStream_t stream);
cudaError_t launch_prologue(const InlineCall &ic, const PrologueArgs &a, cudaStream_t stream);
cudaError_t launch_planes_to_pcm(const float *d_mix, uint32_t plane_stride, uint32_t n,
		uint32_t stereo, int16_t *d_pcm, cudaStream_t stream);
cudaError_t launch_selftest(const float *d_tables, unsigned long long *d_bad, cudaStream_t stream);
int device_sm_count();
size_t device_smem_optin();
cudaError_t read_signature_dump(uint32_t out[36]);
cudaError_t read_team_dump(uint32_t out[32]);
}

using namespace saugen;

static thread_local std::string g_err;
static void set_err(const char *what, cudaError_t e) {
	char b[256];
	snprintf(b, sizeof b, "%s: %s", what, e == cudaSuccess ? "failed" : cudaGetErrorString(e));
	g_err = b;
	fprintf(stderr, "saugen_b200: error: %s\n", b);
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(#call, e_); goto fail; } } while (0)

static uint32_t ms_in_samples(uint64_t ms, uint64_t srate, int *carry) {   /* sau/math.h:35-46 */
	uint64_t t = ms * srate;
	if (carry) { t += *carry; *carry = (int) (t % 1000); }
	return (uint32_t) (t / 1000);
}

/* ---- device wave-table blocks, shared by content ------------------------ */

struct TableBlock { float *d; double *coefs; };
static std::mutex g_tab_mu;
static std::map<std::pair<int, uint64_t>, TableBlock> g_tabs;

/* A batch render creates thousands of generators on the same table set: the upload
 * image and its byte-wise hash are skipped when the caller's struct address and a
 * word-wise hash of EVERY table value and coefficient were seen before on this device
 * (tables edited in place at the same address get a new device copy). */
struct TableSeen { int device; const void *addr; uint64_t fp; TableBlock blk; };
static std::vector<TableSeen> g_tab_seen;
/* every table value and coefficient: four independent multiply-xor chains over 64-bit words
 * (~10 us for the 96 KiB; a batch creates thousands of generators on the same tables) */
static uint64_t table_fingerprint(const saugen_WaveTables *t) {
	uint64_t h[4] = {1469598103934665603ull, 0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull};
	for (int w = 0; w < NUM_WAVES; ++w) {
		const unsigned char *p = (const unsigned char*) t->pilut[w];
		for (int i = 0; i < WAVE_LEN * 4; i += 32) {
			uint64_t v[4];
			memcpy(v, p + i, 32);
			for (int k = 0; k < 4; ++k) h[k] = (h[k] ^ v[k]) * 0x9E3779B97F4A7C15ull + (h[k] >> 31);
		}
		uint32_t a, b; memcpy(&a, &t->amp_scale[w], 4); memcpy(&b, &t->amp_dc[w], 4);
		h[w & 3] = (h[w & 3] ^ (((uint64_t) a << 32) | b)) * 0x9E3779B97F4A7C15ull;
		h[(w + 1) & 3] ^= (uint32_t) t->phase_adj[w] * 0x85EBCA6Bull;
	}
	uint64_t r = 0;
	for (int k = 0; k < 4; ++k) { r = (r ^ h[k]) * 0xFF51AFD7ED558CCDull; r ^= r >> 33; }
	return r;
}

static float *get_device_tables(int device, const saugen_WaveTables *t, double **coefs_out = nullptr) {
	const uint64_t fp = table_fingerprint(t);
	{
		std::lock_guard<std::mutex> lk(g_tab_mu);
		for (const TableSeen &e : g_tab_seen)
			if (e.device == device && e.addr == (const void*) t && e.fp == fp) {
				if (coefs_out) *coefs_out = e.blk.coefs;
				return e.blk.d;
			}
	}
	std::vector<float> host((size_t) NUM_WAVES * WAVE_LEN + sizeof(WaveCoeffs) / sizeof(float));
	for (int w = 0; w < NUM_WAVES; ++w)
		memcpy(&host[(size_t) w * WAVE_LEN], t->pilut[w], sizeof(float) * WAVE_LEN);
	WaveCoeffs wc;
	for (int w = 0; w < NUM_WAVES; ++w) {
		wc.diff_scale[w] = t->amp_scale[w] * 0.125f * 4294967296.f;   /* wave.h:144-145 */
		wc.diff_offset[w] = t->amp_dc[w];
		wc.amp256[w] = t->amp_scale[w] * 256.f;
		wc.phase_adj[w] = t->phase_adj[w];
	}
	memcpy(&host[(size_t) NUM_WAVES * WAVE_LEN], &wc, sizeof(wc));
	uint64_t h = 1469598103934665603ull;
	const unsigned char *b = (const unsigned char*) host.data();
	for (size_t i = 0; i < host.size() * sizeof(float); ++i) { h ^= b[i]; h *= 1099511628211ull; }
	std::lock_guard<std::mutex> lk(g_tab_mu);
	auto key = std::make_pair(device, h);
	auto it = g_tabs.find(key);
	if (it != g_tabs.end()) {
		if (coefs_out) *coefs_out = it->second.coefs;
		if (g_tab_seen.size() < 64) g_tab_seen.push_back(TableSeen{device, (const void*) t, fp, it->second});
		return it->second.d;
	}
	float *d = nullptr;
	double *dc = nullptr;
	if (cudaMalloc(&d, host.size() * sizeof(float)) != cudaSuccess) return nullptr;
	if (cudaMemcpy(d, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
	    cudaMalloc(&dc, coef_table_bytes()) != cudaSuccess) {
		cudaFree(d);
		return nullptr;
	}
	/* per-index cubic coefficients, computed once on the device from the uploaded tables;
	 * a table set whose c1 values are not exact floats gets no planes (never seen) */
	{
		uint32_t *d_flag = nullptr, h_flag = 1;
		bool ok = cudaMalloc(&d_flag, sizeof(uint32_t)) == cudaSuccess &&
			cudaMemset(d_flag, 0, sizeof(uint32_t)) == cudaSuccess &&
			launch_coefs(d, dc, d_flag, 0) == cudaSuccess &&
			cudaMemcpy(&h_flag, d_flag, sizeof h_flag, cudaMemcpyDeviceToHost) == cudaSuccess;
		if (d_flag) cudaFree(d_flag);
		if (!ok) { cudaFree(d); cudaFree(dc); return nullptr; }
		if (h_flag) { cudaFree(dc); dc = nullptr; }
	}
	g_tabs[key] = TableBlock{d, dc};
	if (g_tab_seen.size() < 64) g_tab_seen.push_back(TableSeen{device, (const void*) t, fp, TableBlock{d, dc}});
	if (coefs_out) *coefs_out = dc;
	return d;
}

/* ---- cached device / pinned-host memory --------------------------------- *
 * cudaFree and cudaFreeHost synchronise the whole device and take up to
 * hundreds of milliseconds for the large carrier-row blocks; a batch render
 * creates and destroys thousands of generators.  Blocks therefore go back to a
 * per-device free list by size class and are handed out again (every user
 * synchronises its stream before releasing, so a recycled block is idle). */
namespace {
struct MemPool {
	static const int MAXDEV = 64;
	std::mutex mu;
	std::multimap<size_t, void*> free_[2][MAXDEV];     /* [0] device memory, [1] pinned host */
	std::map<void*, size_t> live;
	std::set<void*> slab_piece;                        /* carved out of a slab: recycled only */
	size_t cached[2][MAXDEV] = {{0}};
	static size_t round_size(size_t n) {
		if (n < 4096) return 4096;
		size_t p = 4096;
		while (p * 2 <= n) p *= 2;
		/* up to 32 MiB: powers of two, so that the row blocks of a batch of small
		 * scripts (a few MiB each, every size different) fall into a handful of
		 * classes and recycle; above: <= 12.5 % slack */
		if (n <= ((size_t) 32 << 20)) return p == n ? n : p * 2;
		const size_t step = p / 8;
		return (n + step - 1) / step * step;
	}
	void *alloc(bool host, int dev, size_t bytes) {
		const size_t r = round_size(bytes);
		const int d = host ? 0 : (dev < 0 || dev >= MAXDEV ? 0 : dev);
		{
			std::lock_guard<std::mutex> lk(mu);
			auto it = free_[host][d].find(r);
			if (it != free_[host][d].end()) {
				void *p = it->second;
				free_[host][d].erase(it);
				cached[host][d] -= r;
				live[p] = r;
				return p;
			}
		}
		/* small classes come out of slabs: one cudaMalloc / cudaHostAlloc per SLAB bytes
		 * instead of one per generator (pinned allocations cost about a millisecond) */
		const size_t SLAB = host ? (size_t) 32 << 20 : (size_t) 128 << 20;
		const size_t want = r <= SLAB / 4 ? SLAB / r * r : r;
		void *p = nullptr;
		cudaError_t e = host ? cudaHostAlloc(&p, want, cudaHostAllocPortable) : cudaMalloc(&p, want);
		if (e != cudaSuccess) {                        /* give the cache back and retry once */
			cudaGetLastError();
			trim(host, d);
			e = host ? cudaHostAlloc(&p, want, cudaHostAllocPortable) : cudaMalloc(&p, want);
			if (e != cudaSuccess) return nullptr;
		}
		std::lock_guard<std::mutex> lk(mu);
		live[p] = r;
		if (want > r) {
			/* the rest of the slab goes straight to the free list; slab pieces are
			 * never handed back to CUDA one by one (trim skips them) */
			for (size_t off = r; off + r <= want; off += r) {
				void *q = (unsigned char*) p + off;
				free_[host][d].insert(std::make_pair(r, q));
				cached[host][d] += r;
				slab_piece.insert(q);
			}
			slab_piece.insert(p);
		}
		return p;
	}
	void release(bool host, int dev, void *p) {
		if (!p) return;
		const int d = host ? 0 : (dev < 0 || dev >= MAXDEV ? 0 : dev);
		const size_t limit = host ? ((size_t) 1 << 30) : ((size_t) 24 << 30);
		size_t r = 0;
		{
			std::lock_guard<std::mutex> lk(mu);
			auto it = live.find(p);
			if (it != live.end()) { r = it->second; live.erase(it); }
			if (r && (cached[host][d] + r <= limit || slab_piece.count(p))) {
				free_[host][d].insert(std::make_pair(r, p));
				cached[host][d] += r;
				return;
			}
		}
		if (host) cudaFreeHost(p); else cudaFree(p);
	}
	void trim(bool host, int d) {
		std::vector<void*> v;
		{
			std::lock_guard<std::mutex> lk(mu);
			std::multimap<size_t, void*> keep;
			size_t kept = 0;
			for (auto &kv : free_[host][d]) {
				if (slab_piece.count(kv.second)) { keep.insert(kv); kept += kv.first; }
				else v.push_back(kv.second);
			}
			free_[host][d].swap(keep);
			cached[host][d] = kept;
		}
		for (void *p : v) { if (host) cudaFreeHost(p); else cudaFree(p); }
	}
};
MemPool g_pool;

/* cudaStreamCreate / cudaStreamDestroy cost a few hundred microseconds and
 * serialise with other threads' launches: generators that own their stream take
 * it from a per-device list and give it back (synchronised) at destroy. */
struct StreamPool {
	std::mutex mu;
	std::vector<cudaStream_t> free_[MemPool::MAXDEV];
	cudaStream_t get(int dev) {
		if (dev >= 0 && dev < MemPool::MAXDEV) {
			std::lock_guard<std::mutex> lk(mu);
			if (!free_[dev].empty()) { cudaStream_t s = free_[dev].back(); free_[dev].pop_back(); return s; }
		}
		cudaStream_t s = nullptr;
		if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
		return s;
	}
	void put(int dev, cudaStream_t s) {
		if (dev >= 0 && dev < MemPool::MAXDEV) {
			std::lock_guard<std::mutex> lk(mu);
			if (free_[dev].size() < 1024) { free_[dev].push_back(s); return; }
		}
		cudaStreamDestroy(s);
	}
};
StreamPool g_streams;

/* carve 256-byte aligned pieces out of one block */
struct Carver {
	size_t off = 0;
	size_t take(size_t bytes) { size_t o = off; off = (off + (bytes ? bytes : 1) + 255) & ~(size_t) 255; return o; }
};
}

/* developer aid: SAUGEN_PROFILE=1 prints where saugen_create spends its time */
namespace {
struct CreateProf {
	std::atomic<long long> ns[6];
	std::atomic<long long> n;
	bool on;
	CreateProf() : on(getenv("SAUGEN_PROFILE") != nullptr) { for (auto &x : ns) x = 0; n = 0; }
	~CreateProf() {
		if (!on || !n) return;
		static const char *nm[6] = {"flatten", "stream+tables", "device alloc", "pinned alloc", "events", "upload+sync"};
		fprintf(stderr, "saugen_create x %lld:", (long long) n);
		for (int i = 0; i < 6; ++i) fprintf(stderr, " %s %.3f ms", nm[i], ns[i] / 1e6 / (double) n);
		fprintf(stderr, "\n");
	}
};
CreateProf g_cprof;
inline long long now_ns() {
	return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}

/* ---- generator object ---------------------------------------------------- */

struct HostOp {
	bool inited = false;
	uint8_t type = 0;
	bool line_set[LINE_COUNT] = {false, false, false, false, false, false};  /* ever given by an event */
	const sauabi_ProgramIDArr *mods[SAUABI_POP_NAMED] = {0};   /* index = use type */
};

struct saugen_Generator {
	const sauabi_Program *prg = nullptr;
	uint32_t srate = 0;
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	cudaStream_t aux_stream = nullptr;    /* saugen_mix_to_pcm */
	cudaStream_t copy_stream = nullptr;   /* read-backs (and saugen_mix_to_pcm): off the kernels' stream, so that
	                                       * the next call's launches never queue behind a device-to-host copy */
	uint32_t vo_count = 0, op_count = 0, nlv = 0;
	uint32_t voice_begin = 0, voice_end = 0;
	uint32_t row_stride = 0;
	uint32_t row_len = 0, nbufs = 1, max_ops = 1, wave_mask = 0, seg_cap = 0, sched = 0;
	bool big_endian = false;           // saugen_Options::pcm_big_endian
	uint32_t nplan = 0;                // block-plan records the largest voice program needs (render_plan.cuh)
	float amp_scale = 0.f;
	/* timeline (host-only integer bookkeeping) */
	std::vector<uint64_t> ev_time;     // absolute sample time of each event
	std::vector<uint32_t> ev_handover; // Flat::ev_handover
	std::vector<uint32_t> group_first; // first segment of every render launch after the first, of the planned call
	size_t next_event = 0;
	uint64_t cur_time = 0;
	bool ended = false;
	/* device */
	GenDesc h_desc;
	GenDesc *d_desc = nullptr;
	float *d_tap = nullptr;           /* saugen_debug_tap */
	void *d_plan_cache = nullptr;     /* GenDesc::plan_cache */
	bool team_cache_failed = false;
	float *d_tables = nullptr;
	double *d_coefs = nullptr;
	bool ctab_ok = false;              // every voice program is fast-path material (see create)
	void *d_ops = nullptr, *d_voices = nullptr, *d_events = nullptr, *d_opdata = nullptr,
	     *d_code = nullptr, *d_prog_ops = nullptr, *d_vev_off = nullptr, *d_vev_idx = nullptr;
	float *d_rows_s = nullptr, *d_rows_r = nullptr, *d_mix = nullptr;
	VoiceSeg *d_vlen = nullptr;
	uint32_t *d_status = nullptr, *d_progress = nullptr;
	UnitDesc *d_units = nullptr, *h_units = nullptr;
	uint32_t unit_cap = 0;
	std::vector<UnitDesc> units_tmp;
	int16_t *d_pcm = nullptr;
	CallDesc *d_call = nullptr;
	SegDesc *d_segs = nullptr;
	/* pinned host staging */
	uint32_t *h_status = nullptr;
	int16_t *h_pcm = nullptr;
	CallDesc *h_call = nullptr;
	SegDesc *h_segs = nullptr;
	uint64_t counters[4] = {0, 0, 0, 0};
	std::vector<SegDesc> segs_tmp;
	/* One call in flight or just completed.  Two alternate: while the caller consumes call k,
	 * call k + 1 may already be rendering (run-ahead, see saugen_run). */
	struct CallSlot {
		bool in_flight = false, timed = false;
		size_t buf_len = 0, host_bytes = 0, gen_base = 0;
		int stereo = 0;
		uint32_t mode = 0, nseg = 0;
		cudaEvent_t done = nullptr;        /* the call's read-back has arrived (copy stream) */
		cudaEvent_t mixed = nullptr;       /* its kernels are through (launch stream) */
		cudaEvent_t ev_t[3] = {nullptr, nullptr, nullptr};
		uint32_t *h_status = nullptr, *d_status = nullptr;
		int16_t *h_pcm = nullptr, *d_pcm = nullptr;
		float *d_mix = nullptr;            /* float L / R planes (saugen_run_mix) + MIX_TAIL floats for the caller */
		CallDesc *h_call = nullptr;        /* pinned staging of the call's descriptors ([call][segs][units]) */
		SegDesc *h_segs = nullptr;
		UnitDesc *h_units = nullptr;
		size_t ev_after = 0;               /* next_event once this call was planned */
	} slot[2];
	int cur_slot = 0;                  /* the slot of the call returned last */
	/* run-ahead: the next call, launched with the parameters of the last one before the caller
	 * asks for it.  The operator / voice state and the host timeline of before it are kept
	 * (d_snap, spec_*) so that it can be undone when the caller asks for something else. */
	bool spec_valid = false;
	int spec_slot = 0;
	size_t spec_next_event = 0;
	uint64_t spec_cur_time = 0;
	void *d_snap = nullptr;
	size_t state_bytes = 0;
	int streak = 0;                    /* consecutive calls with the same parameters, no inspection between */
	bool rows_stale = false;           /* the carrier rows hold a run-ahead call's, not the last returned call's */
	/* device time of the two kernels, measured with events on the launch stream */
	double render_ms = 0.0, mix_ms = 0.0;
	bool timing = false;
	size_t zero_bytes = 0, zero_bytes0 = 0, back_bytes_fixed = 0, units_off_in_call = 0;
	bool compact = true;               /* [vlen..status] and [status][pcm] still adjacent (no growth yet) */
	/* every block this generator took from the pool: (pointer, is pinned host) */
	std::vector<std::pair<void*, bool>> blocks;
	void *take(bool host, size_t bytes) {
		void *p = g_pool.alloc(host, device, bytes);
		if (p) blocks.push_back(std::make_pair(p, host));
		return p;
	}
};

/* ---- bytecode compiler --------------------------------------------------- */

struct Compiler {
	const std::vector<HostOp> &ops;
	std::vector<Instr> out;
	std::vector<uint32_t> prog_ops;       // slot -> operator id of the program being built
	std::vector<int32_t> slot_of;         // operator id -> slot, -1 if not in the program
	std::vector<char> onstack;
	uint32_t max_buf = 0;
	bool too_deep = false;
	int depth = 0;
	Compiler(const std::vector<HostOp> &o) : ops(o), slot_of(o.size(), -1), onstack(o.size(), 0) {}

	uint32_t slot(uint32_t op) {
		if (slot_of[op] < 0) { slot_of[op] = (int32_t) prog_ops.size(); prog_ops.push_back(op); }
		return (uint32_t) slot_of[op];
	}
	/* which fields of an instruction name work buffers: bit 0..4 = a..e, bit 5 =
	 * the buffer after b too (self-PM scratch) */
	std::vector<uint8_t> bufmask;
	static uint8_t buf_fields(uint8_t opc, uint32_t d, uint16_t flags) {
		enum { A = 1, B = 2, Cc = 4, D = 8, E = 16, B1 = 32 };
		switch (opc) {
		case I_LINE: return (uint8_t) ((d ? A : 0) | B);
		case I_VPAN: return A;
		case I_WLEAF: return (uint8_t) (A | E | ((flags & F_MAY_SELFMOD) ? (B | B1) : 0) |
				((flags & F_KEEP_FREQ) ? B : 0));
		case I_WHEAD: return A | B | E;
		case I_WTAIL: return (uint8_t) (A | B | Cc | D | ((flags & F_MAY_SELFMOD) ? B1 : 0));
		case I_END: return 0;
		default: return A | B | Cc | D | E;
		}
	}
	void emit(uint8_t opc, uint32_t op_slot, uint32_t a, uint32_t b = NO_BUF, uint32_t c = NO_BUF,
			uint32_t d = NO_BUF, uint32_t e = NO_BUF, uint16_t flags = 0) {
		Instr i;
		i.opcode = opc; i.a = (uint8_t) a; i.b = (uint8_t) b; i.c = (uint8_t) c;
		i.d = (uint8_t) d; i.e = (uint8_t) e; i.flags = flags; i.op = op_slot; i.aux = 0;
		if (a != NO_BUF && a >= 250) too_deep = true;
		if (b != NO_BUF && b >= 249) too_deep = true;
		if (c != NO_BUF && c >= 250) too_deep = true;
		if (e != NO_BUF && e >= 250) too_deep = true;
		out.push_back(i);
		bufmask.push_back(buf_fields(opc, d, flags));
	}
	/* Renumber the work buffers a finished voice program really uses to 0..n-1,
	 * keeping their order (so "b and the buffer after it" stays adjacent): the
	 * recursive numbering (base + k per nesting level) leaves gaps, and every
	 * buffer costs each warp 1 KiB of shared memory. */
	void compact_buffers() {
		bool used[256] = {false};
		for (size_t k = 0; k < out.size(); ++k) {
			const Instr &i = out[k];
			const uint8_t m = bufmask[k];
			const uint8_t f[5] = {i.a, i.b, i.c, i.d, i.e};
			for (int q = 0; q < 5; ++q) if ((m >> q & 1) && f[q] != NO_BUF) used[f[q]] = true;
			if ((m & 32) && i.b != NO_BUF) used[i.b + 1] = true;
		}
		uint8_t map[256];
		uint32_t n = 0;
		for (int b = 0; b < 255; ++b) map[b] = used[b] ? (uint8_t) n++ : (uint8_t) NO_BUF;
		map[NO_BUF] = NO_BUF;
		for (size_t k = 0; k < out.size(); ++k) {
			Instr &i = out[k];
			const uint8_t m = bufmask[k];
			if (m & 1) i.a = map[i.a];
			if (m & 2) i.b = map[i.b];
			if (m & 4) i.c = map[i.c];
			if (m & 8) i.d = map[i.d];
			if (m & 16) i.e = map[i.e];
		}
		if (n > max_buf) max_buf = n;
	}
	static uint32_t cnt(const sauabi_ProgramIDArr *a) { return a ? a->count : 0; }

	/* run_param_with_rangemod, generator.c:448-477 */
	void param(uint32_t op, uint32_t B, int par, int rpar, int mods_use, int rmods_use,
			uint32_t mulbuf, uint32_t reused_freq, bool is_freq) {
		const HostOp &n = ops[op];
		const uint32_t freq = reused_freq != NO_BUF ? reused_freq : (is_freq ? B : NO_BUF);
		emit(I_LINE, slot(op), B, mulbuf, par, 1);
		const sauabi_ProgramIDArr *rm = n.mods[rmods_use], *m = n.mods[mods_use];
		if (cnt(rm) > 0) {
			emit(I_LINE, slot(op), B + 1, mulbuf, rpar, 1);
			for (uint32_t i = 0; i < rm->count; ++i)
				visit(rm->ids[i], B + 2, freq, true, i > 0 ? F_LAYER : 0);
			emit(I_RANGE, 0, B, B + 1, B + 2);
		} else if (n.line_set[rpar]) {
			/* a never-set line has no state to advance: its sauLine_skip is a no-op */
			emit(I_LINE, slot(op), 0, NO_BUF, rpar, 0);
		}
		for (uint32_t i = 0; i < cnt(m); ++i)
			visit(m->ids[i], B, freq, false, F_LAYER);
	}

	/* run_block + run_block_{amp,noiseg,wosc,rasg}, generator.c:505-729 */
	void visit(uint32_t op, uint32_t base, uint32_t parent_freq, bool wave_env, uint16_t layer_flags) {
		if (op >= ops.size() || !ops[op].inited) {     /* never prepared: renders nothing */
			if (!(layer_flags & (F_LAYER | F_LAYER_PMA))) emit(I_ZERO, 0, base);
			return;
		}
		if (onstack[op]) { emit(I_ZERO, 0, base); return; }   /* generator.c:685-689 */
		if (++depth >= MAX_NEST - 1) { too_deep = true; --depth; return; }
		onstack[op] = 1;
		const HostOp &n = ops[op];
		const uint32_t sl = slot(op);
		const uint16_t mixf = wave_env ? F_WAVEENV : 0;
		if (n.type == SAUABI_POPT_wave) {
			/* fused forms of run_block_wosc where the parameter lists allow */
			const sauabi_ProgramIDArr *pl = n.mods[SAUABI_POP_pmod], *fl = n.mods[SAUABI_POP_fpmod],
				*al = n.mods[SAUABI_POP_apmod];
			const bool simple_freq = !cnt(n.mods[SAUABI_POP_fmod]) && !cnt(n.mods[SAUABI_POP_rfmod]);
			const bool simple_amp = !cnt(n.mods[SAUABI_POP_amod]) && !cnt(n.mods[SAUABI_POP_ramod]) &&
				!cnt(al);
			const bool kids = cnt(pl) || cnt(fl);
			const uint32_t phase = base + 1, freq = base + 2;
			uint16_t wf = layer_flags | mixf;
			if (n.line_set[LINE_FREQ2]) wf |= F_SKIP_FREQ2;
			if (n.line_set[LINE_AMP2]) wf |= F_SKIP_AMP2;
			if (n.line_set[LINE_PMA]) wf |= F_MAY_SELFMOD;
			/* the carrier's pan modulators take its frequency buffer as their
			 * parent frequency (mix_add, generator.c:764-766) */
			if (depth == 1 && cnt(n.mods[SAUABI_POP_camod])) wf |= F_KEEP_FREQ;
			if (simple_freq && simple_amp && !kids) {
				emit(I_WLEAF, sl, base, freq, NO_BUF, NO_BUF, parent_freq, wf);
				onstack[op] = 0; --depth;
				return;
			}
			const size_t enter_at = out.size();
			if (simple_freq) {
				emit(I_WHEAD, sl, base, freq, NO_BUF, NO_BUF, parent_freq, wf);
			} else {
				emit(I_ENTER, sl, base, NO_BUF, NO_BUF, NO_BUF, NO_BUF, layer_flags);
				param(op, freq, LINE_FREQ, LINE_FREQ2, SAUABI_POP_fmod, SAUABI_POP_rfmod,
						parent_freq, NO_BUF, true);
			}
			uint32_t pm = NO_BUF, fpm = NO_BUF;
			for (uint32_t i = 0; i < cnt(pl); ++i) visit(pl->ids[i], base + 3, freq, false, i > 0 ? F_LAYER : 0);
			if (cnt(pl)) pm = base + 3;
			for (uint32_t i = 0; i < cnt(fl); ++i) visit(fl->ids[i], base + 4, freq, false, i > 0 ? F_LAYER : 0);
			if (cnt(fl)) fpm = base + 4;
			if (simple_amp) {
				out[enter_at].aux = (uint32_t) out.size();
				emit(I_WTAIL, sl, base, freq, pm, fpm, NO_BUF, wf);
			} else {
				emit(I_PHASOR, sl, phase, freq, pm, fpm);
				param(op, base + 3, LINE_AMP, LINE_AMP2, SAUABI_POP_amod, SAUABI_POP_ramod,
						NO_BUF, freq, false);
				emit(I_PMA, sl, base + 5);
				for (uint32_t i = 0; i < cnt(al); ++i)
					visit(al->ids[i], base + 5, freq, false, i > 0 ? F_LAYER : F_LAYER_PMA);
				emit(I_WOSC, sl, base + 4, phase, base + 5, NO_BUF, NO_BUF, cnt(al) ? F_HAS_APMODS : 0);
				emit(I_MIX, 0, base, base + 4, base + 3, NO_BUF, NO_BUF, mixf);
				out[enter_at].aux = (uint32_t) out.size();
				emit(I_LEAVE, sl, base, NO_BUF, NO_BUF, NO_BUF, NO_BUF, layer_flags);
			}
			onstack[op] = 0; --depth;
			return;
		}
		const size_t enter_at = out.size();
		emit(I_ENTER, sl, base, NO_BUF, NO_BUF, NO_BUF, NO_BUF, layer_flags);
		switch (n.type) {
		case SAUABI_POPT_amp:
		case SAUABI_POPT_noise: {
			param(op, base + 1, LINE_AMP, LINE_AMP2, SAUABI_POP_amod, SAUABI_POP_ramod,
					NO_BUF, NO_BUF, false);
			if (n.type == SAUABI_POPT_noise) {
				emit(I_NOISE, sl, base + 2);
				emit(I_MIX, 0, base, base + 2, base + 1, NO_BUF, NO_BUF, mixf);
			} else {
				emit(I_MIX, 0, base, NO_BUF, base + 1, NO_BUF, NO_BUF, mixf);
			}
			break; }
		case SAUABI_POPT_raseg: {
			const uint32_t cycle = base + 1, rasg = base + 2, freq = base + 3;
			param(op, freq, LINE_FREQ, LINE_FREQ2, SAUABI_POP_fmod, SAUABI_POP_rfmod,
					parent_freq, NO_BUF, true);
			uint32_t pm = NO_BUF, fpm = NO_BUF;
			const sauabi_ProgramIDArr *pl = n.mods[SAUABI_POP_pmod], *fl = n.mods[SAUABI_POP_fpmod],
				*al = n.mods[SAUABI_POP_apmod];
			for (uint32_t i = 0; i < cnt(pl); ++i) visit(pl->ids[i], base + 4, freq, false, i > 0 ? F_LAYER : 0);
			if (cnt(pl)) pm = base + 4;
			for (uint32_t i = 0; i < cnt(fl); ++i) visit(fl->ids[i], base + 5, freq, false, i > 0 ? F_LAYER : 0);
			if (cnt(fl)) fpm = base + 5;
			emit(I_CYCLOR, sl, cycle, rasg, freq, pm, fpm);
			param(op, base + 4, LINE_AMP, LINE_AMP2, SAUABI_POP_amod, SAUABI_POP_ramod,
					NO_BUF, freq, false);
			emit(I_PMA, sl, base + 5);
			for (uint32_t i = 0; i < cnt(al); ++i)
				visit(al->ids[i], base + 5, freq, false, i > 0 ? F_LAYER : F_LAYER_PMA);
			emit(I_RASG, sl, rasg, cycle, base + 5, NO_BUF, NO_BUF, cnt(al) ? F_HAS_APMODS : 0);
			emit(I_MIX, 0, base, rasg, base + 4, NO_BUF, NO_BUF, mixf);
			break; }
		}
		out[enter_at].aux = (uint32_t) out.size();   /* index of the LEAVE */
		emit(I_LEAVE, sl, base, NO_BUF, NO_BUF, NO_BUF, NO_BUF, layer_flags);
		onstack[op] = 0;
		--depth;
	}

	/* run_voice + mix_add, generator.c:749-788,833-846 */
	void voice(uint32_t carr) {
		out.clear();
		bufmask.clear();
		for (uint32_t op : prog_ops) slot_of[op] = -1;
		prog_ops.clear();
		if (carr >= ops.size() || !ops[carr].inited) { emit(I_END, 0, 0); return; }
		const uint32_t cs = slot(carr);   /* carrier is slot 0 */
		visit(carr, 0, NO_BUF, false, 0);
		const HostOp &n = ops[carr];
		const uint32_t fb = n.type == SAUABI_POPT_wave ? 2 : n.type == SAUABI_POPT_raseg ? 3 : 0;
		const sauabi_ProgramIDArr *cl = n.mods[SAUABI_POP_camod];
		emit(I_VPAN, cs, 1 + fb, NO_BUF, NO_BUF, cnt(cl) ? 1 : 0);
		for (uint32_t i = 0; i < cnt(cl); ++i)
			visit(cl->ids[i], 1 + fb, fb ? fb : NO_BUF, false, F_LAYER);
		emit(I_VOUT, cs, 0, 1 + fb);
		emit(I_END, 0, 0);
		compact_buffers();
	}
};

static void flatten_line(LineDelta *d, const sauabi_Line *s, uint32_t srate) {
	memset(d, 0, sizeof(*d));
	if (!s) return;
	d->present = 1;
	d->v0 = s->v0; d->vt = s->vt;
	d->end_samples = ms_in_samples(s->time_ms, srate, NULL);   /* line.c:325 */
	d->type = s->type; d->flags = s->flags;
}

extern "C" void saugen_destroy(saugen_Generator *o);

/* ---- the flat program: everything saugen_create derives from a sauProgram ----- *
 * (the device-ready, index-based "instruction form").
 * Events and op-data with pointers turned into indices and ms into samples, the
 * per-voice bytecode, the event timeline.  It does not need the sauProgram any
 * more, serialises to one relocatable blob (saugen_flatten) and instantiates on
 * any device (saugen_create_flat): parse and flatten once, render anywhere. */
struct Flat {
	uint32_t srate = 0, vo_count = 0, op_count = 0;
	uint32_t nbufs = 1, max_ops = 1, nplan = 0, wave_mask = 0;
	float amp_scale = 0.f;
	std::vector<uint64_t> ev_time;
	std::vector<EventRec> events;
	std::vector<OpDataRec> opdata;
	std::vector<Instr> code;
	std::vector<uint32_t> prog_ops, vev_off, vev_idx;
	/* Hand-over events.  Each voice's warp applies its own events and renders on its own,
	 * which equals the reference's global event order (generator.c:915-949) as long as an
	 * operator stays with one voice.  parseconv re-homes an operator whose voice slot was
	 * recycled (a later `@label` update gives the same op id a new voice): the event that
	 * touches an operator last touched under ANOTHER voice is a hand-over.  ev_handover[e]
	 * = that other voice + 1 (0 = none); a call is cut into separate render launches there
	 * (plan_call), so the operator's state passes through a kernel boundary. */
	std::vector<uint32_t> ev_handover;
};

static bool flatten_program(const sauabi_Program *prg, uint32_t srate, Flat &f) {
	Flat *o = &f;
	f.srate = srate; f.vo_count = prg->vo_count; f.op_count = prg->op_count;
	std::vector<EventRec> &events = f.events;
	std::vector<OpDataRec> &opdata = f.opdata;
	std::vector<Instr> &code = f.code;
	std::vector<uint32_t> &prog_ops = f.prog_ops, &vev_off = f.vev_off, &vev_idx = f.vev_idx;
	events.assign(prg->ev_count, EventRec());
	f.ev_handover.assign(prg->ev_count, 0u);
	{
		/* (one allocation each instead of a chain of doublings: a 4096-voice script has 12 288 op-data records) */
		size_t nod = 0;
		for (size_t i = 0; i < prg->ev_count; ++i) nod += prg->events[i].op_data_count;
		opdata.reserve(nod);
		code.reserve(4 * nod + 4 * prg->ev_count);
		prog_ops.reserve(nod + prg->ev_count);
	}
	std::vector<uint32_t> op_voice(prg->op_count, 0xffffffffu);   /* voice an operator was last touched under */
	std::vector<HostOp> hops(prg->op_count);
	std::vector<std::vector<uint32_t>> vev(prg->vo_count);
	std::vector<uint32_t> vcarr(prg->vo_count, 0xffffffffu);
	std::vector<std::pair<uint32_t, uint32_t>> vprog(prg->vo_count, {0u, 0u});
	std::vector<std::pair<uint32_t, uint32_t>> vops(prg->vo_count, {0u, 0u});
	o->amp_scale = 0.5f * prg->ampmult;                           /* generator.c:183-185 */
	if (prg->mode & SAUABI_PMODE_AMP_DIV_VOICES) o->amp_scale /= (float) prg->vo_count;

	/* event timeline with carry (generator.c:181-192) */
	{
		int carry = 0;
		uint64_t t = 0;
		o->ev_time.resize(prg->ev_count);
		for (size_t i = 0; i < prg->ev_count; ++i) {
			t += ms_in_samples(prg->events[i].wait_ms, srate, &carry);
			o->ev_time[i] = t;
		}
	}
	/* flatten events + mirror the graph + compile voice programs */
	{
		Compiler comp(hops);
		for (size_t ei = 0; ei < prg->ev_count; ++ei) {
			const sauabi_ProgramEvent *pe = &prg->events[ei];
			EventRec &er = events[ei];
			er.vo_id = pe->vo_id;
			er.carr_op_id = pe->carr_op_id;
			er.opdata_off = (uint32_t) opdata.size();
			er.opdata_count = pe->op_data_count;
			for (uint32_t k = 0; k < pe->op_data_count; ++k) {
				const sauabi_ProgramOpData *od = &pe->op_data[k];
				OpDataRec r;
				memset(&r, 0, sizeof r);
				r.id = od->id; r.params = od->params;
				r.time_samples = ms_in_samples(od->time.v_ms, srate, NULL);  /* generator.c:332 */
				r.time_flags = od->time.flags;
				r.type = od->type; r.use_type = od->use_type;
				r.mode_main = od->mode.main;
				flatten_line(&r.line[LINE_AMP], od->amp, srate);
				flatten_line(&r.line[LINE_AMP2], od->amp2, srate);
				flatten_line(&r.line[LINE_PAN], od->pan, srate);
				flatten_line(&r.line[LINE_FREQ], od->freq, srate);
				flatten_line(&r.line[LINE_FREQ2], od->freq2, srate);
				flatten_line(&r.line[LINE_PMA], od->pm_a, srate);
				r.phase = od->phase; r.seed = od->seed;
				if (od->type == SAUABI_POPT_raseg) {
					r.ras_flags = od->mode.ras.flags; r.ras_func = od->mode.ras.func;
					r.ras_level = od->mode.ras.level; r.ras_alpha = od->mode.ras.alpha;
				}
				if (od->type == SAUABI_POPT_wave) {
					if (od->params & SAUABI_POPP_MODE) o->wave_mask |= 1u << (od->mode.main % NUM_WAVES);
					o->wave_mask |= 1u << SAUABI_WAVE_sin;   /* sau_init_WOsc default */
				}
				opdata.push_back(r);
				if (od->id < hops.size()) {
					HostOp &h = hops[od->id];
					if (!h.inited) { h.inited = true; h.type = od->type; }
					if (od->amp) h.line_set[LINE_AMP] = true;
					if (od->amp2) h.line_set[LINE_AMP2] = true;
					if (od->pan) h.line_set[LINE_PAN] = true;
					if (od->freq) h.line_set[LINE_FREQ] = true;
					if (od->freq2) h.line_set[LINE_FREQ2] = true;
					if (od->pm_a) h.line_set[LINE_PMA] = true;
					if (od->type >= SAUABI_POPT_wave) {            /* generator.c:316-320 */
						if (od->fmods) h.mods[SAUABI_POP_fmod] = od->fmods;
						if (od->rfmods) h.mods[SAUABI_POP_rfmod] = od->rfmods;
						if (od->pmods) h.mods[SAUABI_POP_pmod] = od->pmods;
						if (od->apmods) h.mods[SAUABI_POP_apmod] = od->apmods;
						if (od->fpmods) h.mods[SAUABI_POP_fpmod] = od->fpmods;
					}
					if (od->camods) h.mods[SAUABI_POP_camod] = od->camods;  /* :337-339 */
					if (od->amods) h.mods[SAUABI_POP_amod] = od->amods;
					if (od->ramods) h.mods[SAUABI_POP_ramod] = od->ramods;
				}
			}
			er.code_off = 0; er.code_len = 0; er.ops_off = 0; er.ops_cnt = 0; er.carr_slot = 0;
			if (pe->vo_id != SAUABI_PVO_NO_ID && pe->vo_id < prg->vo_count) {
				vcarr[pe->vo_id] = pe->carr_op_id;
				comp.voice(pe->carr_op_id);
				/* reuse the voice's previous program when the walk is unchanged */
				auto &pv = vprog[pe->vo_id];
				auto &po = vops[pe->vo_id];
				bool same = pv.second == comp.out.size() && pv.second > 0 &&
					memcmp(&code[pv.first], comp.out.data(), pv.second * sizeof(Instr)) == 0 &&
					po.second == comp.prog_ops.size() && (po.second == 0 ||
					memcmp(&prog_ops[po.first], comp.prog_ops.data(), po.second * sizeof(uint32_t)) == 0);
				if (!same) {
					pv.first = (uint32_t) code.size();
					pv.second = (uint32_t) comp.out.size();
					code.insert(code.end(), comp.out.begin(), comp.out.end());
					po.first = (uint32_t) prog_ops.size();
					po.second = (uint32_t) comp.prog_ops.size();
					if (po.second > o->max_ops) o->max_ops = po.second;
					{
						/* records of the fast path's block plan (render_plan.cuh:steady_plan): the
						 * instructions that do something per chunk; a program with any other
						 * kind of instruction never takes that path */
						uint32_t np = 0;
						bool fast = true;
						for (const Instr &in : comp.out) {
							switch (in.opcode) {
							case I_WLEAF: case I_WTAIL: np += 2; break;      /* + the amplitude trajectory's slot */
							case I_WHEAD: case I_RANGE: case I_VOUT:
							case I_PHASOR: case I_WOSC: case I_NOISE: case I_CYCLOR: case I_RASG: case I_MIX:
							case I_PMA:
								++np; break;
							case I_LINE: if (in.d) ++np; break;
							case I_ENTER: case I_VPAN: case I_END: case I_LEAVE: break;
							default: fast = false; break;
							}
						}
						np += 3;                           /* the plan's two header slots and its end mark */
						if (fast && np <= 64 && np > o->nplan) o->nplan = np;
					}
					prog_ops.insert(prog_ops.end(), comp.prog_ops.begin(), comp.prog_ops.end());
				}
				er.code_off = pv.first; er.code_len = pv.second;
				er.ops_off = po.first; er.ops_cnt = po.second; er.carr_slot = 0;
				vev[pe->vo_id].push_back((uint32_t) ei);
				/* operators this event updates or this voice now renders, met under another voice before */
				auto touch = [&](uint32_t id) {
					if (id >= op_voice.size()) return;
					if (op_voice[id] != 0xffffffffu && op_voice[id] != pe->vo_id && !f.ev_handover[ei])
						f.ev_handover[ei] = op_voice[id] + 1u;
					op_voice[id] = pe->vo_id;
				};
				for (uint32_t k = 0; k < pe->op_data_count; ++k) touch(pe->op_data[k].id);
				for (uint32_t id : comp.prog_ops) touch(id);
			}
		}
		if (comp.too_deep) {
			g_err = "saugen_create: operator nesting too deep for the device interpreter";
			fprintf(stderr, "saugen_b200: error: %s\n", g_err.c_str());
			return false;
		}
		o->nbufs = comp.max_buf ? comp.max_buf : 1;
	}
	vev_off.resize(prg->vo_count + 1);
	for (uint32_t v = 0; v < prg->vo_count; ++v) {
		vev_off[v] = (uint32_t) vev_idx.size();
		vev_idx.insert(vev_idx.end(), vev[v].begin(), vev[v].end());
	}
	vev_off[prg->vo_count] = (uint32_t) vev_idx.size();

	return true;
}

static saugen_Generator *create_from_flat(const Flat &f, const saugen_WaveTables *tables,
		const saugen_Options *opt) {
	saugen_Options defopt;
	memset(&defopt, 0, sizeof defopt);
	if (!opt) opt = &defopt;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || opt->device >= ndev) {
		set_err("saugen_create: no usable CUDA device (this back end has no CPU path)", cudaGetLastError());
		return nullptr;
	}
	if (!tables) {
		/* the tables are input data built by the front-end library on the host (sau/wave.c:105-221);
		 * this back end does not regenerate them */
		g_err = "saugen_create: wave tables required (libsau's sauWave_piluts, or saugen_wave_tables_load)";
		fprintf(stderr, "saugen_b200: error: %s\n", g_err.c_str());
		return nullptr;
	}
	saugen_Generator *o = new saugen_Generator();
	long long tp = now_ns();
	auto lap = [&tp](int i) { if (g_cprof.on) { const long long t = now_ns(); g_cprof.ns[i] += t - tp; tp = t; } };
	const std::vector<EventRec> &events = f.events;
	const std::vector<OpDataRec> &opdata = f.opdata;
	const std::vector<Instr> &code = f.code;
	const std::vector<uint32_t> &prog_ops = f.prog_ops, &vev_off = f.vev_off, &vev_idx = f.vev_idx;
	const uint32_t srate = f.srate;
	o->srate = srate; o->device = opt->device; o->sched = opt->sched;
	o->big_endian = opt->pcm_big_endian != 0;
	o->vo_count = f.vo_count; o->op_count = f.op_count;
	o->voice_begin = 0; o->voice_end = f.vo_count;
	if (opt->voice_end > opt->voice_begin) {
		o->voice_begin = opt->voice_begin < f.vo_count ? opt->voice_begin : f.vo_count;
		o->voice_end = opt->voice_end < f.vo_count ? opt->voice_end : f.vo_count;
	}
	o->nlv = o->voice_end - o->voice_begin;
	o->row_len = opt->max_call_len ? opt->max_call_len : ms_in_samples(256, srate, NULL);  /* saugns.c:471 */
	o->row_len = (o->row_len + 3u) & ~3u;
	if (o->row_len < 4) o->row_len = 4;
	/* carrier rows are frame-tile major (device_types.h:ROW_TILE): the stride between
	 * tiles is one 512-byte piece per local voice */
	o->row_stride = (o->nlv ? o->nlv : 1u) * (uint32_t) ROW_TILE;
	o->amp_scale = f.amp_scale;
	o->ev_time = f.ev_time;
	o->ev_handover = f.ev_handover;
	if (o->voice_end - o->voice_begin < f.vo_count) {
		/* a voice shard keeps operator state for its own voices only: a hand-over between a
		 * voice inside and one outside cannot be rendered (multigpu.py shards around them) */
		for (size_t e = 0; e < f.ev_handover.size(); ++e) {
			if (!f.ev_handover[e]) continue;
			const uint32_t a = f.ev_handover[e] - 1u, b = f.events[e].vo_id;
			const bool ia = a >= o->voice_begin && a < o->voice_end, ib = b >= o->voice_begin && b < o->voice_end;
			if (ia != ib) {
				g_err = "saugen_create: an operator moves between voices on different shards "
					"(use saugen_voice_groups to choose shard boundaries)";
				fprintf(stderr, "saugen_b200: error: %s\n", g_err.c_str());
				delete o;
				return nullptr;
			}
		}
	}
	o->nbufs = f.nbufs; o->max_ops = f.max_ops; o->nplan = f.nplan; o->wave_mask = f.wave_mask;
	lap(0);
	/* ---- device allocation: one state block, one row block, one pinned block ---- */
	CK(cudaSetDevice(o->device));
	if (opt->stream) o->stream = (cudaStream_t) opt->stream;
	else {
		o->stream = g_streams.get(o->device);
		if (!o->stream) { set_err("saugen_create: stream", cudaGetLastError()); goto fail; }
		o->own_stream = true;
	}
	o->d_tables = get_device_tables(o->device, tables, &o->d_coefs);
	if (!o->d_tables) { set_err("wave table upload", cudaGetLastError()); goto fail; }
	{
		lap(1);
		const size_t nops = f.op_count ? f.op_count : 1, nvo = f.vo_count ? f.vo_count : 1;
		const size_t nl = o->nlv ? o->nlv : 1;
		o->seg_cap = 64;
		o->unit_cap = 256;
		Carver cv;
		/* static part first: filled in a host staging image, uploaded with one copy */
		const size_t o_events = cv.take(events.size() * sizeof(EventRec));
		const size_t o_opdata = cv.take(opdata.size() * sizeof(OpDataRec));
		const size_t o_code = cv.take(code.size() * sizeof(Instr));
		const size_t o_prog_ops = cv.take(prog_ops.size() * sizeof(uint32_t));
		const size_t o_vev_off = cv.take(vev_off.size() * sizeof(uint32_t));
		const size_t o_vev_idx = cv.take(vev_idx.size() * sizeof(uint32_t));
		const size_t o_desc = cv.take(sizeof(GenDesc));
		const size_t static_bytes = cv.off;
		const size_t o_ops = cv.take(nops * sizeof(OpState));
		const size_t o_voices = cv.take(nvo * sizeof(VoiceState));
		const size_t zero_bytes = cv.off - o_ops;         /* operator + voice state start zeroed */
		/* zeroed before every call with ONE memset: [vlen][progress + ticket][status];
		 * read back after every call with ONE copy: [status][pcm] */
		const size_t o_vlen = cv.take((size_t) o->seg_cap * nl * sizeof(VoiceSeg));
		const size_t o_progress = cv.take((nl + 1) * sizeof(uint32_t));   /* [nl] = ticket counter */
		const size_t o_status = cv.take((1 + o->seg_cap) * sizeof(uint32_t));
		const size_t o_pcm = cv.take(2 * (size_t) o->row_len * sizeof(int16_t));
		o->zero_bytes = o_status - o_vlen;             /* [vlen][progress + ticket]; the slot's status apart */
		o->zero_bytes0 = o_pcm - o_vlen;               /* ... with slot 0's status (the batched calls) */
		o->back_bytes_fixed = o_pcm - o_status;        /* status part of the read-back */
		/* two sets of float planes (one per call slot), each followed by MIX_TAIL floats the
		 * caller may use (multigpu.py folds its control words into the ONE reduced buffer) */
		const size_t mix_floats = 2 * (size_t) o->row_len + SAUGEN_MIX_TAIL;
		const size_t o_mix = cv.take(mix_floats * sizeof(float));
		const size_t o_mix1 = cv.take(mix_floats * sizeof(float));
		const size_t o_status1 = cv.take((1 + o->seg_cap) * sizeof(uint32_t));         /* the alternate call slot's */
		const size_t o_pcm1 = cv.take(2 * (size_t) o->row_len * sizeof(int16_t));
		const size_t o_snap = cv.take(zero_bytes);                                      /* run-ahead: state before it */
		/* written before every call with ONE copy: [call][segs][units] (same layout in
		 * the pinned block) */
		const size_t o_call = cv.take(sizeof(CallDesc));
		const size_t o_segs = cv.take(o->seg_cap * sizeof(SegDesc));
		const size_t o_units = cv.take(o->unit_cap * sizeof(UnitDesc));
		o->units_off_in_call = o_units - o_call;
		unsigned char *base = (unsigned char*) o->take(false, cv.off);
		if (!base) { set_err("saugen_create: device memory", cudaGetLastError()); goto fail; }
		o->d_events = base + o_events; o->d_opdata = base + o_opdata; o->d_code = base + o_code;
		o->d_prog_ops = base + o_prog_ops; o->d_vev_off = base + o_vev_off; o->d_vev_idx = base + o_vev_idx;
		o->d_desc = (GenDesc*) (base + o_desc);
		o->d_ops = base + o_ops; o->d_voices = base + o_voices;
		o->d_vlen = (VoiceSeg*) (base + o_vlen); o->d_status = (uint32_t*) (base + o_status);
		o->d_progress = (uint32_t*) (base + o_progress); o->d_units = (UnitDesc*) (base + o_units);
		o->d_mix = (float*) (base + o_mix); o->d_pcm = (int16_t*) (base + o_pcm);
		o->d_call = (CallDesc*) (base + o_call); o->d_segs = (SegDesc*) (base + o_segs);
		o->d_snap = base + o_snap; o->state_bytes = zero_bytes;
		o->slot[0].d_pcm = o->d_pcm; o->slot[1].d_pcm = (int16_t*) (base + o_pcm1);
		o->slot[0].d_status = o->d_status; o->slot[1].d_status = (uint32_t*) (base + o_status1);
		o->slot[0].d_mix = o->d_mix; o->slot[1].d_mix = (float*) (base + o_mix1);
		static const char *pcenv = getenv("SAUGEN_PLANCACHE");    /* developer knob: 0 = stable plans are not kept */
		if (o->nplan && o->d_coefs && !(pcenv && pcenv[0] == '0')) {
			const size_t pcb = nl * (18 + 2 * (size_t) o->nplan) * 16;     /* header, records, the teams' analysis */
			o->d_plan_cache = o->take(false, pcb);
			if (o->d_plan_cache) CK(cudaMemsetAsync(o->d_plan_cache, 0, pcb, o->stream));
		}
		const size_t ntile = ((size_t) o->row_len + ROW_TILE - 1) / ROW_TILE;
		float *rows = (float*) o->take(false, 2 * ntile * (size_t) o->row_stride * sizeof(float));
		if (!rows) { set_err("saugen_create: device memory (carrier rows)", cudaGetLastError()); goto fail; }
		o->d_rows_s = rows; o->d_rows_r = rows + ntile * (size_t) o->row_stride;
		lap(2);
		Carver hv;
		const size_t h_status = hv.take((1 + o->seg_cap) * sizeof(uint32_t));
		const size_t h_pcm = hv.take(2 * (size_t) o->row_len * sizeof(int16_t));
		const size_t h_call = hv.take(sizeof(CallDesc));
		const size_t h_segs = hv.take(o->seg_cap * sizeof(SegDesc));
		const size_t h_units = hv.take(o->unit_cap * sizeof(UnitDesc));
		const size_t h_status1 = hv.take((1 + o->seg_cap) * sizeof(uint32_t));
		const size_t h_pcm1 = hv.take(2 * (size_t) o->row_len * sizeof(int16_t));
		const size_t h_call1 = hv.take(sizeof(CallDesc));
		const size_t h_segs1 = hv.take(o->seg_cap * sizeof(SegDesc));
		const size_t h_units1 = hv.take(o->unit_cap * sizeof(UnitDesc));
		if (h_units1 - h_call1 != o_units - o_call || h_segs1 - h_call1 != o_segs - o_call) o->compact = false;
		if (h_pcm1 - h_status1 != o_pcm1 - o_status1) o->compact = false;
		if (h_pcm - h_status != o_pcm - o_status || h_units - h_call != o_units - o_call ||
				h_segs - h_call != o_segs - o_call) o->compact = false;
		unsigned char *hb = (unsigned char*) o->take(true, hv.off);
		if (!hb) { set_err("saugen_create: pinned host memory", cudaGetLastError()); goto fail; }
		o->h_status = (uint32_t*) (hb + h_status); o->h_pcm = (int16_t*) (hb + h_pcm);
		o->h_call = (CallDesc*) (hb + h_call); o->h_segs = (SegDesc*) (hb + h_segs);
		o->h_units = (UnitDesc*) (hb + h_units);
		o->slot[0].h_status = o->h_status; o->slot[0].h_pcm = o->h_pcm;
		o->slot[1].h_status = (uint32_t*) (hb + h_status1); o->slot[1].h_pcm = (int16_t*) (hb + h_pcm1);
		o->slot[0].h_call = o->h_call; o->slot[0].h_segs = o->h_segs; o->slot[0].h_units = o->h_units;
		o->slot[1].h_call = (CallDesc*) (hb + h_call1); o->slot[1].h_segs = (SegDesc*) (hb + h_segs1);
		o->slot[1].h_units = (UnitDesc*) (hb + h_units1);
		lap(3);
		lap(4);

		GenDesc &d = o->h_desc;
		memset(&d, 0, sizeof d);
		d.ops = (OpState*) o->d_ops; d.voices = (VoiceState*) o->d_voices;
		d.events = (const EventRec*) o->d_events; d.opdata = (const OpDataRec*) o->d_opdata;
		d.code = (const Instr*) o->d_code;
		d.prog_ops = (const uint32_t*) o->d_prog_ops;
		d.vev_off = (const uint32_t*) o->d_vev_off; d.vev_idx = (const uint32_t*) o->d_vev_idx;
		d.rows_s = o->d_rows_s; d.rows_r = o->d_rows_r;
		d.vlen = o->d_vlen; d.status = o->d_status; d.vlen_cap = o->seg_cap;
		d.progress = o->d_progress; d.ticket = o->d_progress + nl;
		d.mix = o->d_mix; d.pcm = o->d_pcm;
		d.vo_count = o->vo_count; d.op_count = o->op_count;
		d.voice_begin = o->voice_begin; d.voice_end = o->voice_end;
		d.row_len = o->row_len; d.row_stride = o->row_stride; d.nbufs = o->nbufs; d.srate = srate;
		d.coeff = (float) (4294967296.0 / srate);                 /* wosc.h:30, math.h:386 */
		d.amp_scale = o->amp_scale;
		d.wave_mask = o->wave_mask; d.tables = o->d_tables;
		d.plan_cache = (uint4*) o->d_plan_cache; d.plan_cache_recs = o->nplan;

		/* one H2D copy of the static image, one memset of the run-time state; the
		 * stream is synchronised before the staging image goes away */
		std::vector<unsigned char> img(static_bytes);
		auto put = [&](size_t off, const void *src, size_t n) { if (n) memcpy(&img[off], src, n); };
		put(o_events, events.data(), events.size() * sizeof(EventRec));
		put(o_opdata, opdata.data(), opdata.size() * sizeof(OpDataRec));
		put(o_code, code.data(), code.size() * sizeof(Instr));
		put(o_prog_ops, prog_ops.data(), prog_ops.size() * sizeof(uint32_t));
		put(o_vev_off, vev_off.data(), vev_off.size() * sizeof(uint32_t));
		put(o_vev_idx, vev_idx.data(), vev_idx.size() * sizeof(uint32_t));
		put(o_desc, &d, sizeof d);
		CK(cudaMemcpyAsync(base, img.data(), static_bytes, cudaMemcpyHostToDevice, o->stream));
		CK(cudaMemsetAsync(base + o_ops, 0, zero_bytes, o->stream));
		/* (the status blocks are read back whole with the PCM: no uninitialised bytes in that copy) */
		CK(cudaMemsetAsync(base + o_status, 0, o_pcm - o_status, o->stream));
		CK(cudaMemsetAsync(base + o_status1, 0, o_pcm1 - o_status1, o->stream));
		CK(cudaStreamSynchronize(o->stream));
		lap(5);
		if (g_cprof.on) g_cprof.n++;
	}
	return o;
fail:
	saugen_destroy(o);
	return nullptr;
}

extern "C" saugen_Generator *saugen_create(const sauabi_Program *prg, uint32_t srate,
		const saugen_WaveTables *tables, const saugen_Options *opt) {
	if (!prg || !srate) { g_err = "saugen_create: NULL program or zero sample rate"; return nullptr; }
	int ndev = 0;              /* no device: fail before anything else (there is no CPU path) */
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || (opt && opt->device >= ndev)) {
		set_err("saugen_create: no usable CUDA device (this back end has no CPU path)", cudaGetLastError());
		return nullptr;
	}
	Flat f;
	if (!flatten_program(prg, srate, f)) return nullptr;
	return create_from_flat(f, tables, opt);
}

/* Voices linked by operator hand-overs (Flat::ev_handover) must be rendered by one generator:
 * group_of_voice[v] = the smallest voice index of v's group.  Returns the number of hand-over
 * events (0 = every voice is independent), <0 on error.  Needs no GPU. */
extern "C" int saugen_voice_groups(const sauabi_Program *prg, uint32_t srate, uint32_t *group_of_voice) {
	if (!prg || !srate) { g_err = "saugen_voice_groups: NULL program or zero sample rate"; return -1; }
	Flat f;
	if (!flatten_program(prg, srate, f)) return -1;
	std::vector<uint32_t> parent(f.vo_count);
	for (uint32_t v = 0; v < f.vo_count; ++v) parent[v] = v;
	auto find = [&](uint32_t v) { while (parent[v] != v) { parent[v] = parent[parent[v]]; v = parent[v]; } return v; };
	int n = 0;
	for (size_t e = 0; e < f.ev_handover.size(); ++e) {
		if (!f.ev_handover[e]) continue;
		++n;
		const uint32_t a = find(f.ev_handover[e] - 1u), b = find(f.events[e].vo_id);
		if (a != b) { if (a < b) parent[b] = a; else parent[a] = b; }
	}
	if (group_of_voice) for (uint32_t v = 0; v < f.vo_count; ++v) group_of_voice[v] = find(v);
	return n;
}

/* The flat program as one relocatable blob: a header of counts, then the arrays. */
namespace {
struct FlatHeader {
	uint32_t magic, version, srate, vo_count, op_count, nbufs, max_ops, nplan, wave_mask;
	float amp_scale;
	uint64_t n_ev_time, n_events, n_opdata, n_code, n_prog_ops, n_vev_off, n_vev_idx, n_ev_handover;
};
const uint32_t FLAT_MAGIC = 0x46554153u /* "SAUF" */, FLAT_VERSION = 2;
template <typename T> size_t blob_bytes(const std::vector<T> &v) { return (v.size() * sizeof(T) + 7) & ~(size_t) 7; }
}

extern "C" size_t saugen_flatten(const sauabi_Program *prg, uint32_t srate, void *blob, size_t cap) {
	if (!prg || !srate) { g_err = "saugen_flatten: NULL program or zero sample rate"; return 0; }
	Flat f;
	if (!flatten_program(prg, srate, f)) return 0;
	const size_t need = sizeof(FlatHeader) + blob_bytes(f.ev_time) + blob_bytes(f.events) + blob_bytes(f.opdata) +
		blob_bytes(f.code) + blob_bytes(f.prog_ops) + blob_bytes(f.vev_off) + blob_bytes(f.vev_idx) +
		blob_bytes(f.ev_handover);
	if (!blob || cap < need) return need;
	memset(blob, 0, need);
	FlatHeader h;
	memset(&h, 0, sizeof h);
	h.magic = FLAT_MAGIC; h.version = FLAT_VERSION; h.srate = f.srate; h.vo_count = f.vo_count;
	h.op_count = f.op_count; h.nbufs = f.nbufs; h.max_ops = f.max_ops; h.nplan = f.nplan;
	h.wave_mask = f.wave_mask; h.amp_scale = f.amp_scale;
	h.n_ev_time = f.ev_time.size(); h.n_events = f.events.size(); h.n_opdata = f.opdata.size();
	h.n_code = f.code.size(); h.n_prog_ops = f.prog_ops.size(); h.n_vev_off = f.vev_off.size();
	h.n_vev_idx = f.vev_idx.size(); h.n_ev_handover = f.ev_handover.size();
	unsigned char *p = (unsigned char*) blob;
	memcpy(p, &h, sizeof h); p += sizeof h;
	auto put = [&p](const void *src, size_t n, size_t padded) { if (n) memcpy(p, src, n); p += padded; };
	put(f.ev_time.data(), f.ev_time.size() * sizeof(uint64_t), blob_bytes(f.ev_time));
	put(f.events.data(), f.events.size() * sizeof(EventRec), blob_bytes(f.events));
	put(f.opdata.data(), f.opdata.size() * sizeof(OpDataRec), blob_bytes(f.opdata));
	put(f.code.data(), f.code.size() * sizeof(Instr), blob_bytes(f.code));
	put(f.prog_ops.data(), f.prog_ops.size() * sizeof(uint32_t), blob_bytes(f.prog_ops));
	put(f.vev_off.data(), f.vev_off.size() * sizeof(uint32_t), blob_bytes(f.vev_off));
	put(f.vev_idx.data(), f.vev_idx.size() * sizeof(uint32_t), blob_bytes(f.vev_idx));
	put(f.ev_handover.data(), f.ev_handover.size() * sizeof(uint32_t), blob_bytes(f.ev_handover));
	return need;
}

/* A blob comes from outside the process: every index and count in it is checked against
 * the arrays it refers to before anything goes to the device (a stale or corrupt blob
 * must not become out-of-bounds accesses in render_kernel). */
static const char *validate_flat(const Flat &f) {
	const size_t nev = f.events.size();
	if (f.vo_count > 65535u || !f.nbufs || f.nbufs > 250u || !f.max_ops || f.nplan > 64u) return "header counts";
	if (f.vev_off.size() != (size_t) f.vo_count + 1 || f.vev_off[0] != 0) return "voice event index";
	for (uint32_t v = 0; v < f.vo_count; ++v)
		if (f.vev_off[v] > f.vev_off[v + 1]) return "voice event index";
	if (f.vev_off[f.vo_count] != f.vev_idx.size()) return "voice event index";
	for (uint32_t i : f.vev_idx) if (i >= nev) return "voice event list";
	for (size_t i = 1; i < f.ev_time.size(); ++i) if (f.ev_time[i] < f.ev_time[i - 1]) return "event times";
	for (uint32_t h : f.ev_handover) if (h > f.vo_count) return "hand-over voice";
	for (const OpDataRec &r : f.opdata) {
		if (r.id >= f.op_count || r.type > SAUABI_POPT_raseg) return "op-data record";
		if (r.type == SAUABI_POPT_wave && (r.params & SAUABI_POPP_MODE) && r.mode_main >= NUM_WAVES) return "wave id";
		for (int l = 0; l < LINE_COUNT; ++l) if (r.line[l].present && r.line[l].type >= SAUABI_LINE_NAMED) return "line type";
	}
	for (uint32_t id : f.prog_ops) if (id >= f.op_count) return "program operator list";
	for (const EventRec &e : f.events) {
		if ((size_t) e.opdata_off + e.opdata_count > f.opdata.size()) return "event op-data range";
		if ((size_t) e.code_off + e.code_len > f.code.size()) return "event code range";
		if ((size_t) e.ops_off + e.ops_cnt > f.prog_ops.size() || e.ops_cnt > f.max_ops) return "event operator range";
		if (e.vo_id != SAUABI_PVO_NO_ID && e.vo_id >= f.vo_count) return "event voice";
		if (e.code_len && (e.carr_slot >= e.ops_cnt && e.ops_cnt)) return "carrier slot";
		uint32_t depth = 0;
		for (uint32_t k = 0; k < e.code_len; ++k) {
			const Instr &in = f.code[e.code_off + k];
			if (in.opcode < I_ENTER || in.opcode > I_WLEAF) return "opcode";
			const uint8_t m = Compiler::buf_fields(in.opcode, in.d, in.flags);
			const uint8_t fld[5] = {in.a, in.b, in.c, in.d, in.e};
			for (int q = 0; q < 5; ++q)
				if ((m >> q & 1) && fld[q] != NO_BUF && fld[q] >= f.nbufs) return "work buffer id";
			if ((m & 32) && in.b != NO_BUF && (uint32_t) in.b + 1 >= f.nbufs) return "work buffer id";
			if (in.opcode == I_LINE && in.c >= LINE_COUNT) return "line index";
			const bool uses_op = in.opcode != I_ZERO && in.opcode != I_RANGE && in.opcode != I_MIX && in.opcode != I_END;
			if (uses_op && in.op >= (e.ops_cnt ? e.ops_cnt : 1u)) return "operator slot";
			if (in.opcode == I_ENTER || in.opcode == I_WHEAD) {
				if (++depth >= (uint32_t) MAX_NEST - 1) return "nesting depth";
				if (in.aux > e.code_len) return "enter target";
			}
			if ((in.opcode == I_LEAVE || in.opcode == I_WTAIL) && depth) --depth;
		}
	}
	return nullptr;
}

extern "C" saugen_Generator *saugen_create_flat(const void *blob, size_t size,
		const saugen_WaveTables *tables, const saugen_Options *opt) {
	FlatHeader h;
	if (!blob || size < sizeof h) { g_err = "saugen_create_flat: no blob"; return nullptr; }
	memcpy(&h, blob, sizeof h);
	if (h.magic != FLAT_MAGIC || h.version != FLAT_VERSION || !h.srate) {
		g_err = "saugen_create_flat: not a flat program of this library version";
		return nullptr;
	}
	Flat f;
	f.srate = h.srate; f.vo_count = h.vo_count; f.op_count = h.op_count; f.nbufs = h.nbufs;
	f.max_ops = h.max_ops; f.nplan = h.nplan; f.wave_mask = h.wave_mask; f.amp_scale = h.amp_scale;
	const unsigned char *p = (const unsigned char*) blob + sizeof h, *end = (const unsigned char*) blob + size;
	bool ok = true;
	auto get = [&](auto &vec, uint64_t n) {
		typedef typename std::remove_reference<decltype(vec)>::type::value_type T;
		const size_t bytes = (size_t) n * sizeof(T), padded = (bytes + 7) & ~(size_t) 7;
		if (!ok || n > ((uint64_t) 1 << 32) || (size_t) (end - p) < padded) { ok = false; return; }
		vec.resize((size_t) n);
		if (bytes) memcpy(vec.data(), p, bytes);
		p += padded;
	};
	get(f.ev_time, h.n_ev_time); get(f.events, h.n_events); get(f.opdata, h.n_opdata); get(f.code, h.n_code);
	get(f.prog_ops, h.n_prog_ops); get(f.vev_off, h.n_vev_off); get(f.vev_idx, h.n_vev_idx);
	get(f.ev_handover, h.n_ev_handover);
	if (!ok || f.vev_off.size() != (size_t) f.vo_count + 1 || f.ev_time.size() != f.events.size() ||
			f.ev_handover.size() != f.events.size()) {
		g_err = "saugen_create_flat: truncated or inconsistent blob";
		return nullptr;
	}
	if (const char *bad = validate_flat(f)) {
		g_err = std::string("saugen_create_flat: blob fails validation (") + bad + ")";
		return nullptr;
	}
	return create_from_flat(f, tables, opt);
}

extern "C" void saugen_destroy(saugen_Generator *o) {
	if (!o) return;
	cudaSetDevice(o->device);
	if (o->stream) cudaStreamSynchronize(o->stream);
	if (o->copy_stream) { cudaStreamSynchronize(o->copy_stream); g_streams.put(o->device, o->copy_stream); }
	if (o->aux_stream) { cudaStreamSynchronize(o->aux_stream); g_streams.put(o->device, o->aux_stream); }
	if (o->d_tap) cudaFree(o->d_tap);
	for (auto &b : o->blocks) g_pool.release(b.second, o->device, b.first);
	for (auto &sl : o->slot) {
		if (sl.done) cudaEventDestroy(sl.done);
		if (sl.mixed) cudaEventDestroy(sl.mixed);
		for (int i = 0; i < 3; ++i) if (sl.ev_t[i]) cudaEventDestroy(sl.ev_t[i]);
	}
	if (o->own_stream && o->stream) g_streams.put(o->device, o->stream);
	delete o;
}

/* Replay of the PROCESS loop of sauGenerator_run (generator.c:915-949): cut the
 * call into inter-event segments; ev_end says which events are due by then. */
static void plan_call(saugen_Generator *o, uint32_t buf_len, std::vector<SegDesc> &segs) {
	segs.clear();
	o->group_first.clear();
	uint64_t t = o->cur_time;
	const uint64_t t_end = t + buf_len;
	size_t ev = o->next_event;
	const size_t nev = o->ev_time.size();
	/* events due at time t, in order.  A hand-over event (Flat::ev_handover) starts a new
	 * render launch: the events due before it at the same instant get a zero-length
	 * segment of their own in the launch before, so that every earlier touch of the
	 * operator has gone through a kernel boundary when its new voice picks it up. */
	auto take_due = [&](uint64_t now) {
		bool cut = false;
		size_t from = ev;
		while (ev < nev && o->ev_time[ev] <= now) {
			if (o->ev_handover[ev]) {
				if (ev > from) {
					SegDesc z; z.start = (uint32_t) (now - o->cur_time); z.len = 0; z.ev_end = (uint32_t) ev;
					if (cut && !segs.empty()) o->group_first.push_back((uint32_t) segs.size());
					segs.push_back(z);
					from = ev;
				}
				cut = true;
			}
			++ev;
		}
		if (cut && !segs.empty()) o->group_first.push_back((uint32_t) segs.size());
	};
	while (t < t_end) {
		take_due(t);
		uint64_t stop = t_end;
		if (ev < nev && o->ev_time[ev] < stop) stop = o->ev_time[ev];
		SegDesc s;
		s.start = (uint32_t) (t - o->cur_time);
		s.len = (uint32_t) (stop - t);
		s.ev_end = (uint32_t) ev;
		segs.push_back(s);
		t = stop;
	}
	if (buf_len == 0) {                  /* events due now are still handled */
		take_due(t);
		SegDesc s; s.start = 0; s.len = 0; s.ev_end = (uint32_t) ev;
		segs.push_back(s);
	}
	o->next_event = ev;
	o->cur_time = t_end;
}

/* Cut the segments of a call into schedulable units of at most UNIT_BLOCKS
 * reference blocks, on the segment's own block grid. */
static void plan_units(const std::vector<SegDesc> &segs, std::vector<UnitDesc> &units,
		uint32_t unit_blocks = 4) {
	units.clear();
	const uint32_t ul = unit_blocks * REF_BLOCK;
	for (uint32_t s = 0; s < segs.size(); ++s) {
		uint32_t off = 0;
		do {
			UnitDesc u;
			u.seg = s; u.off = off;
			u.len = segs[s].len - off < ul ? segs[s].len - off : ul;
			units.push_back(u);
			off += u.len;
		} while (off < segs[s].len);
	}
}

/* Launch shape of render_kernel for `ntasks` voice tasks.
 * The path is latency-bound: what counts is how many
 * voices are resident per SM at once.  One CTA per SM, with as many warps as it
 * takes to hold every task in ONE resident wave (up to what shared memory
 * allows, at most 32); up to 8 warps run the 128-register kernel, more the
 * 64-register one.  Tables: coefficient planes (render_ops.cuh:CTAB_FLAG, 48 KiB per
 * wave) when the launch uses at most two waves, else the float tables (8 KiB
 * per wave in use). */
static const uint32_t CTAB_FLAG = 0x80000000u;
struct Shape { uint32_t warps; uint32_t mask; uint32_t team; };
/* team: with fewer voices than an SM has room for warps, every voice gets a TEAM of warps of
 * its CTA that split steady stretches along time (render_team.cuh): as many members as fit,
 * one named barrier per team (ids 1..15).  allow_team: not under the ticketed schedulers. */
static Shape pick_shape(uint32_t ntasks, uint32_t wave_mask, uint32_t nbufs, uint32_t max_ops,
		uint32_t nplan, bool have_coefs, bool allow_team = true) {
	const uint32_t sms = (uint32_t) device_sm_count();
	const size_t SMEM_CAP = device_smem_optin();
	int nw = 0;
	for (uint32_t w = 0; w < NUM_WAVES; ++w) if (wave_mask & (1u << w)) ++nw;
	static const char *env = getenv("SAUGEN_CTAB");       /* developer knob: 0 = off */
	static const char *tenv = getenv("SAUGEN_TEAM");      /* developer knob: 0 = off, n = at most n members */
	/* coefficient planes (48 KiB per wave): two waves leave room for a full CTA of voices; with few
	 * voices (fewer warps) up to four waves' planes fit beside them */
	const bool want_ctab = have_coefs && nw >= 1 && nw <= 4 && !(env && env[0] == '0');
	static const char *fenv = getenv("SAUGEN_FUSED");     /* developer knob: 0 = no fused shapes */
	const uint32_t nofuse = (fenv && fenv[0] == '0') ? 0x40000000u : 0u;
	Shape sh;
	sh.team = 1;
	for (int pass = want_ctab ? 0 : 1; pass < 2; ++pass) {
		sh.mask = (pass == 0 ? (wave_mask | CTAB_FLAG) : wave_mask) | nofuse;
		uint32_t fit = 28;                 /* kernels.cu:WIDE_WARPS */
		while (fit > 1 && render_smem_bytes(sh.mask, nbufs, max_ops, nplan, fit, 1) > SMEM_CAP) --fit;
		sh.warps = (ntasks + sms - 1) / sms;
		if (sh.warps < 1) sh.warps = 1;
		if (sh.warps > fit) sh.warps = fit;
		if (render_smem_bytes(sh.mask, nbufs, max_ops, nplan, sh.warps, 1) <= SMEM_CAP &&
				(pass == 1 || fit >= 8) && (pass == 1 || nw <= 2 || fit >= (ntasks + sms - 1) / sms)) {
			/* (teams split lowered plans: coefficient-plane launches only) */
			const uint32_t per_cta = sh.warps;             /* voices per CTA */
			if (allow_team && pass == 0 && nplan && per_cta <= 14 && ntasks <= per_cta * sms &&
					!(tenv && tenv[0] == '0')) {
				uint32_t t = 28 / per_cta;
				if (tenv && atoi(tenv) > 0 && (uint32_t) atoi(tenv) < t) t = (uint32_t) atoi(tenv);
				while (t > 1 && render_smem_bytes(sh.mask, nbufs, max_ops, nplan, per_cta * t, t) > SMEM_CAP) --t;
				/* (two members per voice do not pay: with 14 voices an SM is close to its throughput
				 * bound already, and the lead-in and phase overheads come on top -- measured 0.63 vs 0.47 ms) */
				if (t > 2) { sh.team = t; sh.warps = per_cta * t; }
			}
			return sh;
		}
	}
	/* not even one warp of this voice program fits the SM's shared memory */
	if (render_smem_bytes(sh.mask, nbufs, max_ops, nplan, sh.warps, 1) > SMEM_CAP) sh.warps = 0;
	return sh;
}

/* mode: 0 = PCM in device memory, 1 = float planes */
static cudaError_t read_back(saugen_Generator *o, uint32_t nseg, size_t host_pcm_bytes, cudaStream_t st,
		int si = 0) {
	saugen_Generator::CallSlot &sl = o->slot[si];
	if (o->compact && host_pcm_bytes)                 /* [status][pcm] in one copy */
		return cudaMemcpyAsync(sl.h_status, sl.d_status, o->back_bytes_fixed + host_pcm_bytes,
				cudaMemcpyDeviceToHost, st);
	cudaError_t e = cudaMemcpyAsync(sl.h_status, sl.d_status, (1 + nseg) * sizeof(uint32_t),
			cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess && host_pcm_bytes)
		e = cudaMemcpyAsync(sl.h_pcm, sl.d_pcm, host_pcm_bytes, cudaMemcpyDeviceToHost, st);
	return e;
}

/* Room for n inter-event segments of one call (rare: > 64 events inside a call): new,
 * larger pieces; the old ones stay with the generator until destroy. */
static bool ensure_seg_cap(saugen_Generator *o, size_t n) {
	if (n <= o->seg_cap) return true;
	uint32_t cap = o->seg_cap;
	while (cap < n) cap *= 2;
	size_t nl = o->nlv ? o->nlv : 1;
	cudaStreamSynchronize(o->stream);
	o->d_vlen = (VoiceSeg*) o->take(false, (size_t) cap * nl * sizeof(VoiceSeg));
	o->d_status = (uint32_t*) o->take(false, (1 + cap) * sizeof(uint32_t));
	o->slot[0].d_status = o->d_status;
	o->slot[1].d_status = (uint32_t*) o->take(false, (1 + cap) * sizeof(uint32_t));
	o->d_segs = (SegDesc*) o->take(false, cap * sizeof(SegDesc));
	o->h_status = (uint32_t*) o->take(true, (1 + cap) * sizeof(uint32_t));
	o->h_segs = (SegDesc*) o->take(true, cap * sizeof(SegDesc));
	o->slot[0].h_status = o->h_status;
	o->slot[1].h_status = (uint32_t*) o->take(true, (1 + cap) * sizeof(uint32_t));
	o->slot[0].h_segs = o->h_segs;
	o->slot[1].h_segs = (SegDesc*) o->take(true, cap * sizeof(SegDesc));
	if (!o->d_vlen || !o->d_status || !o->d_segs || !o->h_status || !o->h_segs || !o->slot[1].h_status ||
			!o->slot[1].h_segs) {
		set_err("saugen_run: segment table growth", cudaGetLastError());
		return false;
	}
	o->seg_cap = cap;
	o->compact = false;
	o->h_desc.vlen = o->d_vlen; o->h_desc.status = o->d_status; o->h_desc.vlen_cap = cap;
	return cudaMemcpy(o->d_desc, &o->h_desc, sizeof(GenDesc), cudaMemcpyHostToDevice) == cudaSuccess;
}

/* The cache of a team launch (render_team.cuh): per voice TEAM_SLOTS slots of one call's frames + every
 * member's lead-in.  Made when the first team launch needs it; none (allocation refused) = only plans
 * without frequency modulation are split along time. */
static bool ensure_team_cache(saugen_Generator *o, uint32_t team, bool mail) {
	const uint32_t stride = (o->row_len + (team - 1u) * TEAM_LEAD_CHUNKS * 128u + 3u) & ~3u;
	const size_t nl = o->nlv ? o->nlv : 1;
	bool changed = false;
	if (!(o->h_desc.team_cache && o->h_desc.team_cache_stride >= stride) && !o->team_cache_failed) {
		float *p = (float*) o->take(false, nl * TEAM_SLOTS * stride * sizeof(float));
		if (!p) { o->team_cache_failed = true; cudaGetLastError(); }
		else { o->h_desc.team_cache = p; o->h_desc.team_cache_stride = stride; changed = true; }
	}
	if (mail && !o->h_desc.team_mail && !o->team_cache_failed) {
		/* a team over several CTAs: the voices' mailboxes and {arrivals, epoch} words */
		const uint32_t mb = team_mail_bytes(plan_area_bytes(o->nplan), o->max_ops ? o->max_ops : 1);
		unsigned char *p = (unsigned char*) o->take(false, nl * mb + nl * 64);
		if (!p) { o->team_cache_failed = true; cudaGetLastError(); }
		else {
			cudaMemsetAsync(p, 0, nl * mb + nl * 64, o->stream);
			o->h_desc.team_mail = p; o->h_desc.team_mail_stride = mb;
			o->h_desc.team_hdr = (uint32_t*) (p + nl * mb);
			changed = true;
		}
	}
	if (changed) {
		cudaStreamSynchronize(o->stream);
		if (cudaMemcpy(o->d_desc, &o->h_desc, sizeof(GenDesc), cudaMemcpyHostToDevice) != cudaSuccess) return false;
	}
	return o->h_desc.team_cache != nullptr && (!mail || o->h_desc.team_mail != nullptr);
}

/* Plans and launches one call into slot `si` (kernels + read-back queued on the generator's
 * stream, an event recorded after them); nothing waits.  host_pcm_bytes: PCM bytes to bring to
 * the slot's pinned staging buffer (0 = none).  <0 on error. */
static int launch_call(saugen_Generator *o, int si, size_t buf_len, int stereo, uint32_t mode,
		size_t host_pcm_bytes, bool ahead = false) {
	saugen_Generator::CallSlot &sl = o->slot[si];
	size_t *out_len = nullptr;
	std::vector<SegDesc> &segs = o->segs_tmp;
	plan_call(o, (uint32_t) buf_len, segs);
	/* (a run-ahead call never grows the tables: the call before it is still in flight on them) */
	if (ahead && segs.size() > o->seg_cap) return -2;
	if (!ensure_seg_cap(o, segs.size())) { if (out_len) *out_len = 0; return -1; }
	if (!o->copy_stream) {
		o->copy_stream = g_streams.get(o->device);
		if (!o->copy_stream) { set_err("saugen_run: stream", cudaGetLastError()); return -1; }
	}
	const uint32_t nseg = (uint32_t) segs.size();
	sl.ev_after = o->next_event;
	memcpy(sl.h_segs, segs.data(), nseg * sizeof(SegDesc));
	/* Launch shape and scheduling.  sched: 0 = auto, 1 = one warp per voice, 2 =
	 * persistent grid with (unit, voice) tickets, 3 = balanced contiguous ranges.
	 * Auto picks balanced when the voices would otherwise need a second, partly
	 * filled wave of warps (between 1 and 4 waves), else one warp per voice. */
	Shape shape = pick_shape(o->nlv, o->wave_mask, o->nbufs, o->max_ops, o->nplan, o->d_coefs != nullptr && !o->d_tap,
			(o->sched == 0 || o->sched == 1) && !o->d_tap);
	if (o->d_tap) shape.mask |= 0x20000000u;         /* render_ops.cuh:TAP_FLAG */
	{
		static const char *venv = getenv("SAUGEN_PLAN_VERIFY");
		if (venv && venv[0] == '1') shape.mask |= 0x10000000u;     /* render_ops.cuh:VERIFY_FLAG */
	}
	/* one voice per CTA and SMs to spare: the voice's team spans `multi` CTAs (render_team.cuh) -- as many as
	 * its members can use at this call length, at most what leaves every voice the same number */
	uint32_t multi = 1;
	if (shape.team > 1 && shape.warps == shape.team) {
		static const char *menv = getenv("SAUGEN_MULTI");      /* developer knob: 0 = off, n = at most n CTAs */
		const uint32_t sms = (uint32_t) device_sm_count(), nl = o->nlv ? o->nlv : 1;
		const uint32_t wanted = (uint32_t) (buf_len / 128) / 2u;       /* members at two chunks each */
		uint32_t k = (wanted + shape.team - 1) / shape.team;
		if (k > TEAM_MAX_CTAS) k = TEAM_MAX_CTAS;
		if (k > sms / nl) k = sms / nl;
		if (menv && (uint32_t) atoi(menv) < k) k = (uint32_t) atoi(menv);
		if (k >= 2 && o->max_ops <= TEAM_MAIL_OPS && ensure_team_cache(o, shape.team * k, true)) multi = k;
	}
	if (shape.team > 1 && multi == 1) ensure_team_cache(o, shape.team, false);      /* (none: PM-only plans still split) */
	const uint32_t warps = shape.warps;
	if (!warps) {
		g_err = "saugen_run: a voice program of this script needs more shared memory than one SM has";
		fprintf(stderr, "saugen_b200: error: %s\n", g_err.c_str());
		if (out_len) *out_len = 0;
		return -1;
	}
	uint32_t ticketed_ctas = 0, sched_mode = 0;
	{
		const size_t smem = render_smem_bytes(shape.mask, o->nbufs, o->max_ops, o->nplan, warps, shape.team);
		int per_sm = render_ctas_per_sm(smem, warps);
		if (per_sm < 1) per_sm = 1;
		const uint32_t resident_ctas = (uint32_t) device_sm_count() * (uint32_t) per_sm;
		const uint32_t need = (o->nlv + warps / shape.team - 1) / (warps / shape.team);
		uint32_t sched = o->sched;
		if (sched == 0)
			sched = need > resident_ctas ? 3 : 1;        /* more than one wave: balanced ranges */
		if (sched == 3 && need <= resident_ctas && !getenv("SAUGEN_FORCE_BALANCED")) sched = 1;   /* a single wave is balanced already */
		if (sched == 2 || sched == 3) {
			ticketed_ctas = resident_ctas < need ? resident_ctas : (need ? need : 1);
			sched_mode = sched == 2 ? 1 : 2;
			if (getenv("SAUGEN_ONE_CTA")) ticketed_ctas = 1;
		}
	}
	/* Unit size.  One warp per voice: a unit is a whole segment (the steady-stretch plan
	 * then covers it in one go).  Ticketed: 4 blocks.  Balanced: the warps of the resident
	 * grid each take items / warps (+-1) units, so large units leave a coarse split
	 * (the busiest warp sets the time: ceil(items / warps) units) while small units pay
	 * the per-stretch plan more often (measured: time ~ 1 + 0.57 / blocks per unit);
	 * pick the size with the best product of the two. */
	uint32_t unit_blocks = sched_mode == 1 ? 4u : (1u << 20);
	if (sched_mode == 2) {
		static const char *ub = getenv("SAUGEN_UNIT_BLOCKS");     /* developer knob */
		const double S = (double) ticketed_ctas * warps;
		double best = -1.0;
		unit_blocks = 1;
		for (uint32_t cand : {1u, 2u, 3u, 4u, 6u, 8u, 12u, 16u, 24u, 32u, 48u, 96u}) {
			double units = 0;
			for (const SegDesc &sd : segs) {
				const uint32_t ul = cand * REF_BLOCK;
				units += sd.len ? (sd.len + ul - 1) / ul : 1;
			}
			const double per = units * o->nlv / S;
			const double eff = per / ceil(per) / (1.0 + 0.57 / cand);
			if (eff > best) { best = eff; unit_blocks = cand; }
		}
		if (ub && atoi(ub) > 0) unit_blocks = (uint32_t) atoi(ub);
	}
	if (sched_mode == 1) {
		static const char *ub = getenv("SAUGEN_UNIT_BLOCKS");
		if (ub && atoi(ub) > 0) unit_blocks = (uint32_t) atoi(ub);
	}
	plan_units(segs, o->units_tmp, unit_blocks);
	if (ahead && o->units_tmp.size() > o->unit_cap) return -2;
	if (o->units_tmp.size() > o->unit_cap) {
		uint32_t cap = o->unit_cap;
		while (cap < o->units_tmp.size()) cap *= 2;
		cudaStreamSynchronize(o->stream);
		o->d_units = (UnitDesc*) o->take(false, cap * sizeof(UnitDesc));
		o->h_units = (UnitDesc*) o->take(true, cap * sizeof(UnitDesc));
		o->slot[0].h_units = o->h_units;
		o->slot[1].h_units = (UnitDesc*) o->take(true, cap * sizeof(UnitDesc));
		if (!o->d_units || !o->h_units || !o->slot[1].h_units) {
			set_err("saugen_run: unit table growth", cudaGetLastError());
			return -1;
		}
		o->unit_cap = cap;
		o->compact = false;
	}
	const uint32_t nunits = (uint32_t) o->units_tmp.size();
	memcpy(sl.h_units, o->units_tmp.data(), nunits * sizeof(UnitDesc));
	/* render launches of this call: one, or one per hand-over cut (plan_call) */
	std::vector<uint32_t> gunit;              /* first unit of every launch, then nunits */
	gunit.push_back(0);
	for (uint32_t gs : o->group_first)
		for (uint32_t u = gunit.back(); u < nunits; ++u)
			if (o->units_tmp[u].seg >= gs) { if (u > gunit.back()) gunit.push_back(u); break; }
	gunit.push_back(nunits);
	const size_t ngroups = gunit.size() - 1;
	CallDesc &cd = *sl.h_call;
	cd.gen = o->d_desc; cd.call_len = (uint32_t) buf_len; cd.nseg = nseg; cd.seg_off = 0;
	cd.task_base = 0; cd.stereo = (stereo ? 1u : 0u) | (o->big_endian ? 2u : 0u);
	cd.unit_off = 0; cd.nunits = gunit[1]; cd.more_launches = ngroups > 1 ? 1u : 0u;
	cd.pcm = mode == 1 ? (int16_t*) sl.d_mix : sl.d_pcm;      /* (float planes in mode 1) */
	cd.status = sl.d_status;
	cudaError_t e;
	const bool prologue = o->compact && ngroups == 1 && nseg <= INLINE_SEGS && nunits <= INLINE_UNITS;
	if (prologue) {
		/* the usual call: its descriptors travel as kernel parameters of one small kernel that also
		 * zeroes [vlen][progress] + the slot's status and takes the run-ahead snapshot of the state
		 * -- no copy-engine work between the previous call's kernels and this call's */
		InlineCall ic;
		ic.cd = cd; ic.nseg = nseg; ic.nunits = nunits;
		memcpy(ic.segs, segs.data(), nseg * sizeof(SegDesc));
		memcpy(ic.units, o->units_tmp.data(), nunits * sizeof(UnitDesc));
		PrologueArgs pa;
		pa.d_call = o->d_call; pa.d_segs = o->d_segs; pa.d_units = o->d_units;
		pa.zero_a = (uint32_t*) o->d_vlen; pa.zero_a_words = (uint32_t) (o->zero_bytes / 4);
		pa.zero_b = sl.d_status; pa.zero_b_words = 1 + nseg;
		pa.snap_src = (const uint4*) o->d_ops; pa.snap_dst = (uint4*) o->d_snap;
		pa.snap_n16 = ahead ? (uint32_t) ((o->state_bytes + 15) / 16) : 0u;
		pa.zero_c = o->h_desc.team_hdr; pa.zero_c_words = multi > 1 ? (o->nlv ? o->nlv : 1) * 16u : 0u;
		e = launch_prologue(ic, pa, o->stream);
		o->counters[1]++;
	} else if (o->compact) {
		/* one copy in ([call][segs][units]), the memsets ([vlen][progress], [status]) */
		if (ahead) e = cudaMemcpyAsync(o->d_snap, o->d_ops, o->state_bytes, cudaMemcpyDeviceToDevice, o->stream);
		else e = cudaSuccess;
		if (e == cudaSuccess) e = cudaMemcpyAsync(o->d_call, sl.h_call, o->units_off_in_call + nunits * sizeof(UnitDesc),
				cudaMemcpyHostToDevice, o->stream);
		if (e == cudaSuccess) e = cudaMemsetAsync(o->d_vlen, 0, o->zero_bytes, o->stream);
		if (e == cudaSuccess) e = cudaMemsetAsync(sl.d_status, 0, (1 + nseg) * sizeof(uint32_t), o->stream);
	} else {
		if (ahead) e = cudaMemcpyAsync(o->d_snap, o->d_ops, o->state_bytes, cudaMemcpyDeviceToDevice, o->stream);
		else e = cudaSuccess;
		if (e != cudaSuccess) { set_err("saugen_run: launch", e); return -1; }
		e = cudaMemcpyAsync(o->d_segs, sl.h_segs, nseg * sizeof(SegDesc), cudaMemcpyHostToDevice, o->stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(o->d_units, sl.h_units, nunits * sizeof(UnitDesc), cudaMemcpyHostToDevice, o->stream);
		if (e == cudaSuccess) e = cudaMemsetAsync(o->d_vlen, 0, (size_t) nseg * (o->nlv ? o->nlv : 1) * sizeof(VoiceSeg), o->stream);
		if (e == cudaSuccess && ticketed_ctas) e = cudaMemsetAsync(o->d_progress, 0, ((size_t) o->nlv + 1) * sizeof(uint32_t), o->stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(o->d_call, sl.h_call, sizeof(CallDesc), cudaMemcpyHostToDevice, o->stream);
		if (e == cudaSuccess) e = cudaMemsetAsync(sl.d_status, 0, (1 + nseg) * sizeof(uint32_t), o->stream);
	}
	sl.timed = o->timing;
	if (sl.timed && !sl.ev_t[0])
		for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreate(&sl.ev_t[i]);
	if (!sl.done && e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
	if (!sl.mixed && e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.mixed, cudaEventDisableTiming);
	if (e == cudaSuccess && sl.timed) e = cudaEventRecord(sl.ev_t[0], o->stream);
	for (size_t gi = 0; gi < ngroups && e == cudaSuccess; ++gi) {
		if (gi > 0) {
			/* the next launch's unit range (pageable source: staged before the call returns) */
			CallDesc next = cd;
			next.unit_off = gunit[gi]; next.nunits = gunit[gi + 1] - gunit[gi];
			next.more_launches = gi + 1 < ngroups ? 1u : 0u;
			e = cudaMemcpyAsync(o->d_call, &next, sizeof(CallDesc), cudaMemcpyHostToDevice, o->stream);
			if (e == cudaSuccess && ticketed_ctas)
				e = cudaMemsetAsync(o->d_progress, 0, ((size_t) o->nlv + 1) * sizeof(uint32_t), o->stream);
			if (e != cudaSuccess) break;
		}
		if (multi > 1 && (gi > 0 || !prologue))
			e = cudaMemsetAsync(o->h_desc.team_hdr, 0, (size_t) (o->nlv ? o->nlv : 1) * 64, o->stream);
		if (e == cudaSuccess)
			e = launch_render(o->d_call, 1, o->d_segs, o->d_units, o->nlv, o->d_tables, o->d_coefs,
					shape.mask, o->nbufs, o->max_ops, o->nplan, warps, ticketed_ctas, sched_mode, shape.team, o->stream, multi);
		o->counters[0]++;
	}
	if (e == cudaSuccess && ngroups > 1) {
		/* the mix kernel reads the call's whole unit / segment tables */
		CallDesc all = cd;
		all.nunits = nunits; all.more_launches = 0;
		e = cudaMemcpyAsync(o->d_call, &all, sizeof(CallDesc), cudaMemcpyHostToDevice, o->stream);
	}
	if (e == cudaSuccess && sl.timed) e = cudaEventRecord(sl.ev_t[1], o->stream);
	if (e == cudaSuccess) {
		e = launch_mix(o->d_call, 1, o->d_segs, (uint32_t) buf_len, mode, o->stream);
		o->counters[1]++;
	}
	if (e == cudaSuccess && sl.timed) e = cudaEventRecord(sl.ev_t[2], o->stream);
	/* the read-back: on the copy stream, behind this call's kernels only */
	if (e == cudaSuccess) e = cudaEventRecord(sl.mixed, o->stream);
	if (e == cudaSuccess) e = cudaStreamWaitEvent(o->copy_stream, sl.mixed, 0);
	if (e == cudaSuccess) e = read_back(o, nseg, host_pcm_bytes, o->copy_stream, si);
	if (e == cudaSuccess) e = cudaEventRecord(sl.done, o->copy_stream);
	if (e != cudaSuccess) { set_err("saugen_run: launch", e); return -1; }
	sl.in_flight = true;
	sl.buf_len = buf_len; sl.stereo = stereo; sl.mode = mode; sl.host_bytes = host_pcm_bytes;
	sl.nseg = nseg;
	sl.gen_base = 0;
	for (uint32_t q = 0; q + 1 < nseg; ++q) sl.gen_base += segs[q].len;
	return 1;
}

/* After the slot's call has completed: out_len / return value (generator.c:938-972). */
static int finish_call(saugen_Generator *o, int si, size_t *out_len) {
	saugen_Generator::CallSlot &sl = o->slot[si];
	sl.in_flight = false;
	if (sl.timed) {
		float a = 0.f, b = 0.f;
		if (cudaEventElapsedTime(&a, sl.ev_t[0], sl.ev_t[1]) == cudaSuccess) o->render_ms += a;
		if (cudaEventElapsedTime(&b, sl.ev_t[1], sl.ev_t[2]) == cudaSuccess) o->mix_ms += b;
		sl.timed = false;
	}
	size_t gen_len = sl.gen_base;
	if (sl.nseg) gen_len += sl.h_status[1 + (sl.nseg - 1)];
	const bool alive = sl.h_status[0] != 0;
	const bool more = alive || sl.ev_after < o->ev_time.size();
	if (!more) {
		o->ended = true;
		if (out_len) *out_len = gen_len;
		return 0;
	}
	if (out_len) *out_len = sl.buf_len;
	return 1;
}

/* Undo a run-ahead call the caller did not ask for: wait for it, put the operator / voice state
 * and the host timeline back to where the last returned call left them. */
static void cancel_runahead(saugen_Generator *o) {
	if (!o->spec_valid) return;
	saugen_Generator::CallSlot &sl = o->slot[o->spec_slot];
	cudaEventSynchronize(sl.done);
	sl.in_flight = false; sl.timed = false;
	cudaMemcpyAsync(o->d_ops, o->d_snap, o->state_bytes, cudaMemcpyDeviceToDevice, o->stream);
	o->next_event = o->spec_next_event;
	o->cur_time = o->spec_cur_time;
	o->spec_valid = false;
	o->streak = -1;
	o->rows_stale = true;          /* the rows are the undone call's */
}

/* The generator's state as of the last RETURNED call: the live arrays, or the copy made before
 * a run-ahead call started on them. */
static const unsigned char *state_base(saugen_Generator *o) {
	return (const unsigned char*) (o->spec_valid ? o->d_snap : o->d_ops);
}

static bool runahead_enabled() {
	static const char *env = getenv("SAUGEN_RUNAHEAD");       /* 0 = off */
	return !(env && env[0] == '0');
}

/* sauGenerator_run's driver.  The call the caller asks for is either already in flight (the
 * run-ahead of the previous call: same length, channels and mode) or launched now; after it has
 * completed and while the caller consumes its PCM, the NEXT call is launched with the same
 * parameters (a player calls with one buffer size until the end, saugns.c:601-609), so the GPU
 * never waits for the host between calls.  Run-ahead starts once two
 * consecutive calls had the same parameters with no state inspection in between, and never
 * past the end of the signal.  Returns <0 error, 0 ended before this call, else 1 with *slot. */
static int run_call(saugen_Generator *o, size_t buf_len, int stereo, uint32_t mode, size_t host_bytes,
		int *slot_out, size_t *out_len, int *more_out) {
	if (!o) return -1;
	if (buf_len > o->row_len) { g_err = "saugen_run: buf_len exceeds max_call_len"; return -1; }
	cudaSetDevice(o->device);
	int si;
	if (o->spec_valid) {
		saugen_Generator::CallSlot &sp = o->slot[o->spec_slot];
		if (sp.buf_len == buf_len && sp.stereo == stereo && sp.mode == mode && sp.host_bytes == host_bytes) {
			si = o->spec_slot;
			o->spec_valid = false;
			o->rows_stale = false;
		} else {
			cancel_runahead(o);
			si = -1;
		}
	} else {
		si = -1;
	}
	if (si < 0) {
		if (o->ended) {
			cudaMemsetAsync(o->slot[o->cur_slot].d_pcm, 0, 2 * (size_t) o->row_len * sizeof(int16_t), o->stream);
			if (out_len) *out_len = 0;
			*more_out = 0;
			*slot_out = o->cur_slot;
			return 0;
		}
		const saugen_Generator::CallSlot &last = o->slot[o->cur_slot];
		o->streak = (last.buf_len == buf_len && last.stereo == stereo && last.mode == mode &&
				last.host_bytes == host_bytes) ? o->streak + 1 : 0;
		si = o->cur_slot ^ 1;
		if (launch_call(o, si, buf_len, stereo, mode, host_bytes) < 0) { if (out_len) *out_len = 0; return -1; }
		o->rows_stale = false;
	} else {
		o->streak++;
	}
	saugen_Generator::CallSlot &sl = o->slot[si];
	/* the call after this one goes out BEFORE this one is waited for: its state snapshot and
	 * its kernels queue up behind this call's on the stream, so the GPU runs on while the host
	 * finishes this call (should this call turn out to be the last, the extra one is undone) */
	if (o->streak >= 1 && runahead_enabled() && !o->d_tap) {
		o->spec_next_event = o->next_event;
		o->spec_cur_time = o->cur_time;
		if (launch_call(o, si ^ 1, buf_len, stereo, mode, host_bytes, true) >= 0) {
			o->spec_valid = true;
			o->spec_slot = si ^ 1;
			o->rows_stale = true;
		} else {
			o->next_event = o->spec_next_event;
			o->cur_time = o->spec_cur_time;
		}
	}
	cudaError_t e = cudaEventSynchronize(sl.done);
	if (e != cudaSuccess) { set_err("saugen_run", e); if (out_len) *out_len = 0; return -1; }
	o->cur_slot = si;
	*slot_out = si;
	const int more = finish_call(o, si, out_len);
	*more_out = more;
	if (!more && o->spec_valid) {
		cancel_runahead(o);
		o->ended = true;
	}
	return 1;
}

extern "C" int saugen_run(saugen_Generator *o, int16_t *buf, size_t buf_len, int stereo,
		size_t *out_len) {
	int more = 0, si = 0;
	const size_t bytes = buf_len * (stereo ? 2 : 1) * sizeof(int16_t);
	int r = run_call(o, buf_len, stereo, 0, bytes, &si, out_len, &more);
	if (r < 0) return r;
	if (r == 0) { if (buf) memset(buf, 0, bytes); return 0; }
	if (buf) memcpy(buf, o->slot[si].h_pcm, bytes);
	return more;
}

extern "C" int saugen_run_device(saugen_Generator *o, size_t buf_len, int stereo,
		int16_t **dev_pcm, size_t *out_len) {
	int more = 0, si = 0;
	int r = run_call(o, buf_len, stereo, 0, 0, &si, out_len, &more);
	if (r < 0) return r;
	if (dev_pcm) *dev_pcm = o->slot[si].d_pcm;
	if (r == 0) { cudaStreamSynchronize(o->stream); return 0; }
	return more;
}

extern "C" int saugen_run_mix(saugen_Generator *o, size_t buf_len, float **dev_mix, size_t *out_len) {
	int more = 0, si = 0;
	int r = run_call(o, buf_len, 1, 1, 0, &si, out_len, &more);
	if (r < 0) return r;
	if (dev_mix) *dev_mix = o->slot[si].d_mix;
	if (r == 0) {
		cudaMemsetAsync(o->slot[si].d_mix, 0, 2 * (size_t) o->row_len * sizeof(float), o->stream);
		cudaStreamSynchronize(o->stream);
		return 0;
	}
	return more;
}

extern "C" int saugen_mix_to_pcm(saugen_Generator *o, const float *dev_mix, size_t buf_len,
		int stereo, int16_t *host_buf) {
	if (!o || buf_len > o->row_len) return -1;
	cudaSetDevice(o->device);
	/* on a stream of its own: the launch stream and the copy stream may already hold the next call
	 * (run-ahead), and the conversion of THIS call's reduced planes must not wait behind it
	 * (mix-mode calls leave d_pcm alone) */
	if (!o->aux_stream) o->aux_stream = g_streams.get(o->device);
	cudaStream_t st = o->aux_stream ? o->aux_stream : o->stream;
	cudaError_t e = launch_planes_to_pcm(dev_mix, o->row_len, (uint32_t) buf_len,
			(stereo ? 1u : 0u) | (o->big_endian ? 2u : 0u),
			o->d_pcm, st);
	const size_t bytes = buf_len * (stereo ? 2 : 1) * sizeof(int16_t);
	if (e == cudaSuccess && host_buf)
		e = cudaMemcpyAsync(o->h_pcm, o->d_pcm, bytes, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) { set_err("saugen_mix_to_pcm", e); return -1; }
	if (host_buf) memcpy(host_buf, o->h_pcm, bytes);
	return 0;
}

/* ---- batched calls ----------------------------------------------------------- *
 * One render + one mix launch for n generators (same device and wave tables), in
 * two halves so that a driver can keep the GPU busy: begin() plans, uploads,
 * launches and queues the read-backs on the batch's stream without waiting; end()
 * waits, hands out the PCM and does the end-of-call bookkeeping.  While one batch
 * is between begin and end, the host works on another (saugns_b200/batch.py). */
struct saugen_Batch {
	int device = 0;
	cudaStream_t stream = nullptr;
	std::vector<CallDesc> calls;
	std::vector<SegDesc> segs;
	std::vector<UnitDesc> units;
	std::vector<size_t> call_of;                 // generator index per call
	std::vector<saugen_Generator*> gens;         // of the call in flight
	std::vector<int16_t*> bufs;
	CallDesc *d_calls = nullptr; size_t d_calls_cap = 0;
	SegDesc *d_segs = nullptr; size_t d_segs_cap = 0;
	UnitDesc *d_units = nullptr; size_t d_units_cap = 0;
	size_t buf_len = 0, bytes = 0;
	bool in_flight = false, dest_pinned = false;
};

extern "C" saugen_Batch *saugen_batch_create(int device) {
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
		set_err("saugen_batch_create: no such CUDA device", cudaGetLastError());
		return nullptr;
	}
	cudaSetDevice(device);
	saugen_Batch *b = new saugen_Batch();
	b->device = device;
	b->stream = g_streams.get(device);
	if (!b->stream) { set_err("saugen_batch_create: stream", cudaGetLastError()); delete b; return nullptr; }
	return b;
}

extern "C" void saugen_batch_destroy(saugen_Batch *b) {
	if (!b) return;
	cudaSetDevice(b->device);
	cudaStreamSynchronize(b->stream);
	if (b->d_calls) cudaFree(b->d_calls);
	if (b->d_segs) cudaFree(b->d_segs);
	if (b->d_units) cudaFree(b->d_units);
	g_streams.put(b->device, b->stream);
	delete b;
}

/* Pinned host memory from the library's pool: PCM destinations allocated here take
 * the device-to-host copy directly (saugen_batch_begin, dest_pinned). */
extern "C" void *saugen_pinned_alloc(size_t bytes) { return g_pool.alloc(true, 0, bytes ? bytes : 1); }
extern "C" void saugen_pinned_free(void *p) { g_pool.release(true, 0, p); }

extern "C" int saugen_batch_begin(saugen_Batch *b, saugen_Generator *const *gens, size_t n,
		int16_t *const *bufs, size_t buf_len, int stereo, int dest_pinned) {
	if (!b || b->in_flight) { g_err = "saugen_batch_begin: batch missing or still in flight"; return -1; }
	b->calls.clear(); b->segs.clear(); b->call_of.clear(); b->units.clear();
	b->gens.assign(gens, gens + n);
	b->bufs.assign(n, nullptr);
	if (bufs) for (size_t i = 0; i < n; ++i) b->bufs[i] = bufs[i];
	b->buf_len = buf_len;
	b->bytes = buf_len * (stereo ? 2 : 1) * sizeof(int16_t);
	b->dest_pinned = dest_pinned != 0;
	if (n == 0) return 0;
	cudaSetDevice(b->device);
	cudaStream_t st = b->stream;
	saugen_Generator *g0 = nullptr;
	uint32_t ntasks = 0, wave_mask = 0, nbufs = 1, max_ops = 1, nplan = 0;
	std::vector<std::vector<CallDesc>> later;    /* render launches after the first (hand-over cuts; rare) */
	/* every generator is checked BEFORE any timeline moves: a refused batch leaves all
	 * of them where they were */
	for (size_t i = 0; i < n; ++i) {
		saugen_Generator *o = gens[i];
		if (!o || o->ended) continue;
		if (!g0) g0 = o;
		if (buf_len > o->row_len || o->d_tables != g0->d_tables || o->device != b->device) {
			g_err = "saugen_run_many: generators must share device and tables and fit buf_len";
			return -1;
		}
	}
	for (size_t i = 0; i < n; ++i) {
		saugen_Generator *o = gens[i];
		if (o) cancel_runahead(o);
		if (o && o->ended && b->bufs[i]) memset(b->bufs[i], 0, b->bytes);   /* as saugen_run does */
		if (!o || o->ended) continue;
		plan_call(o, (uint32_t) buf_len, o->segs_tmp);
		if (!ensure_seg_cap(o, o->segs_tmp.size())) return -1;
		CallDesc cd;
		cd.gen = o->d_desc; cd.call_len = (uint32_t) buf_len; cd.nseg = (uint32_t) o->segs_tmp.size();
		cd.seg_off = (uint32_t) b->segs.size(); cd.task_base = ntasks;
		cd.stereo = (stereo ? 1u : 0u) | (o->big_endian ? 2u : 0u); cd.more_launches = 0;
		cd.pcm = o->d_pcm;
		cd.status = o->d_status;
		{
			saugen_Generator::CallSlot &sl = o->slot[0];
			sl.buf_len = buf_len; sl.stereo = stereo; sl.mode = 0; sl.host_bytes = b->bytes;
			sl.nseg = cd.nseg; sl.gen_base = 0; sl.timed = false; sl.ev_after = o->next_event;
			for (uint32_t q = 0; q + 1 < cd.nseg; ++q) sl.gen_base += o->segs_tmp[q].len;
			o->cur_slot = 0; o->streak = 0;
		}
		plan_units(o->segs_tmp, o->units_tmp, 1u << 20);
		cd.unit_off = (uint32_t) b->units.size(); cd.nunits = (uint32_t) o->units_tmp.size();
		b->units.insert(b->units.end(), o->units_tmp.begin(), o->units_tmp.end());
		if (!o->group_first.empty()) {
			/* hand-over cuts (plan_call): this call's later launches go into later rounds */
			uint32_t first = 0, round = 0;
			const uint32_t nu = cd.nunits;
			std::vector<uint32_t> cutu;
			for (uint32_t gs : o->group_first)
				for (uint32_t u = first; u < nu; ++u)
					if (o->units_tmp[u].seg >= gs) { if (u > first) { cutu.push_back(u); first = u; } break; }
			cutu.push_back(nu);
			first = cutu[0];
			for (size_t k = 1; k < cutu.size(); ++k) {
				CallDesc nx = cd;
				nx.unit_off = cd.unit_off + first; nx.nunits = cutu[k] - first;
				nx.more_launches = k + 1 < cutu.size() ? 1u : 0u;
				if (later.size() < ++round) later.emplace_back();
				later[round - 1].push_back(nx);
				first = cutu[k];
			}
			if (cutu.size() > 1) { cd.nunits = cutu[0]; cd.more_launches = 1; }
		}
		if (o->compact) cudaMemsetAsync(o->d_vlen, 0, o->zero_bytes0, st);
		else cudaMemsetAsync(o->d_vlen, 0, (size_t) cd.nseg * (o->nlv ? o->nlv : 1) * sizeof(VoiceSeg), st);
		b->segs.insert(b->segs.end(), o->segs_tmp.begin(), o->segs_tmp.end());
		b->calls.push_back(cd);
		b->call_of.push_back(i);
		ntasks += o->nlv;
		wave_mask |= o->wave_mask;
		if (o->nbufs > nbufs) nbufs = o->nbufs;
		if (o->max_ops > max_ops) max_ops = o->max_ops;
		if (o->nplan > nplan) nplan = o->nplan;
		if (!o->compact) cudaMemsetAsync(o->d_status, 0, (1 + cd.nseg) * sizeof(uint32_t), st);
	}
	if (b->calls.empty()) return 0;
	cudaError_t e = cudaSuccess;
	auto grow = [&](void **p, size_t *cap, size_t need, size_t elem) {
		if (e != cudaSuccess || need <= *cap) return;
		cudaStreamSynchronize(st);
		if (*p) cudaFree(*p);
		*cap = need * 2;
		e = cudaMalloc(p, *cap * elem);
	};
	size_t ncalls_all = b->calls.size();
	for (auto &r : later) {
		uint32_t tb = 0;
		for (CallDesc &c : r) {                     /* tasks of a round: its own calls' voices */
			c.task_base = tb;
			for (size_t k = 0; k < b->calls.size(); ++k)
				if (b->calls[k].gen == c.gen) { tb += gens[b->call_of[k]]->nlv; break; }
		}
		ncalls_all += r.size();
	}
	grow((void**) &b->d_calls, &b->d_calls_cap, ncalls_all, sizeof(CallDesc));
	grow((void**) &b->d_segs, &b->d_segs_cap, b->segs.size(), sizeof(SegDesc));
	grow((void**) &b->d_units, &b->d_units_cap, b->units.size(), sizeof(UnitDesc));
	if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_units, b->units.data(), b->units.size() * sizeof(UnitDesc), cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_calls, b->calls.data(), b->calls.size() * sizeof(CallDesc), cudaMemcpyHostToDevice, st);
	{
		size_t at = b->calls.size();
		for (auto &r : later) {
			if (e == cudaSuccess)       /* pageable source: staged before the call returns */
				e = cudaMemcpyAsync(b->d_calls + at, r.data(), r.size() * sizeof(CallDesc), cudaMemcpyHostToDevice, st);
			at += r.size();
		}
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_segs, b->segs.data(), b->segs.size() * sizeof(SegDesc), cudaMemcpyHostToDevice, st);
	/* kernel times of the batch accumulate on the first generator */
	saugen_Generator::CallSlot &t0 = g0->slot[0];
	t0.timed = g0->timing;
	if (t0.timed && !t0.ev_t[0])
		for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreate(&t0.ev_t[i]);
	if (e == cudaSuccess && t0.timed) e = cudaEventRecord(t0.ev_t[0], st);
	if (e == cudaSuccess) {
		const Shape shape = pick_shape(ntasks, wave_mask, nbufs, max_ops, nplan, g0->d_coefs != nullptr);
		if (!shape.warps) e = cudaErrorInvalidConfiguration;
		else e = launch_render(b->d_calls, (uint32_t) b->calls.size(), b->d_segs, b->d_units, ntasks, g0->d_tables,
				g0->d_coefs, shape.mask, nbufs, max_ops, nplan, shape.warps, 0, 0, shape.team, st);
		g0->counters[0]++;
		size_t at = b->calls.size();
		for (auto &r : later) {
			uint32_t nt = 0;
			for (size_t k = 0; k < b->calls.size(); ++k)
				for (const CallDesc &c : r) if (b->calls[k].gen == c.gen) nt += gens[b->call_of[k]]->nlv;
			if (e == cudaSuccess && shape.warps)
				e = launch_render(b->d_calls + at, (uint32_t) r.size(), b->d_segs, b->d_units, nt, g0->d_tables,
						g0->d_coefs, shape.mask, nbufs, max_ops, nplan, shape.warps, 0, 0, shape.team, st);
			g0->counters[0]++;
			at += r.size();
		}
	}
	if (e == cudaSuccess && t0.timed) e = cudaEventRecord(t0.ev_t[1], st);
	if (e == cudaSuccess) {
		e = launch_mix(b->d_calls, (uint32_t) b->calls.size(), b->d_segs, (uint32_t) buf_len, 0, st);
		g0->counters[1]++;
	}
	if (e == cudaSuccess && t0.timed) e = cudaEventRecord(t0.ev_t[2], st);
	for (size_t c = 0; c < b->calls.size() && e == cudaSuccess; ++c) {
		const size_t i = b->call_of[c];
		saugen_Generator *o = gens[i];
		int16_t *dst = b->bufs[i];
		if (dst && b->dest_pinned) {
			/* status to the generator's staging block, PCM straight to its destination */
			e = cudaMemcpyAsync(o->h_status, o->d_status, (1 + b->calls[c].nseg) * sizeof(uint32_t),
					cudaMemcpyDeviceToHost, st);
			if (e == cudaSuccess) e = cudaMemcpyAsync(dst, o->d_pcm, b->bytes, cudaMemcpyDeviceToHost, st);
		} else {
			e = read_back(o, b->calls[c].nseg, dst ? b->bytes : 0, st);
		}
	}
	if (e != cudaSuccess) { set_err("saugen_run_many", e); cudaStreamSynchronize(st); return -1; }
	b->in_flight = true;
	return 1;
}

extern "C" int saugen_batch_end(saugen_Batch *b, size_t *out_lens, int *more) {
	if (!b) return -1;
	const size_t n = b->gens.size();
	for (size_t i = 0; i < n; ++i) {
		if (out_lens) out_lens[i] = 0;
		if (more) more[i] = 0;
	}
	if (!b->in_flight) return 0;
	b->in_flight = false;
	cudaSetDevice(b->device);
	cudaError_t e = cudaStreamSynchronize(b->stream);
	if (e != cudaSuccess) { set_err("saugen_run_many", e); return -1; }
	/* pinned staging -> the callers' buffers: fresh destination pages fault on first
	 * touch, so a large batch is copied by a few threads */
	if (!b->dest_pinned) {
		const size_t nc = b->calls.size();
		const size_t bytes = b->bytes;
		auto copy_range = [b, bytes](size_t lo, size_t hi) {
			for (size_t c = lo; c < hi; ++c) {
				const size_t i = b->call_of[c];
				if (b->bufs[i]) memcpy(b->bufs[i], b->gens[i]->h_pcm, bytes);
			}
		};
		size_t nthr = nc * bytes >= ((size_t) 4 << 20) ? std::thread::hardware_concurrency() / 2 : 1;
		if (nthr > 8) nthr = 8;
		if (nthr < 2) copy_range(0, nc);
		else {
			std::vector<std::thread> th;
			for (size_t t = 0; t < nthr; ++t)
				th.emplace_back(copy_range, nc * t / nthr, nc * (t + 1) / nthr);
			for (auto &x : th) x.join();
		}
	}
	int any = 0;
	for (size_t c = 0; c < b->calls.size(); ++c) {
		const size_t i = b->call_of[c];
		size_t ol = 0;
		int m = finish_call(b->gens[i], 0, &ol);
		if (out_lens) out_lens[i] = ol;
		if (more) more[i] = m;
		any |= m;
	}
	return any;
}

/* == the two halves back to back, on a per-thread batch per device */
extern "C" int saugen_run_many(saugen_Generator *const *gens, size_t n, int16_t *const *bufs,
		size_t buf_len, int stereo, size_t *out_lens, int *more) {
	if (!gens || n == 0) return 0;
	int device = 0;
	for (size_t i = 0; i < n; ++i) if (gens[i]) { device = gens[i]->device; break; }
	struct Holder {
		std::map<int, saugen_Batch*> by_dev;
		~Holder() { for (auto &kv : by_dev) saugen_batch_destroy(kv.second); }
	};
	static thread_local Holder holder;
	saugen_Batch *&b = holder.by_dev[device];
	if (!b) b = saugen_batch_create(device);
	if (!b) return -1;
	const int r = saugen_batch_begin(b, gens, n, bufs, buf_len, stereo, 0);
	if (r < 0) return r;
	return saugen_batch_end(b, out_lens, more);
}

/* ---- introspection ------------------------------------------------------- */

static void view_line(saugen_LineView *d, const OpState *o, int li) {
	const LineState *s = &o->line[li];
	d->v0 = s->v0; d->vt = s->vt; d->pos = s->pos; d->end = s->end;
	d->type = LM_TYPE(o->lmeta[li]); d->flags = LM_FLAGS(o->lmeta[li]);
}

extern "C" int saugen_read_op(saugen_Generator *o, uint32_t op_id, saugen_OpView *out) {
	if (!o || op_id >= o->op_count) return -1;
	cudaSetDevice(o->device);
	OpState s;
	cudaStreamSynchronize(o->stream);
	o->streak = -1;                 /* inspected: not a streaming caller */
	if (cudaMemcpy(&s, (const OpState*) state_base(o) + op_id, sizeof s, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
	memset(out, 0, sizeof *out);
	out->inited = (s.flags & ON_INIT) != 0;
	if (!out->inited) return 0;
	out->type = s.type;
	out->flags = s.flags & (ON_INIT | ON_TIME_INF);   /* reference bits only */
	out->time = s.time;
	view_line(&out->amp, &s, LINE_AMP); view_line(&out->amp2, &s, LINE_AMP2);
	view_line(&out->pan, &s, LINE_PAN);
	if (s.type >= SAUABI_POPT_wave) {
		view_line(&out->freq, &s, LINE_FREQ); view_line(&out->freq2, &s, LINE_FREQ2);
		view_line(&out->pm_a, &s, LINE_PMA);
	}
	out->i0 = s.i0; out->i1 = s.i1; out->mode = s.mode;
	switch (s.type) {
	case SAUABI_POPT_wave:
		out->oscflags = s.oscflags; out->prev_Is = s.prev_Is;
		out->prev_s = s.prev_s; out->fb_s = s.fb_s; break;
	case SAUABI_POPT_raseg:
		out->oscflags = s.ras_flags | (s.ras_func << 16) | (s.ras_level << 24);
		out->prev_s = s.prev_s; out->fb_s = s.fb_s;
		out->alpha = s.ras_alpha; out->rate2x = s.oscflags & 1; break;
	}
	return 0;
}

extern "C" int saugen_read_voice(saugen_Generator *o, uint32_t vo_id, uint32_t out[4]) {
	if (!o || vo_id >= o->vo_count) return -1;
	cudaSetDevice(o->device);
	VoiceState s;
	cudaStreamSynchronize(o->stream);
	o->streak = -1;
	if (cudaMemcpy(&s, (const VoiceState*) (state_base(o) + ((const unsigned char*) o->d_voices - (const unsigned char*) o->d_ops)) + vo_id,
			sizeof s, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
	out[0] = s.duration; out[1] = s.flags; out[2] = s.carr_op; out[3] = 0;
	return 0;
}

extern "C" int saugen_read_voice_rows(saugen_Generator *o, uint32_t vo_id, float *s, float *r, size_t n) {
	if (!o || vo_id < o->voice_begin || vo_id >= o->voice_end || n > o->row_len) return -1;
	cudaSetDevice(o->device);
	o->streak = -1;
	if (o->rows_stale) {
		/* a run-ahead call has overwritten the last returned call's rows (a caller that inspects
		 * rows does so from the first calls on, before run-ahead starts; SAUGEN_RUNAHEAD=0 turns it off) */
		g_err = "saugen_read_voice_rows: the rows hold a run-ahead call (inspect before streaming, or SAUGEN_RUNAHEAD=0)";
		return -1;
	}
	cudaStreamSynchronize(o->stream);
	const size_t lv = vo_id - o->voice_begin;
	/* gather the voice's 512-byte pieces, one per frame tile (device_types.h:ROW_TILE) */
	const size_t ntile = (n + ROW_TILE - 1) / ROW_TILE;
	std::vector<float> tmp(ntile * ROW_TILE);
	const float *src[2] = {o->d_rows_s, o->d_rows_r};
	float *dst[2] = {s, r};
	for (int k = 0; k < 2; ++k) {
		if (!dst[k] || !ntile) continue;
		if (cudaMemcpy2D(tmp.data(), ROW_TILE * sizeof(float), src[k] + lv * ROW_TILE,
				(size_t) o->row_stride * sizeof(float), ROW_TILE * sizeof(float), ntile,
				cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
		memcpy(dst[k], tmp.data(), n * sizeof(float));
	}
	return 0;
}

extern "C" int saugen_counters(saugen_Generator *o, uint64_t out[4]) {
	if (!o) return -1;
	for (int i = 0; i < 4; ++i) out[i] = o->counters[i];
	out[2] = o->nbufs; out[3] = o->wave_mask;
	return 0;
}
/* Per-kernel device time: CUDA events on the launch stream around each kernel. */
extern "C" int saugen_set_timing(saugen_Generator *o, int on) {
	if (!o) return -1;
	/* (a call that ran ahead was launched under the old setting: it is undone and launched again, so that
	 * every call between two settings is timed, or none) */
	if ((on != 0) != o->timing) { cudaSetDevice(o->device); cancel_runahead(o); }
	o->timing = on != 0;                     /* (the events are made per call slot when first timed) */
	o->render_ms = o->mix_ms = 0.0;
	return 0;
}
extern "C" int saugen_kernel_ms(saugen_Generator *o, double out[2]) {
	if (!o) return -1;
	out[0] = o->render_ms; out[1] = o->mix_ms;
	return 0;
}
/* Device arithmetic self-test (kernels.cu:selftest_kernel): number of inputs on
 * which a fast-path primitive differs from the statement it replaces, <0 on error. */
extern "C" long long saugen_selftest(int device, const saugen_WaveTables *tables) {
	if (cudaSetDevice(device) != cudaSuccess) { set_err("saugen_selftest", cudaGetLastError()); return -1; }
	if (!tables) { g_err = "saugen_selftest: wave tables required"; return -1; }
	float *d_tab = get_device_tables(device, tables);
	unsigned long long *d_bad = nullptr, h_bad = 0;
	if (!d_tab || cudaMalloc(&d_bad, sizeof h_bad) != cudaSuccess) { set_err("saugen_selftest", cudaGetLastError()); return -1; }
	cudaMemset(d_bad, 0, sizeof h_bad);
	cudaError_t e = launch_selftest(d_tab, d_bad, 0);
	if (e == cudaSuccess) e = cudaMemcpy(&h_bad, d_bad, sizeof h_bad, cudaMemcpyDeviceToHost);
	cudaFree(d_bad);
	if (e != cudaSuccess) { set_err("saugen_selftest", e); return -1; }
	return (long long) h_bad;
}
/* developer aid (not in the header): out[0] = records, out[1..] = their codes (render_fast.cuh:rec_code) */
extern "C" int saugen_debug_signature(saugen_Generator *o, uint32_t out[36]) {
	if (!o) return -1;
	cudaSetDevice(o->device);
	cudaStreamSynchronize(o->stream);
	return read_signature_dump(out) == cudaSuccess ? 0 : -1;
}
/* developer aid (tests; not in the header): keep every operator's output buffer of the calls that
 * follow -- what the reference's run_block leaves in the operator's mix_buf (generator.c:664-730).
 * The voices then run on the general interpreter only (no steady plans, teams or run-ahead). */
extern "C" int saugen_debug_tap(saugen_Generator *o) {
	if (!o) return -1;
	cudaSetDevice(o->device);
	cancel_runahead(o);
	cudaStreamSynchronize(o->stream);
	if (!o->d_tap) {
		const size_t bytes = (size_t) (o->h_desc.op_count ? o->h_desc.op_count : 1) * o->row_len * sizeof(float);
		if (cudaMalloc(&o->d_tap, bytes) != cudaSuccess) return -1;
		cudaMemset(o->d_tap, 0, bytes);
		o->h_desc.tap = o->d_tap;
		if (cudaMemcpy(o->d_desc, &o->h_desc, sizeof(GenDesc), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
	}
	return 0;
}
extern "C" int saugen_debug_read_tap(saugen_Generator *o, uint32_t op_id, float *out, size_t n) {
	if (!o || !o->d_tap || op_id >= o->h_desc.op_count || n > o->row_len) return -1;
	cudaSetDevice(o->device);
	cudaStreamSynchronize(o->stream);
	return cudaMemcpy(out, o->d_tap + (size_t) op_id * o->row_len, n * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
/* developer aid (not in the header): render_team.cuh:g_team_dump */
extern "C" int saugen_debug_team(saugen_Generator *o, uint32_t out[32]) {
	if (!o) return -1;
	cudaSetDevice(o->device);
	cudaStreamSynchronize(o->stream);
	return read_team_dump(out) == cudaSuccess ? 0 : -1;
}
extern "C" float saugen_amp_scale(saugen_Generator *o) { return o ? o->amp_scale : 0.f; }
extern "C" const char *saugen_last_error(void) { return g_err.c_str(); }
extern "C" int saugen_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

/* The table file tools/make_wave_tables.py writes (libsau's host-built tables, verbatim):
 * "SAUT", version, 12, 2048, 12 x 2048 floats, amp_scale[12], amp_dc[12], phase_adj[12]. */
extern "C" saugen_WaveTables *saugen_wave_tables_load(const char *path) {
	FILE *f = path ? fopen(path, "rb") : nullptr;
	if (!f) { g_err = std::string("saugen_wave_tables_load: cannot open ") + (path ? path : "(null)"); return nullptr; }
	uint32_t hdr[4] = {0, 0, 0, 0};
	const size_t nt = (size_t) NUM_WAVES * WAVE_LEN;
	unsigned char *blk = (unsigned char*) malloc(sizeof(saugen_WaveTables) + nt * sizeof(float));
	saugen_WaveTables *t = (saugen_WaveTables*) blk;
	float *tab = blk ? (float*) (blk + sizeof(saugen_WaveTables)) : nullptr;
	bool ok = blk && fread(hdr, sizeof hdr, 1, f) == 1 && hdr[0] == 0x54554153u /* "SAUT" */ && hdr[1] == 1 &&
		hdr[2] == (uint32_t) NUM_WAVES && hdr[3] == (uint32_t) WAVE_LEN &&
		fread(tab, sizeof(float), nt, f) == nt &&
		fread(t->amp_scale, sizeof(float), NUM_WAVES, f) == (size_t) NUM_WAVES &&
		fread(t->amp_dc, sizeof(float), NUM_WAVES, f) == (size_t) NUM_WAVES &&
		fread(t->phase_adj, sizeof(int32_t), NUM_WAVES, f) == (size_t) NUM_WAVES;
	fclose(f);
	if (!ok) { free(blk); g_err = std::string("saugen_wave_tables_load: not a wave table file: ") + path; return nullptr; }
	for (int w = 0; w < NUM_WAVES; ++w) t->pilut[w] = tab + (size_t) w * WAVE_LEN;
	return t;
}
extern "C" void saugen_wave_tables_free(saugen_WaveTables *t) { free(t); }

extern "C" size_t saugen_abi_layout(uint32_t *out, size_t cap) {
	const uint32_t v[] = {
		sizeof(sauabi_Line), offsetof(sauabi_Line, v0), offsetof(sauabi_Line, vt),
		offsetof(sauabi_Line, pos), offsetof(sauabi_Line, end),
		offsetof(sauabi_Line, time_ms), offsetof(sauabi_Line, type), offsetof(sauabi_Line, flags),
		sizeof(sauabi_Time), offsetof(sauabi_Time, v_ms), offsetof(sauabi_Time, flags),
		sizeof(sauabi_RasOpt), offsetof(sauabi_RasOpt, line), offsetof(sauabi_RasOpt, alpha),
		sizeof(sauabi_ProgramIDArr), offsetof(sauabi_ProgramIDArr, ids),
		sizeof(sauabi_ProgramOpData),
		offsetof(sauabi_ProgramOpData, id), offsetof(sauabi_ProgramOpData, params),
		offsetof(sauabi_ProgramOpData, time), offsetof(sauabi_ProgramOpData, pan),
		offsetof(sauabi_ProgramOpData, amp), offsetof(sauabi_ProgramOpData, amp2),
		offsetof(sauabi_ProgramOpData, freq), offsetof(sauabi_ProgramOpData, freq2),
		offsetof(sauabi_ProgramOpData, pm_a), offsetof(sauabi_ProgramOpData, phase),
		offsetof(sauabi_ProgramOpData, seed), offsetof(sauabi_ProgramOpData, use_type),
		offsetof(sauabi_ProgramOpData, type), offsetof(sauabi_ProgramOpData, mode),
		offsetof(sauabi_ProgramOpData, camods), offsetof(sauabi_ProgramOpData, amods),
		offsetof(sauabi_ProgramOpData, ramods), offsetof(sauabi_ProgramOpData, fmods),
		offsetof(sauabi_ProgramOpData, rfmods), offsetof(sauabi_ProgramOpData, pmods),
		offsetof(sauabi_ProgramOpData, apmods), offsetof(sauabi_ProgramOpData, fpmods),
		sizeof(sauabi_ProgramEvent),
		offsetof(sauabi_ProgramEvent, wait_ms), offsetof(sauabi_ProgramEvent, vo_id),
		offsetof(sauabi_ProgramEvent, carr_op_id), offsetof(sauabi_ProgramEvent, op_count),
		offsetof(sauabi_ProgramEvent, op_data_count), offsetof(sauabi_ProgramEvent, op_list),
		offsetof(sauabi_ProgramEvent, op_data),
		sizeof(sauabi_Program),
		offsetof(sauabi_Program, events), offsetof(sauabi_Program, ev_count),
		offsetof(sauabi_Program, mode), offsetof(sauabi_Program, vo_count),
		offsetof(sauabi_Program, op_count), offsetof(sauabi_Program, op_nest_depth),
		offsetof(sauabi_Program, duration_ms), offsetof(sauabi_Program, ampmult),
		offsetof(sauabi_Program, name),
		SAUABI_WAVE_NAMED, SAUABI_LINE_NAMED, SAUABI_NOISE_NAMED, SAUABI_RAS_FUNCTIONS,
	};
	const size_t n = sizeof(v) / sizeof(v[0]);
	for (size_t i = 0; i < n && i < cap; ++i) out[i] = v[i];
	return n;
}
