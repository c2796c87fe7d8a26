/* kernels.cu -- sm_100a kernels of the saugns generator back end.
 *
 *   render_kernel : one warp per (call, voice).  The warp owns its voice's
 *     timeline for the call: it applies the voice's due events
 *     (handle_event, sau/generator.c:348-377), then renders every inter-event
 *     segment by interpreting the voice's bytecode (the flattened run_block
 *     recursion, generator.c:448-729): in 128-sample chunks through the
 *     general interpreter (run_chunk: any operator, any state), or -- for
 *     stretches of whole 1024-sample blocks in which nothing changes shape,
 *     almost all of a render -- from a PLAN the warp compiles once per stretch
 *     into its shared memory (steady_plan -> run_block_fast -> steady_update).
 *     Each lane owns 4 consecutive samples of a chunk; integer phase
 *     accumulation is a warp shuffle scan over the rounded increments, or a
 *     closed form when the frequency is uniform (bit-exact with
 *     sauPhasor_fill / sauCyclor_fill either way); self-PM operators run as a
 *     serial loop on one lane with their state in registers.  Wave tables --
 *     or their per-index cubic coefficients -- are staged into shared memory
 *     with TMA bulk copies (cp.async.bulk + mbarrier).  The carrier chunk,
 *     scaled, goes to the voice's 512-byte piece of the frame tile in HBM
 *     (device_types.h:ROW_TILE) with 128-bit streaming stores.
 *   mix_kernel : fused epilogue.  One CTA per frame tile, one thread per
 *     output frame; the tile's voice pieces stream through a TMA-filled ring
 *     and are summed in voice order (same float summation order as mix_add,
 *     generator.c:749-788), clamped and rounded to int16
 *     (mix_write_stereo/mono, generator.c:795-825).
 *
 * Compiled with -fmad=false: every float/double operation is a separately
 * rounded IEEE operation, in the order sau_arith.h states.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include "device_types.h"
#include "sau_arith.h"
#include "../../include/sau_program_abi.h"

namespace saugen {

#define FULL 0xffffffffu

/* per-warp shared memory: operator-state cache, work buffers, len stack */
constexpr int FAST_NS = 4;            // samples per lane in the steady-block fast path (8 measured
                                      // slower: its registers allow 16 resident warps per SM, 4 allows 32)
constexpr uint32_t BUF_FLOATS = 32 * (FAST_NS > SPL ? FAST_NS : SPL);   // per work buffer
/* The len stacks (general interpreter, live inside one chunk only) and the block
 * plan (fast path, live across one steady block, 32 bytes per record) share one
 * area: whichever is larger. */
constexpr int WIDE_WARPS = 28;        // most warps of a render_kernel_wide CTA: 72 registers each
constexpr uint32_t STACK_BYTES = 3 * MAX_NEST * (uint32_t) sizeof(uint32_t);
/* team > 1 (render_team.cuh): a second set of operator states (the member's work copy), a
 * second plan area (its executable plan) and the team's command block */
__host__ __device__ inline uint32_t warp_plan_bytes(uint32_t nplan) {
	const uint32_t plan_bytes = nplan * 32u;
	return plan_bytes > STACK_BYTES ? plan_bytes : STACK_BYTES;
}
/* a member's executable plan: the master's records + a save per cache slot */
__host__ __device__ inline uint32_t team_plan_bytes(uint32_t nplan) {
	return warp_plan_bytes(nplan) + TEAM_SLOTS * 32u;
}
/* one warp per voice: operator states, work buffers, plan / len stacks */
__host__ __device__ inline uint32_t warp_smem_bytes(uint32_t nbufs, uint32_t nslots, uint32_t nplan) {
	return nslots * (uint32_t) sizeof(OpState) + nbufs * BUF_FLOATS * (uint32_t) sizeof(float) + warp_plan_bytes(nplan);
}
/* a team of warps per voice: the voice's part (its operator states, the master plan / the leader's len
 * stacks, the command block) and, per member, a work copy of the operator states, work buffers and the
 * member's executable plan */
__host__ __device__ inline uint32_t team_lead_bytes(uint32_t nslots, uint32_t nplan) {
	return nslots * (uint32_t) sizeof(OpState) + warp_plan_bytes(nplan) + TEAM_CMD_BYTES;
}
__host__ __device__ inline uint32_t team_member_bytes(uint32_t nbufs, uint32_t nslots, uint32_t nplan) {
	return nslots * (uint32_t) sizeof(OpState) + nbufs * BUF_FLOATS * (uint32_t) sizeof(float) + team_plan_bytes(nplan);
}
__host__ __device__ inline uint32_t team_smem_bytes(uint32_t nbufs, uint32_t nslots, uint32_t nplan, uint32_t team) {
	return team_lead_bytes(nslots, nplan) + team * team_member_bytes(nbufs, nslots, nplan);
}

#include "render_ops.cuh"
#include "render_interp.cuh"
#include "render_plan.cuh"
#include "render_fast.cuh"
#include "render_team.cuh"
#include "render_kernel.cuh"
#include "mix_kernel.cuh"

/* ---- arithmetic self-test ------------------------------------------------ *
 * The hand-expanded primitives of the fast path against the plain statements
 * they replace: div_scale_by_int vs IEEE `/` for every non-zero int32 divisor
 * (strided over the grid) and each wave's diff_scale; ftoi_lo32 vs the low
 * word of sau::ftoi64 over a float bit-pattern sweep. */
__global__ void selftest_kernel(const float *tables, unsigned long long *bad) {
	const WaveCoeffs *wc = reinterpret_cast<const WaveCoeffs*>(tables + NUM_WAVES * WAVE_LEN);
	const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t nth = (uint64_t) gridDim.x * blockDim.x;
	unsigned long long nbad = 0;
	for (uint64_t u = tid; u < 0x100000000ull; u += nth) {
		const int32_t d = (int32_t) (uint32_t) u;
		if (d != 0) {
			const uint32_t w = (uint32_t) (u % NUM_WAVES);
			const float ds = wc->diff_scale[w];
			const float want = ds / (float) d;
			const float got = div_scale_by_int(ds, d);
			if (__float_as_uint(want) != __float_as_uint(got)) ++nbad;
		}
		const float x = __uint_as_float((uint32_t) u);
		if ((uint32_t) sau::ftoi64(x) != ftoi_lo32(x)) ++nbad;
	}
	if (nbad) atomicAdd(bad, nbad);
}

/* multiprocessors of the current device (queried once per device) */
int device_sm_count() {
	static std::mutex mu;
	static int sms[64] = {0};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) dev = 0;
	std::lock_guard<std::mutex> lk(mu);
	if (!sms[dev]) {
		int n = 0;
		if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 1;
		sms[dev] = n;
	}
	return sms[dev];
}
/* the most dynamic shared memory one CTA may opt in to on the current device */
size_t device_smem_optin() {
	static std::mutex mu;
	static size_t cap[64] = {0};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) dev = 0;
	std::lock_guard<std::mutex> lk(mu);
	if (!cap[dev]) {
		int n = 0;
		if (cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || n <= 0) n = 48 * 1024;
		cap[dev] = (size_t) n;
	}
	return cap[dev];
}

cudaError_t launch_selftest(const float *d_tables, unsigned long long *d_bad, cudaStream_t stream) {
	selftest_kernel<<<device_sm_count() * 8, 256, 0, stream>>>(d_tables, d_bad);
	return cudaGetLastError();
}

/* ---- host-callable launchers -------------------------------------------- */

/* wave_mask may carry CTAB_FLAG (coefficient tables in shared memory) */
size_t render_smem_bytes(uint32_t wave_mask, uint32_t nbufs, uint32_t nslots_ops, uint32_t nplan,
		uint32_t warps, uint32_t team) {
	uint32_t nslots = 0;
	for (uint32_t w = 0; w < NUM_WAVES; ++w) if (wave_mask & (1u << w)) ++nslots;
	const size_t slot = (wave_mask & CTAB_FLAG) ? CTAB_WAVE_BYTES : TAB_STRIDE * sizeof(float);
	if (team > 1u)
		return 128 + (size_t) nslots * slot + (size_t) (warps / team) * team_smem_bytes(nbufs, nslots_ops, nplan, team);
	return 128 + (size_t) nslots * slot + (size_t) warps * warp_smem_bytes(nbufs, nslots_ops, nplan);
}

/* ---- per-index cubic coefficients of every wave table -------------------- *
 * sauWave_get_herp (wave.h:127-141, as compiled: sau::herp_poly) forms c1, c2,
 * c3 from the four taps around an index before it touches the phase fraction;
 * they depend on the index alone.  One thread per (wave, index) evaluates the
 * SAME expressions once; the fast path then loads {c3, c2} (doubles) with one
 * 128-bit and {c1, c0} with one 64-bit shared-memory load.  c1 = 0.5 * (s2 - s0)
 * and c0 = s1 are float values held as floats (24 bytes per index instead of
 * 32: two waves leave room for 32 warps per SM); the kernel checks that the
 * float form of c1 is exact and flags the table set otherwise (the runtime then
 * stays with the float tables).  Layout per wave: 2048 x {c3, c2}, then
 * 2048 x {c1, c0}. */
__global__ void coef_kernel(const float *tables, double *coefs, uint32_t *inexact) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= NUM_WAVES * WAVE_LEN) return;
	const uint32_t w = t / WAVE_LEN, i = t % WAVE_LEN;
	const float *lut = tables + w * WAVE_LEN;
	const float s0 = lut[(i - 1) & sau::WAVE_LENMASK], s1 = lut[i];
	const float s2 = lut[(i + 1) & sau::WAVE_LENMASK], s3 = lut[(i + 2) & sau::WAVE_LENMASK];
	double c1, c2, c3;
	sau::herp_coefs(s0, s1, s2, s3, &c1, &c2, &c3);
	unsigned char *base = reinterpret_cast<unsigned char*>(coefs) + (size_t) w * CTAB_WAVE_BYTES;
	reinterpret_cast<double2*>(base)[i] = make_double2(c3, c2);
	const float c1f = (float) c1;
	if ((double) c1f != c1) atomicOr(inexact, 1u);
	reinterpret_cast<float2*>(base + CTAB_PLANE_BYTES)[i] = make_float2(c1f, s1);
}
size_t coef_table_bytes() { return (size_t) NUM_WAVES * CTAB_WAVE_BYTES; }
cudaError_t launch_coefs(const float *d_tables, double *d_coefs, uint32_t *d_inexact, cudaStream_t stream) {
	coef_kernel<<<(NUM_WAVES * WAVE_LEN + 255) / 256, 256, 0, stream>>>(d_tables, d_coefs, d_inexact);
	return cudaGetLastError();
}

/* grid: one warp per voice task, or a persistent grid of `ticketed_ctas` CTAs with
 * sched_mode 1 = (unit, voice) tickets, 2 = balanced contiguous ranges */
/* The opt-in shared-memory size of a kernel is set ONCE per device, to the device's
 * maximum, under a lock: batch driver threads launch concurrently on one device, and a
 * per-size cache could let one thread lower what another had just raised. */
static cudaError_t ensure_smem(bool wide, size_t smem) {
	static std::mutex mu;
	static bool configured[2][64] = {{false}, {false}};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
	if (smem > device_smem_optin()) return cudaErrorInvalidValue;
	std::lock_guard<std::mutex> lk(mu);
	if (!configured[wide][dev]) {
		const int cap = (int) device_smem_optin();
		cudaError_t e = wide ?
			cudaFuncSetAttribute(render_kernel_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, cap) :
			cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
		if (e != cudaSuccess) return e;
		configured[wide][dev] = true;
	}
	return cudaSuccess;
}
uint32_t plan_area_bytes(uint32_t nplan) { return warp_plan_bytes(nplan); }
int render_ctas_per_sm(size_t smem, uint32_t warps) {
	int n = 0;
	if (ensure_smem(warps > 8, smem) != cudaSuccess) return 0;
	cudaError_t e = warps > 8 ?
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, render_kernel_wide, (int) warps * 32, smem) :
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, render_kernel, (int) warps * 32, smem);
	return e == cudaSuccess ? n : 0;
}
cudaError_t launch_render(const CallDesc *d_calls, uint32_t ncalls, const SegDesc *d_segs,
		const UnitDesc *d_units, uint32_t ntasks, const float *d_tables, const double *d_coefs,
		uint32_t wave_mask, uint32_t nbufs, uint32_t nslots_ops, uint32_t nplan, uint32_t warps,
		uint32_t ticketed_ctas, uint32_t sched_mode, uint32_t team, cudaStream_t stream, uint32_t multi) {
	if (ntasks == 0) return cudaSuccess;
	if (nslots_ops == 0) nslots_ops = 1;
	if (team < 1 || ticketed_ctas) team = 1;
	const size_t smem = render_smem_bytes(wave_mask, nbufs, nslots_ops, nplan, warps, team);
	const bool wide = warps > 8;
	{
		cudaError_t e = ensure_smem(wide, smem);
		if (e != cudaSuccess) return e;
	}
	const uint32_t per_cta = warps / team;      /* voices per CTA */
	uint32_t grid = ticketed_ctas ? ticketed_ctas : (ntasks + per_cta - 1) / per_cta;
	if (multi > 1u && team > 1u && per_cta == 1u) {
		/* a team per voice over `multi` CTAs: they wait for one another through global memory, so all of
		 * them have to be resident at once -- a cooperative launch (it fails rather than hang) */
		grid *= multi;
		uint32_t team_arg = team | multi << 8, sched_arg = 0u;
		void *args[] = {&d_calls, &ncalls, &d_segs, &d_units, &ntasks, &d_tables, &d_coefs, &wave_mask, &nbufs,
			&nslots_ops, &nplan, &warps, &sched_arg, &team_arg};
		cudaError_t e = cudaLaunchCooperativeKernel(wide ? (const void*) render_kernel_wide : (const void*) render_kernel,
				dim3(grid), dim3(warps * 32), args, smem, stream);
		if (e == cudaSuccess) return e;
		/* refused (the device cannot hold the grid at once, e.g. under MPS): one CTA per voice */
		cudaGetLastError();
		grid /= multi;
	}
	if (wide)
		render_kernel_wide<<<grid, warps * 32, smem, stream>>>(d_calls, ncalls, d_segs, d_units, ntasks,
				d_tables, d_coefs, wave_mask, nbufs, nslots_ops, nplan, warps, ticketed_ctas ? sched_mode : 0u, team);
	else
		render_kernel<<<grid, warps * 32, smem, stream>>>(d_calls, ncalls, d_segs, d_units, ntasks,
				d_tables, d_coefs, wave_mask, nbufs, nslots_ops, nplan, warps, ticketed_ctas ? sched_mode : 0u, team);
	return cudaGetLastError();
}

/* see device_types.h:PrologueArgs */
__global__ void prologue_kernel(const __grid_constant__ InlineCall ic, const __grid_constant__ PrologueArgs a) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
	if (blockIdx.x == 0) {
		if (threadIdx.x == 0) *a.d_call = ic.cd;
		for (uint32_t i = threadIdx.x; i < ic.nseg; i += blockDim.x) a.d_segs[i] = ic.segs[i];
		for (uint32_t i = threadIdx.x; i < ic.nunits; i += blockDim.x) a.d_units[i] = ic.units[i];
	}
	for (uint32_t i = t; i < a.zero_a_words; i += nt) a.zero_a[i] = 0u;
	for (uint32_t i = t; i < a.zero_b_words; i += nt) a.zero_b[i] = 0u;
	for (uint32_t i = t; i < a.zero_c_words; i += nt) a.zero_c[i] = 0u;
	for (uint32_t i = t; i < a.snap_n16; i += nt) a.snap_dst[i] = a.snap_src[i];
}
cudaError_t launch_prologue(const InlineCall &ic, const PrologueArgs &a, cudaStream_t stream) {
	const uint32_t work = a.snap_n16 + (a.zero_a_words + a.zero_b_words) / 4u;
	uint32_t blocks = (work + 1023u) / 1024u;           /* ~4 items per thread */
	const uint32_t cap = 2u * (uint32_t) device_sm_count();
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
	prologue_kernel<<<blocks, 256, 0, stream>>>(ic, a);
	return cudaGetLastError();
}

cudaError_t launch_mix(const CallDesc *d_calls, uint32_t ncalls, const SegDesc *d_segs,
		uint32_t max_call_len, uint32_t mode, cudaStream_t stream) {
	if (ncalls == 0 || max_call_len == 0) return cudaSuccess;
	dim3 grid((max_call_len + MIX_FRAMES - 1) / MIX_FRAMES, ncalls);
	{
		static std::mutex mu;
		static bool configured[64] = {false};
		int dev = 0;
		cudaGetDevice(&dev);
		if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
		std::lock_guard<std::mutex> lk(mu);
		if (!configured[dev]) {
			cudaError_t e = cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
					(int) sizeof(MixSmem));
			if (e != cudaSuccess) return e;
			configured[dev] = true;
		}
	}
	mix_kernel<<<grid, MIX_FRAMES + 32, sizeof(MixSmem), stream>>>(d_calls, d_segs, mode);
	return cudaGetLastError();
}

/* developer aid: the lowered-plan signature of the first stretch of the launch's first voice */
cudaError_t read_signature_dump(uint32_t out[36]) {
	return cudaMemcpyFromSymbol(out, g_sig_dump, sizeof(uint32_t) * 36);
}
cudaError_t read_team_dump(uint32_t out[32]) {
	return cudaMemcpyFromSymbol(out, g_team_dump, sizeof(uint32_t) * 32);
}

cudaError_t launch_planes_to_pcm(const float *d_mix, uint32_t plane_stride, uint32_t n,
		uint32_t stereo, int16_t *d_pcm, cudaStream_t stream) {
	if (!n) return cudaSuccess;
	planes_to_pcm_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_mix, plane_stride, n, stereo, d_pcm);
	return cudaGetLastError();
}

} // namespace saugen
