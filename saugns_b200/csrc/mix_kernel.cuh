/* mix_kernel.cuh -- part of kernels.cu (one translation unit; included inside namespace saugen):
 * the fused mix-down + clip epilogue and the plane-to-PCM kernel. */
#pragma once

/* ---- mix + clip epilogue ------------------------------------------------- */

/* One CTA mixes one frame tile (ROW_TILE = 128 consecutive frames), one thread
 * per frame: the sum over voices must run in voice order in ONE thread (float
 * addition is not associative and mix_add adds voice after voice,
 * generator.c:773-786).  The tile's voice pieces lie side by side in HBM
 * (device_types.h:ROW_TILE), so the CTA reads ONE contiguous stream: the
 * producer warp moves MIX_TV voices (16 KiB) per stage with a single TMA bulk copy
 * (cp.async.bulk + mbarrier transaction count) into a ring of MIX_STAGES
 * stages, the four consumer warps add behind it.
 * A voice whose pan stands still contributes r = s * pan, computed here
 * (VoiceSeg); only moving pans have an r piece, which the consumers read straight
 * from HBM (a coalesced 128-byte line per warp; rare).
 * The producer also classifies each stage: when its MIX_TV voices all run through the
 * whole tile with pans standing still (MixSmem::slow == 0, the common case) the
 * consumers take a loop with unconditional loads and no per-voice decisions, whose
 * critical path is the two dependent additions per voice and channel (with one warp
 * per scheduler every exposed latency counts: the per-voice compare + predicated load
 * of the general loop held a CTA to 25 GB/s).  The bulk copies carry an L2
 * evict-first hint (render_ops.cuh:tma_bulk_g2s_stream). */
constexpr int MIX_FRAMES = ROW_TILE;           // = consumer threads (one per frame)
constexpr int MIX_TV = 32;                     // voices per stage (16 KiB per bulk copy)
constexpr int MIX_STAGES = 5;
constexpr int MIX_CWARPS = MIX_FRAMES / 32;    // consumer warps; one more warp produces
struct MixSmem {
	float s[MIX_STAGES][MIX_TV][MIX_FRAMES];
	uint2 vi[MIX_STAGES][MIX_TV];              // the tile's VoiceSeg records
	float pan[MIX_STAGES][MIX_TV];             // their (constant) pans, packed for 128-bit reads
	uint64_t full[MIX_STAGES], empty[MIX_STAGES];
	uint32_t slow[MIX_STAGES];                 // the stage has a moving pan or a voice ending inside the tile
};

__device__ __forceinline__ void mix_store(const GenDesc *g, const CallDesc *cd, uint32_t mode,
		uint32_t f, float L, float R) {
	if (mode == 1) {                  /* float planes: CallDesc::pcm is the call slot's plane block */
		float *mix = reinterpret_cast<float*>(cd->pcm);
		mix[f] = L;
		mix[g->row_len + f] = R;
		return;
	}
	/* CallDesc::stereo: bit 0 = two channels, bit 1 = big-endian samples (the AU stream
	 * of `saugns -o -`, player/sndfile.c:160-168: the byte swap folded into the epilogue) */
	const bool be = (cd->stereo & 2u) != 0;
	if (cd->stereo & 1u) {                                         /* generator.c:795-810 */
		L = sau::fclampf(L, -1.f, 1.f);
		R = sau::fclampf(R, -1.f, 1.f);
		uint32_t w = ((uint32_t) (uint16_t) (short) __float2int_rn(L * 32767.f)) |
			((uint32_t) (uint16_t) (short) __float2int_rn(R * 32767.f) << 16);
		if (be) w = __byte_perm(w, 0u, 0x2301);
		reinterpret_cast<uint32_t*>(cd->pcm)[f] = w;
	} else {                                                       /* generator.c:812-825 */
		float m = (L + R) * 0.5f;
		m = sau::fclampf(m, -1.f, 1.f);
		uint32_t w = (uint16_t) (short) __float2int_rn(m * 32767.f);
		if (be) w = __byte_perm(w, 0u, 0x3201);
		reinterpret_cast<uint16_t*>(cd->pcm)[f] = (uint16_t) w;
	}
}

__global__ void __launch_bounds__(MIX_FRAMES + 32)
mix_kernel(const CallDesc *calls, const SegDesc *segs, uint32_t mode /*0 pcm, 1 float planes*/) {
	extern __shared__ __align__(128) unsigned char mix_smem_raw[];
	MixSmem &sm = *reinterpret_cast<MixSmem*>(mix_smem_raw);
	const CallDesc *cd = &calls[blockIdx.y];
	const GenDesc *g = cd->gen;
	/* frame tiles in DESCENDING order of launch: the render kernel wrote the call's last tiles last, so the
	 * L2 still holds a good part of them (dirty); taking them first turns those reads into L2 hits */
	const uint32_t bx = gridDim.x - 1u - blockIdx.x;
	const uint32_t f0 = bx * MIX_FRAMES;
	if (f0 >= cd->call_len) return;
	const uint32_t tid = threadIdx.x;
	const bool producer = tid >= (uint32_t) MIX_FRAMES;          /* the last warp */
	const uint32_t f = f0 + (producer ? 0u : tid);
	const bool valid = !producer && f < cd->call_len;
	const uint32_t nlv = g->voice_end - g->voice_begin;
	const uint32_t tstride = g->row_stride;
	/* the segment holding each thread's frame */
	uint32_t si = 0;
	for (; si < cd->nseg; ++si) {
		const SegDesc sd = segs[cd->seg_off + si];
		if (f >= sd.start && f < sd.start + sd.len) break;
	}
	const bool in_seg = valid && si < cd->nseg;
	const uint32_t fi = in_seg ? f - segs[cd->seg_off + si].start : 0u;
	__shared__ uint32_t seg0, mixed, active;
	if (tid == 0) { seg0 = si; mixed = 0; active = 0; }
	__syncthreads();
	if (valid && si != seg0) mixed = 1;
	if (in_seg && fi < cd->status[1 + si]) active = 1;
	if (tid == 0) {
		for (int st = 0; st < MIX_STAGES; ++st) {
			mbar_init(&sm.full[st], 1);                /* the producer's arrive.expect_tx */
			mbar_init(&sm.empty[st], MIX_CWARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	float L = 0.f, R = 0.f;
	if (mixed || seg0 >= cd->nseg) {
		/* an event boundary inside these frames: each thread walks its own segment's
		 * voice list straight from global memory (rare) */
		if (in_seg && fi < cd->status[1 + si]) {
			const uint2 *vl = reinterpret_cast<const uint2*>(g->vlen + (size_t) si * nlv);
			for (uint32_t lv = 0; lv < nlv; ++lv) {
				const uint2 v = vl[lv];
				if (fi < v.x) {
					const size_t at = (size_t) lv * ROW_TILE + row_index(f, tstride);
					const float s = g->rows_s[at];
					const float r = (v.y == PAN_DYNAMIC) ? g->rows_r[at] : s * __uint_as_float(v.y);
					L = (L + s) - r;
					R = (R + s) + r;
				}
			}
		}
		if (valid) mix_store(g, cd, mode, f, L, R);
		return;
	}
	if (!active) {                       /* nothing was rendered for these frames */
		if (valid) mix_store(g, cd, mode, f, 0.f, 0.f);
		return;
	}
	const uint2 *vl = reinterpret_cast<const uint2*>(g->vlen + (size_t) seg0 * nlv);
	const uint32_t ntiles = (nlv + MIX_TV - 1) / MIX_TV;
	if (producer) {
		/* The producer warp: per stage, the VoiceSeg records (one lane each), then lane 0
		 * posts the transaction count and issues the bulk copies: the MIX_TV voices' s
		 * pieces are contiguous (one copy), r pieces only for moving pans. */
		const uint32_t lane = tid & 31u;
		const float *tile_s = g->rows_s + (size_t) bx * tstride;
		/* the records are fetched three stages ahead of their use (their L2 latency
		 * would otherwise sit in this loop's critical path) */
		auto fetch = [&](uint32_t t) {
			const uint32_t v = t * MIX_TV + lane;
			return (lane < (uint32_t) MIX_TV && v < nlv) ? __ldg(vl + v) : make_uint2(0u, 0u);
		};
		uint2 pre0 = fetch(0), pre1 = fetch(1), pre2 = fetch(2);
		/* the tile's last frame, counted from the segment start: a voice running beyond it
		 * contributes to every frame of the tile */
		const uint32_t last = cd->call_len - 1u - f0 < (uint32_t) (MIX_FRAMES - 1) ?
			cd->call_len - 1u - f0 : (uint32_t) (MIX_FRAMES - 1);
		const uint32_t fi_last = f0 - segs[cd->seg_off + seg0].start + last;
		for (uint32_t t = 0; t < ntiles; ++t) {
			const uint32_t st = t % MIX_STAGES, v0 = t * MIX_TV;
			const uint32_t nv = nlv - v0 < (uint32_t) MIX_TV ? nlv - v0 : (uint32_t) MIX_TV;
			const uint2 info = pre0;
			pre0 = pre1; pre1 = pre2; pre2 = fetch(t + 3);
			if (t >= (uint32_t) MIX_STAGES) mbar_wait(&sm.empty[st], ((t / MIX_STAGES) - 1u) & 1u);
			const bool has = lane < nv;
			if (has) { sm.vi[st][lane] = info; sm.pan[st][lane] = __uint_as_float(info.y); }
			const uint32_t slowmask = __ballot_sync(FULL, !has || info.y == PAN_DYNAMIC || info.x <= fi_last);
			if (lane == 0) sm.slow[st] = slowmask;
			__syncwarp();                      /* vi, pan, slow written before lane 0's arrive publishes them */
			if (lane == 0) {
				const uint32_t piece = ROW_TILE * (uint32_t) sizeof(float);
				mbar_expect_tx(&sm.full[st], nv * piece);
				tma_bulk_g2s_stream(&sm.s[st][0][0], tile_s + (size_t) v0 * ROW_TILE, nv * piece, &sm.full[st]);
			}
		}
		return;
	}
	const uint32_t fx = in_seg ? fi : 0xffffffffu;               /* frames outside take nothing */
	const float *tile_r = g->rows_r + (size_t) bx * tstride + tid;
	for (uint32_t t = 0; t < ntiles; ++t) {
		const uint32_t st = t % MIX_STAGES, v0 = t * MIX_TV;
		const uint32_t nv = nlv - v0 < (uint32_t) MIX_TV ? nlv - v0 : (uint32_t) MIX_TV;
		mbar_wait(&sm.full[st], (t / MIX_STAGES) & 1u);
		const float *sp = &sm.s[st][0][tid];
		const float *rp = tile_r + (size_t) v0 * ROW_TILE;
		const uint2 *ip = &sm.vi[st][0];
		if (sm.slow[st] == 0) {
			/* the common stage: MIX_TV voices that all run through the whole tile with pans
			 * standing still -- unconditional loads, no per-voice decisions; the two
			 * dependent additions per voice and channel are the critical path */
			const float4 *pp = reinterpret_cast<const float4*>(&sm.pan[st][0]);
#pragma unroll
			for (int k = 0; k < MIX_TV; k += 4) {
				const float4 pan = pp[k >> 2];
				const float s0 = sp[(k + 0) * MIX_FRAMES], s1 = sp[(k + 1) * MIX_FRAMES];
				const float s2 = sp[(k + 2) * MIX_FRAMES], s3 = sp[(k + 3) * MIX_FRAMES];
				const float r0 = s0 * pan.x, r1 = s1 * pan.y, r2 = s2 * pan.z, r3 = s3 * pan.w;
				L = (L + s0) - r0; R = (R + s0) + r0;          /* as compiled, Appendix B.3 */
				L = (L + s1) - r1; R = (R + s1) + r1;
				L = (L + s2) - r2; R = (R + s2) + r2;
				L = (L + s3) - r3; R = (R + s3) + r3;
			}
		} else {
			for (uint32_t k = 0; k < nv; ++k) {
				const uint2 info = ip[k];
				const bool on = fx < info.x;
				const float s = on ? sp[k * MIX_FRAMES] : 0.f;
				float rr;
				if (info.y == PAN_DYNAMIC) rr = on ? rp[k * MIX_FRAMES] : 0.f;
				else rr = s * __uint_as_float(info.y);
				L = (L + s) - rr;
				R = (R + s) + rr;
			}
		}
		__syncwarp();
		if ((tid & 31u) == 0) mbar_arrive(&sm.empty[st]);        /* this warp is done with the stage */
	}
	if (valid) mix_store(g, cd, mode, f, L, R);
}

/* float planes (already reduced over ranks) -> int16, for voice-sharded runs */
__global__ void planes_to_pcm_kernel(const float *mix, uint32_t plane_stride, uint32_t n,
		uint32_t stereo, int16_t *pcm) {
	const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n) return;
	float L = mix[f], R = mix[plane_stride + f];
	const bool be = (stereo & 2u) != 0;          /* flags as CallDesc::stereo */
	uint16_t *out = reinterpret_cast<uint16_t*>(pcm);
	auto put = [be](uint16_t *p, int v) {
		const uint16_t u = (uint16_t) (short) v;
		*p = be ? (uint16_t) ((u << 8) | (u >> 8)) : u;
	};
	if (stereo & 1u) {
		L = sau::fclampf(L, -1.f, 1.f);
		R = sau::fclampf(R, -1.f, 1.f);
		put(out + 2 * f, __float2int_rn(L * 32767.f));
		put(out + 2 * f + 1, __float2int_rn(R * 32767.f));
	} else {
		float m = sau::fclampf((L + R) * 0.5f, -1.f, 1.f);
		put(out + f, __float2int_rn(m * 32767.f));
	}
}
