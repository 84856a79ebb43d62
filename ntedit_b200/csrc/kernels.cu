// sm_100a kernels of the ntEdit hot path -- see kernels.cuh for the map to the reference.
#include "kernels.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace ntb {

// ------------------------------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP), cache-hinted loads
__device__ __forceinline__ uint32_t
smem_addr(const void* p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void
mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void
mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void
mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile("{\n"
	             ".reg .pred p;\n"
	             "NTB_WAIT:\n"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	             "@p bra NTB_DONE;\n"
	             "bra NTB_WAIT;\n"
	             "NTB_DONE:\n"
	             "}\n" ::"r"(smem_addr(bar)),
	             "r"(parity)
	             : "memory");
}

__device__ __forceinline__ void
bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
	             : "memory");
}

// filter probes are uniformly random over a multi-GB array: read-only path, do not allocate in L1
__device__ __forceinline__ uint32_t
ld_filter_u8(const uint8_t* p)
{
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

// Resident CTAs per SM of a kernel on the CURRENT device (after raising its dynamic shared memory limit to `smem`).  cudaFuncSetAttribute is per device (and per
// context), so the result is cached per (kernel, device): a process that drives several GPUs from several host threads
// (ntedit-b200 --gpus N) sets the attribute on each of them.
constexpr int NTB_MAX_DEVICES = 64;
struct OccCache
{
	std::mutex lock;
	int per_sm[NTB_MAX_DEVICES] = {};
};

template<class Kernel>
static cudaError_t
walker_occupancy(Kernel kernel, OccCache& cache, size_t smem, int threads, int* out)
{
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) {
		return e;
	}
	if (dev < 0 || dev >= NTB_MAX_DEVICES) {
		return cudaErrorInvalidDevice;
	}
	std::lock_guard<std::mutex> guard(cache.lock);
	if (cache.per_sm[dev] == 0) {
		if (smem > 0) {
			e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (e != cudaSuccess) {
				return e;
			}
		}
		int n = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
		if (e != cudaSuccess) {
			return e;
		}
		cache.per_sm[dev] = n > 0 ? n : 1;
	}
	*out = cache.per_sm[dev];
	return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------------------------
// K1
// class byte of a text byte: bits 0-2 forward seed code, bits 3-5 reverse seed code (btllib's SEED_TAB[c & 7] path), bit 6
// accepted (ntedit.cpp:493-499 after toupper)
__device__ __forceinline__ uint8_t
class_of(unsigned c)
{
	const unsigned fc = base_code((unsigned char)c);
	const unsigned rc = rev_code((unsigned char)c);
	return (uint8_t)(fc | (rc << 3) | (is_accepted_any_case((unsigned char)c) ? 0x40u : 0u));
}

template<int H, bool COUNTING, bool EXTRA>
__global__ void __launch_bounds__(SCAN_THREADS, 3)
scan_kernel(const __grid_constant__ ScanArgs a)
{
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t* stage_buf = smem;                                               // SCAN_STAGES * SCAN_STAGE_BYTES
	uint32_t* sbits = (uint32_t*)(smem + SCAN_STAGES * SCAN_STAGE_BYTES);    // visit bits of the tile
	uint32_t* vbits = sbits + SCAN_BITWORDS;                                 // valid bits (EXTRA only)
	uint8_t* cls = (uint8_t*)(sbits + (EXTRA ? 2 : 1) * SCAN_BITWORDS);      // 256
	uint64_t* tab = (uint64_t*)(cls + 256);                                  // seed[8], rotk[8]
	uint64_t* bars = tab + 16;                                               // SCAN_STAGES mbarriers

	const int tid = threadIdx.x;
	for (int i = tid; i < 256; i += SCAN_THREADS) {
		cls[i] = class_of((unsigned)i);
	}
	for (int i = tid; i < SCAN_BITWORDS * (EXTRA ? 2 : 1); i += SCAN_THREADS) {
		sbits[i] = 0;
	}
	if (tid < 8) {
		tab[tid] = tid < 5 ? a.seed[tid] : 0;
		tab[8 + tid] = tid < 5 ? a.rotk[tid] : 0;
	}
	if (tid == 0) {
		for (int s = 0; s < SCAN_STAGES; s++) {
			mbar_init(&bars[s], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const uint32_t k = a.k;
	const int warm = (int)((k + 3) & ~3u);       // warm-up bytes: multiple of 4, >= k
	const int oshift = 8 * (int)((0u - k) & 3u); // byte phase of the outgoing character stream
	const int strip0 = SCAN_HALO + tid * SCAN_STRIP;
	const FilterView& F = a.filter;

	uint64_t tile = blockIdx.x;
	uint32_t phase[SCAN_STAGES];
	for (int s = 0; s < SCAN_STAGES; s++) {
		phase[s] = 0;
	}
	int stage = 0;
	if (tid == 0 && tile < a.n_tiles) {
		mbar_expect_tx(&bars[0], SCAN_STAGE_BYTES);
		bulk_copy_g2s(stage_buf, a.text + tile * SCAN_TILE - SCAN_HALO, SCAN_STAGE_BYTES, &bars[0]);
	}
	for (; tile < a.n_tiles; tile += gridDim.x) {
		const uint64_t next = tile + gridDim.x;
		if (tid == 0 && next < a.n_tiles) {
			// the other stage was released by the __syncthreads that closed the previous tile
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			mbar_expect_tx(&bars[stage ^ 1], SCAN_STAGE_BYTES);
			bulk_copy_g2s(stage_buf + (stage ^ 1) * SCAN_STAGE_BYTES, a.text + next * SCAN_TILE - SCAN_HALO, SCAN_STAGE_BYTES,
			              &bars[stage ^ 1]);
		}
		mbar_wait(&bars[stage], phase[stage]);
		phase[stage] ^= 1;
		const uint8_t* st = stage_buf + stage * SCAN_STAGE_BYTES;

		// ---- warm-up: build the hash of the window that ends just before the strip
		uint64_t f = 0, r = 0;
		uint32_t run = 0;
		{
			const int w0 = strip0 - warm;
			const uint32_t first = *(const uint32_t*)(st + w0);
			for (int j = 0; j < warm; j += 4) {
				const uint32_t w = *(const uint32_t*)(st + w0 + j);
#pragma unroll
				for (int b = 0; b < 4; b++) {
					const uint32_t ci = cls[(w >> (8 * b)) & 0xFF];
					uint64_t fo = 0, ro = 0;
					if ((uint32_t)(j + b) >= k) {
						const uint32_t co = cls[(first >> (8 * (j + b - (int)k))) & 0xFF];
						fo = tab[8 + (co & 7)];
						ro = tab[(co >> 3) & 7];
					}
					f = srol1(f) ^ tab[ci & 7] ^ fo;
					r = sror1(r ^ tab[8 + ((ci >> 3) & 7)] ^ ro);
					run = (ci & 0x40) ? run + 1 : 0;
				}
			}
		}

		// ---- the strip: 4 positions per step
		const int o0 = strip0 - (int)k; // byte index of the character leaving the window at the strip's first position
		uint32_t wlo = *(const uint32_t*)(st + (o0 & ~3));
		uint32_t cur = 0, curv = 0;
		int bitpos = tid * SCAN_STRIP;
		const uint64_t gbase = tile * SCAN_TILE + (uint64_t)tid * SCAN_STRIP;
		for (int wi = 0; wi < SCAN_STRIP / 4; wi++) {
			const uint32_t win = *(const uint32_t*)(st + strip0 + 4 * wi);
			const uint32_t whi = *(const uint32_t*)(st + (o0 & ~3) + 4 * wi + 4);
			const uint32_t wout = __funnelshift_r(wlo, whi, oshift);
			wlo = whi;
			uint64_t base[4];
			bool valid[4];
#pragma unroll
			for (int b = 0; b < 4; b++) {
				const uint32_t ci = cls[(win >> (8 * b)) & 0xFF];
				const uint32_t co = cls[(wout >> (8 * b)) & 0xFF];
				f = srol1(f) ^ tab[ci & 7] ^ tab[8 + (co & 7)];
				r = sror1(r ^ tab[8 + ((ci >> 3) & 7)] ^ tab[(co >> 3) & 7]);
				run = (ci & 0x40) ? run + 1 : 0;
				valid[b] = run >= k;
				base[b] = f + r;
			}
			uint32_t nib = 0, nibv = 0;
			if (a.snv && !EXTRA) {
#pragma unroll
				for (int b = 0; b < 4; b++) {
					nib |= valid[b] ? (1u << b) : 0u;
				}
			} else {
				// issue every probe of the 4 windows before looking at any of them
				uint32_t got[4][H];
				uint32_t sh[4][H];
#pragma unroll
				for (int b = 0; b < 4; b++) {
#pragma unroll
					for (int i = 0; i < H; i++) {
						uint64_t hv = base[b];
						if (i > 0) {
							hv *= a.mult[i];
							hv ^= hv >> MULTISHIFT;
						}
						const uint64_t slot = filter_slot(F, hv);
						got[b][i] = 0;
						if (COUNTING) {
							sh[b][i] = 0;
							if (valid[b]) {
								got[b][i] = ld_filter_u8(F.data + slot);
							}
						} else {
							sh[b][i] = (uint32_t)slot & 7;
							if (valid[b]) {
								got[b][i] = ld_filter_u8(F.data + (slot >> 3));
							}
						}
					}
				}
#pragma unroll
				for (int b = 0; b < 4; b++) {
					uint32_t cnt;
					if (COUNTING) {
						cnt = 255;
#pragma unroll
						for (int i = 0; i < H; i++) {
							cnt = min(cnt, got[b][i]);
						}
					} else {
						cnt = 1;
#pragma unroll
						for (int i = 0; i < H; i++) {
							cnt &= got[b][i] >> sh[b][i];
						}
					}
					if (!valid[b]) {
						cnt = 0;
					}
					const bool site = valid[b] && (a.snv ? true : (cnt == 0 || (COUNTING && cnt < a.min_threshold)));
					nib |= site ? (1u << b) : 0u;
					if (EXTRA) {
						nibv |= valid[b] ? (1u << b) : 0u;
						if (a.counts) {
							a.counts[gbase + 4 * wi + b] = (uint8_t)cnt;
						}
					}
				}
			}
			cur |= nib << (bitpos & 31);
			if (EXTRA) {
				curv |= nibv << (bitpos & 31);
			}
			bitpos += 4;
			if ((bitpos & 31) == 0) {
				if (cur) {
					atomicOr(&sbits[(bitpos - 1) >> 5], cur);
				}
				if (EXTRA && curv) {
					atomicOr(&vbits[(bitpos - 1) >> 5], curv);
				}
				cur = 0;
				curv = 0;
			}
		}
		if (cur) {
			atomicOr(&sbits[bitpos >> 5], cur);
		}
		if (EXTRA && curv) {
			atomicOr(&vbits[bitpos >> 5], curv);
		}
		__syncthreads(); // tile consumed: its stage may be refilled, its bit words are complete
		for (int w = tid; w < SCAN_BITWORDS; w += SCAN_THREADS) {
			a.visit[tile * SCAN_BITWORDS + w] = sbits[w];
			sbits[w] = 0;
			if (EXTRA) {
				if (a.valid) {
					a.valid[tile * SCAN_BITWORDS + w] = vbits[w];
				}
				vbits[w] = 0;
			}
		}
		__syncthreads();
		stage ^= 1;
	}
}

template<int H>
static cudaError_t
launch_scan_h(const ScanArgs& a, bool counting, bool extra, int grid, cudaStream_t stream)
{
	const size_t smem = scan_smem_bytes(extra);
#define NTB_LAUNCH(C, E)                                                                                              \
	do {                                                                                                              \
		cudaError_t e = cudaFuncSetAttribute(scan_kernel<H, C, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
		if (e != cudaSuccess) {                                                                                       \
			return e;                                                                                                 \
		}                                                                                                             \
		scan_kernel<H, C, E><<<grid, SCAN_THREADS, smem, stream>>>(a);                                                \
	} while (0)
	if (counting) {
		if (extra) {
			NTB_LAUNCH(true, true);
		} else {
			NTB_LAUNCH(true, false);
		}
	} else {
		if (extra) {
			NTB_LAUNCH(false, true);
		} else {
			NTB_LAUNCH(false, false);
		}
	}
#undef NTB_LAUNCH
	return cudaGetLastError();
}

cudaError_t
launch_scan(const ScanArgs& a, bool counting, bool extra, int grid, cudaStream_t stream)
{
	switch (a.filter.hash_num) {
	case 1: return launch_scan_h<1>(a, counting, extra, grid, stream);
	case 2: return launch_scan_h<2>(a, counting, extra, grid, stream);
	case 3: return launch_scan_h<3>(a, counting, extra, grid, stream);
	case 4: return launch_scan_h<4>(a, counting, extra, grid, stream);
	case 5: return launch_scan_h<5>(a, counting, extra, grid, stream);
	case 6: return launch_scan_h<6>(a, counting, extra, grid, stream);
	case 7: return launch_scan_h<7>(a, counting, extra, grid, stream);
	case 8: return launch_scan_h<8>(a, counting, extra, grid, stream);
	default: return cudaErrorInvalidValue;
	}
}

// ------------------------------------------------------------------------------------------------------------------
// K1b: binned scan (see kernels.cuh)
__device__ __forceinline__ uint64_t
policy_evict_first()
{
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}

__device__ __forceinline__ uint64_t
policy_evict_last()
{
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}

// a probe record leaves the SM once and is read once by the probe kernel much later: keep it from displacing the filter region
// the probe kernel holds in L2 (NTB_BIN_STORE_HINT: 0 plain store, 1 st.global.cs, 2 L2 evict-first policy)
#ifndef NTB_BIN_STORE_HINT
#define NTB_BIN_STORE_HINT 0
#endif
__device__ __forceinline__ void
st_record(uint64_t* p, uint64_t v, uint64_t policy)
{
#if NTB_BIN_STORE_HINT == 1
	asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#elif NTB_BIN_STORE_HINT == 2
	asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(policy) : "memory");
#else
	asm volatile("st.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); // (the row pointers come out of shared memory: say "global")
#endif
}

// the probe a record stands for, done on the spot: is the k-mer's bit clear / its counter below the threshold?
template<bool COUNTING>
__device__ __forceinline__ bool
probe_misses(const uint8_t* data, uint64_t slot, uint32_t thr)
{
	if (COUNTING) {
		return ld_filter_u8(data + slot) < thr;
	}
	return ((ld_filter_u8(data + (slot >> 3)) >> ((uint32_t)slot & 7u)) & 1u) == 0;
}

template<int H, bool COUNTING, bool POW2>
__global__ void __launch_bounds__(SCAN_THREADS, BIN_CTAS_PER_SM)
bin_kernel(const __grid_constant__ BinArgs A)
{
	const ScanArgs& a = A.scan;
	constexpr int RR = SCAN_THREADS * BIN_POS_PER_ROUND * H; // records a round can produce
	constexpr uint32_t NONE = BIN_MAX_BUCKETS;               // "bucket" of a window that is not valid: 32 spare counters, one per lane
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t* stage_buf = smem;                                                // BIN_STAGES * SCAN_STAGE_BYTES
	// the round's records, grouped by bucket: slot in region : 32 | bucket : 16 | position in tile : 16 (SCAN_TILE < 2^16)
	uint64_t* sorted = (uint64_t*)(smem + BIN_STAGES * SCAN_STAGE_BYTES);
	uint64_t** rowp = (uint64_t**)(sorted + RR);                              // per bucket: where sorted[j] goes, minus j
	uint32_t* cnt = (uint32_t*)(rowp + BIN_MAX_BUCKETS);                      // per bucket: records of this round (+ 32 spare)
	uint32_t* off = cnt + BIN_MAX_BUCKETS + 32;                               // per bucket: start inside sorted[] (+ 32 spare)
	uint32_t* wsum = off + BIN_MAX_BUCKETS + 32;                              // [0..8) warp totals, [8] round total
	uint8_t* cls = (uint8_t*)(wsum + 16);                                     // 256
	uint64_t* tab = (uint64_t*)(cls + 256);                                   // seed[8], rotk[8]
	uint64_t* bars = tab + 16;                                                // BIN_STAGES mbarriers

	const int tid = threadIdx.x;
	const int lane = tid & 31, warp = tid >> 5;
	for (int i = tid; i < 256; i += SCAN_THREADS) {
		cls[i] = class_of((unsigned)i);
	}
	for (int i = tid; i < BIN_MAX_BUCKETS + 32; i += SCAN_THREADS) {
		cnt[i] = 0;
		off[i] = 0;
	}
	if (tid < 8) {
		tab[tid] = tid < 5 ? a.seed[tid] : 0;
		tab[8 + tid] = tid < 5 ? a.rotk[tid] : 0;
	}
	if (tid == 0) {
		for (int s = 0; s < BIN_STAGES; s++) {
			mbar_init(&bars[s], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const uint32_t k = a.k;
	const int warm = (int)((k + 3) & ~3u);
	const int oshift = 8 * (int)((0u - k) & 3u);
	const int strip0 = SCAN_HALO + tid * SCAN_STRIP;
	const FilterView& F = a.filter;
	const uint32_t rl = A.region_log2;
	const uint64_t rmask = (1ULL << rl) - 1ULL;
	const uint32_t nb = A.n_buckets;
	const uint32_t cap = A.bucket_cap;
	const uint32_t thr = a.min_threshold > 1u ? a.min_threshold : 1u;
	const uint64_t pol_stream = NTB_BIN_STORE_HINT == 2 ? policy_evict_first() : 0;
	const uint32_t none = NONE + (uint32_t)lane;

	uint64_t tile = blockIdx.x;
	uint32_t phase[BIN_STAGES];
	for (int s = 0; s < BIN_STAGES; s++) {
		phase[s] = 0;
	}
	int stage = 0;
	if (tid == 0 && tile < a.n_tiles) {
		mbar_expect_tx(&bars[0], SCAN_STAGE_BYTES);
		bulk_copy_g2s(stage_buf, a.text + tile * SCAN_TILE - SCAN_HALO, SCAN_STAGE_BYTES, &bars[0]);
	}
	for (; tile < a.n_tiles; tile += gridDim.x) {
		const uint64_t next = tile + gridDim.x;
		if (BIN_STAGES == 2 && tid == 0 && next < a.n_tiles) {
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			mbar_expect_tx(&bars[stage ^ 1], SCAN_STAGE_BYTES);
			bulk_copy_g2s(stage_buf + (stage ^ 1) * SCAN_STAGE_BYTES, a.text + next * SCAN_TILE - SCAN_HALO, SCAN_STAGE_BYTES,
			              &bars[stage ^ 1]);
		}
		mbar_wait(&bars[stage], phase[stage]);
		phase[stage] ^= 1;
		const uint8_t* st = stage_buf + stage * SCAN_STAGE_BYTES;

		// ---- warm-up: hash of the window that ends just before the strip (as in scan_kernel)
		uint64_t f = 0, r = 0;
		uint32_t run = 0;
		{
			const int w0 = strip0 - warm;
			const uint32_t first = *(const uint32_t*)(st + w0);
			for (int j = 0; j < warm; j += 4) {
				const uint32_t w = *(const uint32_t*)(st + w0 + j);
#pragma unroll
				for (int b = 0; b < 4; b++) {
					const uint32_t ci = cls[(w >> (8 * b)) & 0xFF];
					uint64_t fo = 0, ro = 0;
					if ((uint32_t)(j + b) >= k) {
						const uint32_t co = cls[(first >> (8 * (j + b - (int)k))) & 0xFF];
						fo = tab[8 + (co & 7)];
						ro = tab[(co >> 3) & 7];
					}
					f = srol1(f) ^ tab[ci & 7] ^ fo;
					r = sror1(r ^ tab[8 + ((ci >> 3) & 7)] ^ ro);
					run = (ci & 0x40) ? run + 1 : 0;
				}
			}
		}

		const int o0 = strip0 - (int)k;
		uint32_t wlo = *(const uint32_t*)(st + (o0 & ~3));
		const uint32_t tile_pos = (uint32_t)(tile * SCAN_TILE); // position of the tile inside the chunk
		const uint32_t pit0 = (uint32_t)tid * SCAN_STRIP;       // position of the strip inside the tile
		for (int wi = 0; wi < SCAN_STRIP / 4; wi++) {
			const uint32_t win = *(const uint32_t*)(st + strip0 + 4 * wi);
			const uint32_t whi = *(const uint32_t*)(st + (o0 & ~3) + 4 * wi + 4);
			const uint32_t wout = __funnelshift_r(wlo, whi, oshift);
			wlo = whi;
			// ---- (a) the round's records: slot inside its region, bucket, rank inside the bucket (shared atomics).  A window that
			// is not valid counts on the lane's spare counter and is dropped in (c): no branch around the atomics
			uint32_t rsir[4 * H];
			uint32_t rbr[4 * H]; // bucket << 16 | rank ; bucket >= NONE = no record
#pragma unroll
			for (int b = 0; b < 4; b++) {
				const uint32_t ci = cls[(win >> (8 * b)) & 0xFF];
				const uint32_t co = cls[(wout >> (8 * b)) & 0xFF];
				f = srol1(f) ^ tab[ci & 7] ^ tab[8 + (co & 7)];
				r = sror1(r ^ tab[8 + ((ci >> 3) & 7)] ^ tab[(co >> 3) & 7]);
				run = (ci & 0x40) ? run + 1 : 0;
				const bool valid = run >= k;
				const uint64_t base = f + r;
#pragma unroll
				for (int i = 0; i < H; i++) {
					uint64_t hv = base;
					if (i > 0) {
						hv *= a.mult[i];
						hv ^= hv >> MULTISHIFT;
					}
					const uint64_t slot = POW2 ? (hv & F.mask) : filter_slot(F, hv);
					const uint32_t bucket = valid ? (uint32_t)(slot >> rl) : none;
					rsir[b * H + i] = (uint32_t)(slot & rmask);
					rbr[b * H + i] = (bucket << 16) | (atomicAdd(&cnt[bucket], 1u) & 0xFFFFu);
				}
			}
			__syncthreads();
			// ---- (b) one thread per bucket: reserve the global rows, scan the counts
			uint32_t c = 0;
			if ((uint32_t)tid < nb) {
				c = cnt[tid];
				cnt[tid] = 0;
			}
			uint32_t incl = c;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
				if (lane >= o) {
					incl += v;
				}
			}
			if (lane == 31) {
				wsum[warp] = incl;
			}
			uint32_t g = 0;
			if (c) {
				g = atomicAdd(&A.cursor[tid], c);
			}
			// a row that runs full in this round (heavily repeated k-mers): the round takes the checked copy-out
			const int over = __syncthreads_or((uint64_t)g + c > (uint64_t)cap);
			uint32_t woff = 0;
#pragma unroll
			for (int w = 0; w < SCAN_THREADS / 32; w++) {
				woff += w < warp ? wsum[w] : 0u;
			}
			if ((uint32_t)tid < nb) {
				const uint32_t o = woff + incl - c;
				off[tid] = o;
				// (with "over" the pointer is not used: it carries the row index of sorted[o] instead)
				rowp[tid] = over ? (uint64_t*)(uintptr_t)(uint64_t)g : A.records + ((uint64_t)tid * cap + g) - o;
			}
			if (tid == SCAN_THREADS - 1) {
				wsum[8] = woff + incl;
			}
			__syncthreads();
			// ---- (c) group the records by bucket
#pragma unroll
			for (int q = 0; q < 4 * H; q++) {
				const uint32_t bucket = rbr[q] >> 16;
				const uint32_t j = off[bucket] + (rbr[q] & 0xFFFFu);
				if (bucket < NONE) {
					sorted[j] = ((uint64_t)rsir[q] << 32) | (uint64_t)((rbr[q] & 0xFFFF0000u) | (pit0 + 4 * wi + q / H));
				}
			}
			__syncthreads();
			// ---- (d) contiguous runs out to the buckets' rows
			const uint32_t total = wsum[8];
			if (!over) {
#pragma unroll 4
				for (uint32_t j = tid; j < total; j += SCAN_THREADS) {
					const uint2 sr = *(const uint2*)&sorted[j];
					st_record(rowp[sr.x >> 16] + j, ((uint64_t)sr.y << 32) | (tile_pos + (sr.x & 0xFFFFu)), pol_stream);
				}
			} else {
				for (uint32_t j = tid; j < total; j += SCAN_THREADS) {
					const uint2 sr = *(const uint2*)&sorted[j];
					const uint32_t bucket = sr.x >> 16;
					const uint32_t pos = tile_pos + (sr.x & 0xFFFFu);
					const uint64_t idx = (uint64_t)(uintptr_t)rowp[bucket] + (j - off[bucket]);
					if (idx < cap) {
						st_record(&A.records[(uint64_t)bucket * cap + idx], ((uint64_t)sr.y << 32) | pos, pol_stream);
					} else {
						// the bucket's rows are full: probe directly
						const uint64_t slot = ((uint64_t)bucket << rl) | sr.y;
						if (probe_misses<COUNTING>(F.data, slot, thr)) {
							atomicOr(&a.visit[pos >> 5], 1u << (pos & 31));
						}
					}
				}
			}
			// the next round's shared atomics may start: cnt[] was cleared in (b); sorted[] / off[] / rowp[] are only written
			// again after the next round's first barrier, which every thread reaches after finishing (d)
		}
		__syncthreads(); // tile consumed: its stage may be refilled
		if (BIN_STAGES == 2) {
			stage ^= 1;
		} else if (tid == 0 && next < a.n_tiles) {
			// one stage: the next tile is fetched now; the other CTAs of the SM cover the wait
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			mbar_expect_tx(&bars[0], SCAN_STAGE_BYTES);
			bulk_copy_g2s(stage_buf, a.text + next * SCAN_TILE - SCAN_HALO, SCAN_STAGE_BYTES, &bars[0]);
		}
	}
}

__device__ __forceinline__ ulonglong2
ld_records_stream(const uint64_t* p, uint64_t policy)
{
	ulonglong2 v;
	asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(policy));
	return v;
}

__device__ __forceinline__ uint32_t
ld_filter_keep(const uint8_t* p, uint64_t policy)
{
	uint32_t v;
	asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
	return v;
}

template<bool COUNTING>
__global__ void __launch_bounds__(256)
probe_bin_kernel(const __grid_constant__ BinArgs A)
{
	constexpr int U = 4; // 16-byte record loads in flight per thread
	const uint64_t pol_stream = policy_evict_first();
	const uint64_t pol_keep = policy_evict_last();
	const uint32_t rl = A.region_log2;
	const uint32_t cap = A.bucket_cap;
	const uint32_t thr = A.scan.min_threshold > 1u ? A.scan.min_threshold : 1u;
	uint32_t* visit = A.scan.visit;
	const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t gstride = (uint64_t)gridDim.x * blockDim.x;
	// pacing: the probes only hit L2 while every CTA works on (nearly) the same region, so a CTA starts bucket b only when
	// every CTA has finished bucket b - pace_lag: one region hot at a time with pace_lag = 1 (the default: regions of half
	// the L2), two with pace_lag = 2.  A.cursor[BIN_MAX_BUCKETS] counts the (CTA, bucket) pairs that are done; the grid is
	// launched cooperatively, so every CTA is resident and the wait ends.
	unsigned int* done = A.cursor + BIN_MAX_BUCKETS;
	for (uint32_t b = 0; b < A.n_buckets; b++) {
		if (b >= A.pace_lag) {
			if (threadIdx.x == 0) {
				const unsigned int target = (b + 1 - A.pace_lag) * gridDim.x;
				unsigned int seen;
				do {
					asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done) : "memory");
				} while (seen < target);
			}
			__syncthreads();
		}
		const uint32_t have = A.cursor[b];
		const uint32_t n = have < cap ? have : cap;
		const uint64_t* recs = A.records + (uint64_t)b * cap;
		const uint8_t* region = A.scan.filter.data + (COUNTING ? ((uint64_t)b << rl) : ((uint64_t)b << (rl - 3)));
		const uint64_t npair = n >> 1;
		for (uint64_t p0 = gtid; p0 < npair; p0 += gstride * U) {
			ulonglong2 rec[U];
#pragma unroll
			for (int u = 0; u < U; u++) {
				const uint64_t p = p0 + (uint64_t)u * gstride;
				rec[u] = make_ulonglong2(~0ULL, ~0ULL);
				if (p < npair) {
					rec[u] = ld_records_stream(recs + 2 * p, pol_stream);
				}
			}
			uint32_t got[2 * U];
#pragma unroll
			for (int u = 0; u < U; u++) {
				const uint64_t r0 = rec[u].x, r1 = rec[u].y;
				const uint32_t s0 = (uint32_t)(r0 >> 32), s1 = (uint32_t)(r1 >> 32);
				got[2 * u] = 0xFFu;
				got[2 * u + 1] = 0xFFu;
				if (r0 != ~0ULL) {
					got[2 * u] = ld_filter_keep(region + (COUNTING ? s0 : (s0 >> 3)), pol_keep);
				}
				if (r1 != ~0ULL) {
					got[2 * u + 1] = ld_filter_keep(region + (COUNTING ? s1 : (s1 >> 3)), pol_keep);
				}
			}
#pragma unroll
			for (int u = 0; u < U; u++) {
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const uint64_t rr = h ? rec[u].y : rec[u].x;
					const uint32_t v = got[2 * u + h];
					const bool miss = COUNTING ? (v < thr) : (((v >> ((uint32_t)(rr >> 32) & 7u)) & 1u) == 0);
					if (rr != ~0ULL && miss) {
						const uint32_t pos = (uint32_t)rr;
						atomicOr(&visit[pos >> 5], 1u << (pos & 31));
					}
				}
			}
		}
		if ((n & 1u) && gtid == 0) {
			const uint64_t rr = recs[n - 1];
			const uint64_t slot = ((uint64_t)b << rl) | (rr >> 32);
			if (probe_misses<COUNTING>(A.scan.filter.data, slot, thr)) {
				const uint32_t pos = (uint32_t)rr;
				atomicOr(&visit[pos >> 5], 1u << (pos & 31));
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			atomicAdd(done, 1u);
		}
	}
}

template<int H, bool COUNTING, bool POW2>
static cudaError_t
launch_bin_hcp(const BinArgs& a, int grid, cudaStream_t stream)
{
	size_t smem = bin_smem_bytes(H);
	if (const char* v = std::getenv("NTB_BIN_SMEM_PAD_KB")) {
		smem += (size_t)std::strtoul(v, nullptr, 10) << 10; // experiment: fewer bin CTAs per SM, room for another kernel beside them
	}
	const cudaError_t e = cudaFuncSetAttribute(bin_kernel<H, COUNTING, POW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) {
		return e;
	}
	bin_kernel<H, COUNTING, POW2><<<grid, SCAN_THREADS, smem, stream>>>(a);
	return cudaGetLastError();
}

template<int H>
static cudaError_t
launch_bin_h(const BinArgs& a, bool counting, int grid, cudaStream_t stream)
{
	const bool pow2 = a.scan.filter.mask != 0;
	if (counting) {
		return pow2 ? launch_bin_hcp<H, true, true>(a, grid, stream) : launch_bin_hcp<H, true, false>(a, grid, stream);
	}
	return pow2 ? launch_bin_hcp<H, false, true>(a, grid, stream) : launch_bin_hcp<H, false, false>(a, grid, stream);
}

cudaError_t
launch_bin(const BinArgs& a, bool counting, int grid, cudaStream_t stream)
{
	if (a.n_buckets == 0 || a.n_buckets > (uint32_t)BIN_MAX_BUCKETS || a.region_log2 < 3 || a.region_log2 > 32) {
		return cudaErrorInvalidValue;
	}
	switch (a.scan.filter.hash_num) {
	case 1: return launch_bin_h<1>(a, counting, grid, stream);
	case 2: return launch_bin_h<2>(a, counting, grid, stream);
	case 3: return launch_bin_h<3>(a, counting, grid, stream);
	case 4: return launch_bin_h<4>(a, counting, grid, stream);
	case 5: return launch_bin_h<5>(a, counting, grid, stream);
	case 6: return launch_bin_h<6>(a, counting, grid, stream);
	case 7: return launch_bin_h<7>(a, counting, grid, stream);
	case 8: return launch_bin_h<8>(a, counting, grid, stream);
	default: return cudaErrorInvalidValue;
	}
}

cudaError_t
launch_probe_bin(const BinArgs& a, bool counting, int grid_probe, cudaStream_t stream)
{
	cudaError_t e;
	// cooperative launch: all CTAs resident (the kernel paces itself across CTAs); grid_probe = CTAs per SM wanted
	static OccCache cache[2];
	const void* fn = counting ? (const void*)probe_bin_kernel<true> : (const void*)probe_bin_kernel<false>;
	int per_sm = 0;
	e = walker_occupancy(fn, cache[counting ? 1 : 0], 0, 256, &per_sm);
	if (e != cudaSuccess) {
		return e;
	}
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int want = grid_probe > 0 && grid_probe < per_sm ? grid_probe : per_sm;
	BinArgs args = a;
	void* params[1] = { (void*)&args };
	return cudaLaunchCooperativeKernel(fn, dim3((unsigned)(sms * want)), dim3(256), params, 0, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// K2
// K2: persistent warps, one task (contig segment) per warp at a time, tasks handed out through an atomic counter.
// The walker state of every warp lives in shared memory (engine.h: WalkerState); lane 0 is the leader.
// NCAP (capacity of the local rope copy) is 2.5 k + 32 rounded up: 160 serves k <= 48, 352 serves k <= KMAX.
// the CTA's copy of the parameters, the rotation table and the class table in front of the team states (engine.h finds them
// by this layout); ends with a CTA barrier
__device__ __forceinline__ void
walk_smem_init(uint8_t* walk_smem, const KParams& kp)
{
	for (uint32_t q = threadIdx.x; q < 256; q += blockDim.x) {
		walk_smem[WALK_KP_BYTES + WALK_ROT_BYTES + q] = class_of(q);
	}
	uint64_t* rot = reinterpret_cast<uint64_t*>(walk_smem + WALK_KP_BYTES);
	for (uint32_t q = threadIdx.x; q < sizeof(KParams) / 4; q += blockDim.x) {
		reinterpret_cast<uint32_t*>(walk_smem)[q] = reinterpret_cast<const uint32_t*>(&kp)[q];
	}
	for (uint32_t q = threadIdx.x; q < ROT_WORDS; q += blockDim.x) {
		rot[q] = rot_entry(q);
	}
	__syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// Pre-evaluation (ntb_common.h: SiteRec; engine.h: pre_run).
// heads_kernel: one warp per task lists the heads of its nominal range [start, end) -- the ranges partition every contig.
// A head is a flagged position with no flagged position among the `gap` (< 128) positions in front of it: every lane takes
// one bitmap word, gets the three words in front of it from its neighbours and smears that 128-bit window instead of
// testing bit by bit.
struct Bits128
{
	uint64_t lo, hi;
};

__device__ __forceinline__ Bits128
shl128(Bits128 x, uint32_t n) // n < 64
{
	Bits128 r;
	r.hi = n ? (x.hi << n) | (x.lo >> (64 - n)) : x.hi;
	r.lo = x.lo << n;
	return r;
}

__global__ void __launch_bounds__(256)
heads_kernel(const uint32_t* visit, const Task* tasks, uint32_t n_tasks, uint32_t gap, uint2* items, uint32_t cap, Counters* ctr)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t ti = warp; ti < n_tasks; ti += n_warps) {
		const Task t = tasks[ti];
		const uint64_t g0 = t.text_off + t.start, g1 = t.text_off + t.end;
		const uint64_t gc = t.text_off; // flagged positions in front of the contig's first base do not count
		for (uint64_t w0 = g0 >> 5; (w0 << 5) < g1; w0 += 32) {
			const uint64_t w = w0 + lane;
			// this lane's word and the three in front of it (bits in front of the contig masked away)
			auto word = [&](uint64_t wi) -> uint32_t {
				if ((int64_t)wi < (int64_t)(gc >> 5) || (wi << 5) >= g1 + 32) {
					return 0u;
				}
				uint32_t b = visit[wi];
				if (wi == (gc >> 5)) {
					b &= 0xFFFFFFFFu << (gc & 31);
				}
				return b;
			};
			const uint32_t cur = (w << 5) < g1 ? word(w) : 0u;
			uint32_t p1 = __shfl_up_sync(0xFFFFFFFFu, cur, 1), p2 = __shfl_up_sync(0xFFFFFFFFu, cur, 2), p3 = __shfl_up_sync(0xFFFFFFFFu, cur, 3);
			if (lane < 1) {
				p1 = word(w - 1);
			}
			if (lane < 2) {
				p2 = word(w - 2);
			}
			if (lane < 3) {
				p3 = word(w - 3);
			}
			Bits128 x;
			x.lo = ((uint64_t)p2 << 32) | p3;
			x.hi = ((uint64_t)cur << 32) | p1;
			// near = positions with a flagged position 1 .. gap in front of them: x smeared upwards by gap, less x itself
			Bits128 s = shl128(x, 1);
			for (uint32_t covered = 1; covered < gap;) { // s covers distances 1 .. covered; gap <= KMAX - 1 keeps every step below 64
				const uint32_t step = covered < gap - covered ? covered : gap - covered;
				const Bits128 sh = shl128(s, step);
				s.lo |= sh.lo;
				s.hi |= sh.hi;
				covered += step;
			}
			uint32_t heads = cur & ~(uint32_t)(s.hi >> 32);
			// only positions of [g0, g1) are this task's
			if (w == (g0 >> 5)) {
				heads &= 0xFFFFFFFFu << (g0 & 31);
			}
			if (((w + 1) << 5) > g1) {
				heads &= (w << 5) < g1 ? (0xFFFFFFFFu >> (32 - (g1 & 31))) : 0u;
			}
			// warp-aggregated append
			const uint32_t n = (uint32_t)__popc(heads);
			uint32_t incl = n;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
				if (lane >= (uint32_t)o) {
					incl += v;
				}
			}
			const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
			if (total == 0) {
				continue;
			}
			uint32_t base = 0;
			if (lane == 0) {
				base = atomicAdd(&ctr->n_items, total);
			}
			base = __shfl_sync(0xFFFFFFFFu, base, 0) + incl - n;
			for (uint32_t rest = heads; rest; rest &= rest - 1, base++) {
				if (base < cap) {
					items[base] = make_uint2(ti, (uint32_t)((w << 5) + ((uint32_t)__ffs((int)rest) - 1u) - t.text_off));
				}
			}
		}
	}
}

// presite_kernel<.., SECOND = false>: first pass, one item (head + chain) per warp at a time, tryIndels left out of the code;
// <.., SECOND = true>: second pass over the sites the first one left pending (and the chains behind them).  Same shared-memory layout and persistent-warp
// scheme as walk_kernel.
template<int NCAP, bool COMMON, bool POW2, bool SECOND>
__global__ void __launch_bounds__(WALK_THREADS, NTB_WALK_MIN_CTAS)
presite_kernel(const uint8_t* text, const uint32_t* visit, FilterView bloom, FilterView rep, const __grid_constant__ KParams kp,
               const Task* tasks, const uint2* items, uint32_t items_cap, SiteRec* table, uint32_t table_mask, PendingSite* pending,
               uint32_t pending_cap, Counters* ctr)
{
	extern __shared__ __align__(16) uint8_t walk_smem[];
	WalkerState<NCAP>* states = reinterpret_cast<WalkerState<NCAP>*>(walk_smem + WALK_KP_BYTES + WALK_ROT_BYTES + WALK_CLS_BYTES);
	walk_smem_init(walk_smem, kp);
	WalkerState<NCAP>& S = states[threadIdx.x / NTB_TEAM];
	const uint64_t* rot = reinterpret_cast<const uint64_t*>(walk_smem + WALK_KP_BYTES);
	const uint32_t lane = lane_id();
	Walker<NCAP, COMMON, POW2> w(S, kp);
	const uint32_t n_units = SECOND ? min(ctr->n_pending, pending_cap) : min(ctr->n_items, items_cap);
	for (;;) {
		uint32_t i = 0;
		if (lane == 0) {
			i = atomicAdd(SECOND ? &ctr->next_pending : &ctr->next_item, 1u);
		}
		i = __shfl_sync(team_mask(), i, (int)team_base());
		if (i >= n_units) {
			break;
		}
		uint32_t ti, pos;
		if (SECOND) {
			const PendingSite ps = pending[i];
			ti = ps.task;
			pos = ps.pos;
		} else {
			const uint2 it = items[i];
			ti = it.x;
			pos = it.y;
		}
		const Task task = tasks[ti];
		warp_sync();
		if (lane == 0) {
			S.io.text = text + task.text_off;
			S.io.len = task.len;
			S.io.visit = visit;
			S.io.goff = task.text_off;
			S.io.bloom = bloom;
			S.io.rep = rep;
			S.io.events = nullptr;
			S.io.ev_cap = 0;
			S.io.ctr = ctr;
			S.io.rot = rot;
			S.io.table = table;
			S.io.table_mask = table_mask;
			S.io.pending = pending;
			S.io.pending_cap = pending_cap;
		}
		warp_sync();
		w.pre_begin();
		w.pre_run(ti, pos, SECOND);
	}
}

// presite_dense_kernel: the first pass with one THREAD per site (site_dense.h) -- 32 sites per warp in flight, every probe
// of a stage issued before any is consumed.  It runs in rounds.  Round 0 takes the heads, one per thread.  A site that made
// no edit hands its chain to the next round (a thread that followed its chain itself would hold its whole warp for the
// ~3 % of heads that fail): there DENSE_GROUP consecutive lanes take one chain and evaluate its next DENSE_GROUP sites
// side by side -- sites are functions of the text alone -- then keep the records up to the first one at which the main
// loop would stop going on (an edit, or tryIndels pending), exactly the records the strictly sequential chain files; only
// when all of them let it go on does the chain move to the round after.  SITE_CHAIN_MAX / DENSE_GROUP rounds.
constexpr int DENSE_THREADS = 128;
constexpr uint32_t DENSE_ROUNDS = 1 + SITE_CHAIN_MAX / DENSE_GROUP;

__global__ void
presite_round_kernel(Counters* ctr, uint32_t items_cap)
{
	ctr->item_lo = ctr->item_hi;
	ctr->item_hi = min(ctr->n_items, items_cap);
}

template<int KCAP, bool CHAIN>
__global__ void __launch_bounds__(DENSE_THREADS)
presite_dense_kernel(const uint8_t* text, const uint32_t* visit, FilterView bloom, FilterView rep, const __grid_constant__ KParams kp,
                     const Task* tasks, uint2* items, uint32_t items_cap, SiteRec* table, uint32_t table_mask, PendingSite* pending,
                     uint32_t pending_cap, Counters* ctr)
{
	const uint32_t lo = ctr->item_lo, hi = ctr->item_hi;
	if (lo >= hi) {
		return;
	}
	__shared__ uint64_t rot[ROT_WORDS];
	__shared__ uint8_t cls[256];
	for (uint32_t q = threadIdx.x; q < ROT_WORDS; q += blockDim.x) {
		rot[q] = rot_entry(q);
	}
	for (uint32_t q = threadIdx.x; q < 256; q += blockDim.x) {
		cls[q] = class_of(q);
	}
	__syncthreads();
	DenseCtx C;
	C.kp = &kp;
	C.bloom = bloom;
	C.rep = rep;
	C.rot = rot;
	C.cls = cls;
	const uint32_t gap = kp.k - 1;
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
	if (!CHAIN) {
		for (uint32_t i = lo + gtid; i < hi; i += gthreads) {
			const uint2 it = items[i];
			const Task task = tasks[it.x];
			SiteRec r;
			const uint32_t st = dense_site<KCAP>(C, text + task.text_off, task.len, it.y, r);
			if (dense_commit(r, st, task.text_off, it.x, it.y, table, table_mask, pending, pending_cap, ctr) && dense_continues(st, r) &&
			    dense_chain_next(visit, task.text_off, task.len, it.y, gap) != NONE32) {
				const uint32_t j = atomicAdd(&ctr->n_items, 1u);
				if (j < items_cap) {
					items[j] = it; // the chain behind this site
				}
			}
		}
		return;
	}
	// chain rounds: item = (task, position the chain goes on BEHIND); the whole warp iterates together
	const uint32_t lane = threadIdx.x & 31u, sub = lane % DENSE_GROUP, gbase = lane - sub;
	const uint32_t n_groups = gthreads / DENSE_GROUP;
	const uint32_t n_iter = (hi - lo + n_groups - 1) / n_groups;
	for (uint32_t iter = 0; iter < n_iter; iter++) {
		const uint32_t i = lo + iter * n_groups + gtid / DENSE_GROUP;
		bool active = i < hi;
		uint2 it = make_uint2(0, 0);
		Task task;
		uint32_t pos = NONE32;
		if (active) {
			it = items[i];
			task = tasks[it.x];
			// this lane's site: the (sub + 1)-th of the chain behind it.y
			pos = it.y;
			for (uint32_t q = 0; q <= sub && pos != NONE32; q++) {
				pos = dense_chain_next(visit, task.text_off, task.len, pos, gap);
			}
			active = pos != NONE32;
		}
		SiteRec r;
		uint32_t st = SITE_NONE;
		if (active) {
			st = dense_site<KCAP>(C, text + task.text_off, task.len, pos, r);
		}
		// the first lane of the group at which the chain stops: its end, an edit, or a pending site
		const bool stops = !active || !dense_continues(st, r);
		const uint32_t stop_mask = (__ballot_sync(0xFFFFFFFFu, stops) >> gbase) & ((1u << DENSE_GROUP) - 1u);
		const uint32_t first_stop = stop_mask ? (uint32_t)__ffs((int)stop_mask) - 1u : (uint32_t)DENSE_GROUP;
		// a no-edit record learns how many of the chain's next sites (the lanes behind it, up to the stop) make no edit and
		// emit nothing either: the walker jumps over them (SITE_FL_SKIP)
		if (!kp.mask && !kp.snv) {
			const uint64_t info = !active ? 0ull : (uint64_t)st | ((uint64_t)r.best_type << 8) | ((uint64_t)r.flags << 16) | ((uint64_t)dense_pack_bases(r) << 32);
			uint32_t T = SKIP_IDENTITY;
			uint32_t n_skip = 0, dist = 0;
			bool run = active && sub < first_stop && dense_skippable(st, r.best_type, r.flags);
#pragma unroll
			for (uint32_t d = 1; d < (uint32_t)DENSE_GROUP; d++) {
				const int src = (int)(gbase + ((sub + d) % DENSE_GROUP));
				const uint64_t oi = __shfl_sync(0xFFFFFFFFu, info, src);
				const uint32_t op = __shfl_sync(0xFFFFFFFFu, pos, src);
				run = run && sub + d < first_stop && dense_skippable((uint32_t)(oi & 0xFF), (uint32_t)((oi >> 8) & 0xFF), (uint32_t)((oi >> 16) & 0xFF));
				if (run) {
					T = dense_skip_compose(T, (uint32_t)(oi & 0xFF), (uint32_t)(oi >> 32));
					n_skip++;
					dist = op - pos;
				}
			}
			if (n_skip) {
				dense_skip_store(r, n_skip, dist, T);
			}
		}
		bool ok = true;
		if (active && sub <= first_stop) {
			ok = dense_commit(r, st, task.text_off, it.x, pos, table, table_mask, pending, pending_cap, ctr);
		}
		if (first_stop == (uint32_t)DENSE_GROUP && sub == DENSE_GROUP - 1 && ok) {
			const uint32_t j = atomicAdd(&ctr->n_items, 1u);
			if (j < items_cap) {
				items[j] = make_uint2(it.x, pos);
			}
		}
	}
}

// K3 snv_dense_kernel (-s 1, ntedit.cpp:1806 "opt::snv ||": every valid window is a site): one thread per position
// evaluates the site from the text (site_dense.h: dense_site -- draft baseline from the check subset, gates, sampled
// trial windows, best / alternates), files a record for the positions whose commit does something (an accepted
// substitution, a variant record, a case change) and marks them in a second bitmap.  The walkers then jump through THAT
// bitmap: between two such positions the main loop's iterations have no observable effect.  One warp per task, lanes on
// consecutive positions (their text and their filter probes are independent; neighbouring windows share cache lines of text).
template<int KCAP>
__global__ void __launch_bounds__(DENSE_THREADS)
snv_dense_kernel(const uint8_t* text, const uint32_t* visit, uint32_t* visit2, FilterView bloom, FilterView rep, const __grid_constant__ KParams kp,
                 const Task* tasks, uint32_t n_tasks, SiteRec* table, uint32_t table_mask, Counters* ctr)
{
	__shared__ uint64_t rot[ROT_WORDS];
	__shared__ uint8_t cls[256];
	for (uint32_t q = threadIdx.x; q < ROT_WORDS; q += blockDim.x) {
		rot[q] = rot_entry(q);
	}
	for (uint32_t q = threadIdx.x; q < 256; q += blockDim.x) {
		cls[q] = class_of(q);
	}
	__syncthreads();
	DenseCtx C;
	C.kp = &kp;
	C.bloom = bloom;
	C.rep = rep;
	C.rot = rot;
	C.cls = cls;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t ti = warp; ti < n_tasks; ti += n_warps) {
		const Task t = tasks[ti];
		for (uint32_t p0 = t.start; p0 < t.end; p0 += 32) {
			const uint32_t pos = p0 + lane;
			const uint64_t g = t.text_off + pos;
			if (pos >= t.end || !((visit[g >> 5] >> (g & 31)) & 1u)) {
				continue;
			}
			SiteRec r;
			const uint32_t st = dense_site<KCAP>(C, text + t.text_off, t.len, pos, r);
			if (dense_has_effect(st, r, kp.mask != 0)) {
				atomicOr(&visit2[g >> 5], 1u << (g & 31));
				dense_commit(r, st, t.text_off, ti, pos, table, table_mask, nullptr, 0, ctr);
			}
		}
	}
}

template<int NCAP, bool COMMON, bool POW2>
__global__ void __launch_bounds__(WALK_THREADS, NTB_WALK_MIN_CTAS)
walk_kernel(const uint8_t* text, const uint32_t* visit, FilterView bloom, FilterView rep, const __grid_constant__ KParams kp,
            const Task* tasks, const uint32_t* order, TaskResult* results, uint32_t n_tasks, Event* events, uint32_t ev_cap, Counters* ctr,
            SiteRec* table, uint32_t table_mask)
{
	extern __shared__ __align__(16) uint8_t walk_smem[];
	// [KParams copy][rotation table][class table][team states] -- engine.h's Walker finds all of them by this layout
	WalkerState<NCAP>* states = reinterpret_cast<WalkerState<NCAP>*>(walk_smem + WALK_KP_BYTES + WALK_ROT_BYTES + WALK_CLS_BYTES);
	walk_smem_init(walk_smem, kp);
	WalkerState<NCAP>& S = states[threadIdx.x / NTB_TEAM];
	const uint64_t* rot = reinterpret_cast<const uint64_t*>(walk_smem + WALK_KP_BYTES);
	const uint32_t lane = lane_id();
	Walker<NCAP, COMMON, POW2> w(S, kp);
	bool have = false;
	uint32_t i = 0;
	Task task;
	long long c0 = 0;
	// every team runs its own task; the teams of a warp re-converge at the top of each iteration
	for (;;) {
		if (!have) {
			if (lane == 0) {
				i = atomicAdd(&ctr->next_task, 1u);
			}
			i = __shfl_sync(team_mask(), i, (int)team_base());
			if (i < n_tasks) {
				if (order) {
					i = order[i];
				}
				task = tasks[i];
				warp_sync();
				if (lane == 0) {
					S.io.text = text + task.text_off;
					S.io.len = task.len;
					S.io.visit = visit;
					S.io.goff = task.text_off;
					S.io.bloom = bloom;
					S.io.rep = rep;
					S.io.events = events;
					S.io.ev_cap = ev_cap;
					S.io.ctr = ctr;
					S.io.rot = rot;
					S.io.table = table;
					S.io.table_mask = table_mask;
					S.io.pending = nullptr;
					S.io.pending_cap = 0;
				}
				warp_sync();
				c0 = clock64();
#if defined(NTB_PHASE_PROF)
				if (lane == 0) {
					for (int q = 0; q < 16; q++) {
						S.prof[q] = 0;
					}
				}
#endif
				w.begin(task);
				have = true;
			}
		}
		if (!have) {
			break;
		}
		if (!w.step(task)) {
			TaskResult res;
			w.finish(res);
			if (lane == 0) {
				res.kcycles = (uint32_t)((clock64() - c0) >> 10);
				results[i] = res;
#if defined(NTB_PHASE_PROF)
				for (int q = 0; q < 16; q++) {
					atomicAdd(&ctr->prof[q], (unsigned long long)S.prof[q]);
				}
#endif
			}
			have = false;
		}
	}
}

// Puts the tasks with many flagged positions at the front of the work queue (they take the longest: every flagged
// position in an unfixable stretch can cost a full insertion / deletion enumeration), the rest behind them.
__global__ void __launch_bounds__(256)
order_tasks_kernel(const uint32_t* visit, const Task* tasks, uint32_t n_tasks, uint32_t dense_threshold, uint32_t* order, Counters* ctr)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_tasks) {
		return;
	}
	const Task t = tasks[i];
	const uint64_t g0 = t.text_off + t.start, g1 = t.text_off + t.end;
	uint32_t cnt = 0;
	for (uint64_t w = g0 >> 5; (w << 5) < g1 && cnt < dense_threshold; w++) {
		uint32_t bits = visit[w];
		if (w == (g0 >> 5)) {
			bits &= 0xFFFFFFFFu << (g0 & 31);
		}
		if (((w + 1) << 5) > g1) {
			bits &= 0xFFFFFFFFu >> (32 - (g1 & 31));
		}
		cnt += __popc(bits);
	}
	if (cnt >= dense_threshold) {
		order[atomicAdd(&ctr->n_front, 1u)] = i;
	} else {
		order[n_tasks - 1 - atomicAdd(&ctr->n_back, 1u)] = i;
	}
}

// The walkers append events to one arena through an atomic cursor, so the events of a walker are scattered and only linked
// backwards.  One thread per task copies its chain into a contiguous run, first event first, and points the result at it:
// the host then reads every walker's events sequentially instead of chasing `prev` through a 100 MB arena.
__global__ void __launch_bounds__(256)
compact_events_kernel(const Event* in, Event* out, TaskResult* results, uint32_t n_tasks, Counters* ctr)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_tasks) {
		return;
	}
	const uint32_t n = results[i].n_events;
	if (n == 0) {
		results[i].last_event = NONE32;
		return;
	}
	const uint32_t base = atomicAdd(&ctr->n_compact, n);
	uint32_t e = results[i].last_event;
	for (uint32_t j = n; j > 0 && e != NONE32; j--) {
		const Event ev = in[e];
		out[base + j - 1] = ev;
		e = ev.prev;
	}
	results[i].last_event = base;
}

// tasks of a round, fetched from the pinned host buffer by the SMs themselves: a cudaMemcpyAsync would queue on the
// host-to-device copy engine behind every piece of a text that is still being uploaded
__global__ void __launch_bounds__(256)
fetch_host_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src_host, size_t n16)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
		dst[i] = src_host[i];
	}
}

cudaError_t
launch_fetch_host(void* dst, const void* src_host, size_t bytes, cudaStream_t stream)
{
	const size_t n16 = (bytes + 15) / 16; // (both buffers are allocated in whole 16-byte units)
	if (n16 == 0) {
		return cudaSuccess;
	}
	const unsigned grid = (unsigned)std::min<size_t>((n16 + 255) / 256, 1184);
	fetch_host_kernel<<<grid, 256, 0, stream>>>((uint4*)dst, (const uint4*)src_host, n16);
	return cudaGetLastError();
}

cudaError_t
launch_compact_events(const Event* in, Event* out, TaskResult* results, uint32_t n_tasks, Counters* ctr, cudaStream_t stream)
{
	if (n_tasks == 0) {
		return cudaSuccess;
	}
	compact_events_kernel<<<(n_tasks + 255) / 256, 256, 0, stream>>>(in, out, results, n_tasks, ctr);
	return cudaGetLastError();
}

template<int NCAP>
static size_t
walk_smem_bytes()
{
	return WALK_KP_BYTES + WALK_ROT_BYTES + WALK_CLS_BYTES + (size_t)WALK_TEAMS * sizeof(WalkerState<NCAP>);
}

template<int NCAP, bool COMMON, bool POW2>
static cudaError_t
launch_walk_n(const WalkArgs& a, cudaStream_t stream)
{
	static OccCache cache;
	const size_t smem = walk_smem_bytes<NCAP>();
	int blocks_per_sm = 0;
	cudaError_t e = walker_occupancy(walk_kernel<NCAP, COMMON, POW2>, cache, smem, WALK_THREADS, &blocks_per_sm);
	if (e != cudaSuccess) {
		return e;
	}
	if (const char* cap = std::getenv("NTB_WALK_BLOCKS_PER_SM")) { // tuning aid
		const int c = std::atoi(cap);
		if (c > 0 && c < blocks_per_sm) {
			blocks_per_sm = c;
		}
	}
	const uint64_t want = ((uint64_t)a.n_tasks + WALK_TEAMS - 1) / WALK_TEAMS;
	const uint64_t cap = (uint64_t)a.sm_count * (uint64_t)blocks_per_sm;
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	if (grid == 0) {
		return cudaSuccess;
	}
	walk_kernel<NCAP, COMMON, POW2><<<grid, WALK_THREADS, smem, stream>>>(a.text, a.visit, a.bloom, a.rep, a.kp, a.tasks, a.order, a.results,
	                                                                      a.n_tasks, a.events, a.ev_cap, a.ctr, a.table, a.table_mask);
	return cudaGetLastError();
}

template<int NCAP, bool COMMON, bool POW2, bool SECOND>
static cudaError_t
launch_presite_n(const WalkArgs& a, cudaStream_t stream)
{
	static OccCache cache;
	const size_t smem = walk_smem_bytes<NCAP>();
	int blocks_per_sm = 0;
	cudaError_t e = walker_occupancy(presite_kernel<NCAP, COMMON, POW2, SECOND>, cache, smem, WALK_THREADS, &blocks_per_sm);
	if (e != cudaSuccess) {
		return e;
	}
	if (const char* cap = std::getenv("NTB_WALK_BLOCKS_PER_SM")) { // tuning aid
		const int c = std::atoi(cap);
		if (c > 0 && c < blocks_per_sm) {
			blocks_per_sm = c;
		}
	}
	// persistent grid: the number of items is only known on the device
	const unsigned grid = (unsigned)(a.sm_count * blocks_per_sm);
	presite_kernel<NCAP, COMMON, POW2, SECOND><<<grid, WALK_THREADS, smem, stream>>>(a.text, a.visit, a.bloom, a.rep, a.kp, a.tasks, a.items,
	                                                                                a.items_cap, a.table, a.table_mask, a.pending,
	                                                                                a.pending_cap, a.ctr);
	return cudaGetLastError();
}

// the specialised instantiations serve the common configuration (see engine.h: Walker<NCAP, COMMON, POW2>)
#define NTB_WALK_DISPATCH(FN, ...)                                                                   \
	do {                                                                                             \
		const bool common = !a.kp.counting && !a.kp.h_rep && !a.kp.snv && !a.kp.mask;                \
		const bool pow2 = common && a.bloom.mask != 0;                                               \
		if (a.kp.k <= 48) {                                                                          \
			return pow2     ? FN<160, true, true __VA_ARGS__>(a, stream)                             \
			       : common ? FN<160, true, false __VA_ARGS__>(a, stream)                            \
			                : FN<160, false, false __VA_ARGS__>(a, stream);                          \
		}                                                                                            \
		return pow2     ? FN<352, true, true __VA_ARGS__>(a, stream)                                 \
		       : common ? FN<352, true, false __VA_ARGS__>(a, stream)                                \
		                : FN<352, false, false __VA_ARGS__>(a, stream);                              \
	} while (0)

cudaError_t
launch_walk(const WalkArgs& a, cudaStream_t stream)
{
	if (a.order && a.n_tasks) {
		order_tasks_kernel<<<(a.n_tasks + 255) / 256, 256, 0, stream>>>(a.visit, a.tasks, a.n_tasks, 24u, a.order, a.ctr);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) {
			return e;
		}
	}
	NTB_WALK_DISPATCH(launch_walk_n);
}

cudaError_t
launch_heads(const WalkArgs& a, cudaStream_t stream)
{
	if (a.n_tasks == 0) {
		return cudaSuccess;
	}
	const unsigned warps_per_cta = 8;
	const unsigned want = (a.n_tasks + warps_per_cta - 1) / warps_per_cta;
	const unsigned cap = (unsigned)a.sm_count * 8u;
	heads_kernel<<<want < cap ? want : cap, warps_per_cta * 32, 0, stream>>>(a.visit, a.tasks, a.n_tasks, a.kp.k - 1, a.items, a.items_cap, a.ctr);
	return cudaGetLastError();
}

static cudaError_t
launch_presite_first(const WalkArgs& a, cudaStream_t stream)
{
#define NTB_COMMA_FALSE , false
	NTB_WALK_DISPATCH(launch_presite_n, NTB_COMMA_FALSE);
#undef NTB_COMMA_FALSE
}

static cudaError_t
launch_presite_second(const WalkArgs& a, cudaStream_t stream)
{
#define NTB_COMMA_TRUE , true
	NTB_WALK_DISPATCH(launch_presite_n, NTB_COMMA_TRUE);
#undef NTB_COMMA_TRUE
}

template<int KCAP>
static cudaError_t
launch_presite_dense_k(const WalkArgs& a, cudaStream_t stream)
{
	static OccCache cache[2];
	int per_sm = 0, per_sm_chain = 0;
	cudaError_t e = walker_occupancy(presite_dense_kernel<KCAP, false>, cache[0], 0, DENSE_THREADS, &per_sm);
	if (e == cudaSuccess) {
		e = walker_occupancy(presite_dense_kernel<KCAP, true>, cache[1], 0, DENSE_THREADS, &per_sm_chain);
	}
	if (e != cudaSuccess) {
		return e;
	}
	for (uint32_t round = 0; round < DENSE_ROUNDS; round++) {
		presite_round_kernel<<<1, 1, 0, stream>>>(a.ctr, a.items_cap);
		// (later rounds hold a few percent of the heads, then nothing: an empty round returns at once)
		if (round == 0) {
			presite_dense_kernel<KCAP, false><<<(unsigned)(a.sm_count * per_sm), DENSE_THREADS, 0, stream>>>(
			    a.text, a.visit, a.bloom, a.rep, a.kp, a.tasks, a.items, a.items_cap, a.table, a.table_mask, a.pending, a.pending_cap, a.ctr);
		} else {
			presite_dense_kernel<KCAP, true><<<(unsigned)(a.sm_count * per_sm_chain), DENSE_THREADS, 0, stream>>>(
			    a.text, a.visit, a.bloom, a.rep, a.kp, a.tasks, a.items, a.items_cap, a.table, a.table_mask, a.pending, a.pending_cap, a.ctr);
		}
	}
	return cudaGetLastError();
}

cudaError_t
launch_presite(const WalkArgs& a, bool second, cudaStream_t stream)
{
	if (second) {
		return launch_presite_second(a, stream);
	}
	// NTB_PRESITE_DENSE=0 (testing aid): the warp-per-site form of the first pass
	const char* dv = std::getenv("NTB_PRESITE_DENSE");
	if (dv && dv[0] == '0') {
		return launch_presite_first(a, stream);
	}
	return a.kp.k <= 48 ? launch_presite_dense_k<48>(a, stream) : launch_presite_dense_k<(int)KMAX>(a, stream);
}

template<int KCAP>
static cudaError_t
launch_snv_dense_k(const WalkArgs& a, uint32_t* visit2, cudaStream_t stream)
{
	static OccCache cache;
	int per_sm = 0;
	cudaError_t e = walker_occupancy(snv_dense_kernel<KCAP>, cache, 0, DENSE_THREADS, &per_sm);
	if (e != cudaSuccess) {
		return e;
	}
	const uint64_t warps_per_cta = DENSE_THREADS / 32;
	const uint64_t want = ((uint64_t)a.n_tasks + warps_per_cta - 1) / warps_per_cta, cap = (uint64_t)a.sm_count * (uint64_t)per_sm;
	snv_dense_kernel<KCAP><<<(unsigned)(want < cap ? want : cap), DENSE_THREADS, 0, stream>>>(a.text, a.visit, visit2, a.bloom, a.rep, a.kp, a.tasks,
	                                                                                      a.n_tasks, a.table, a.table_mask, a.ctr);
	return cudaGetLastError();
}

cudaError_t
launch_snv_dense(const WalkArgs& a, uint32_t* visit2, cudaStream_t stream)
{
	if (a.n_tasks == 0) {
		return cudaSuccess;
	}
	return a.kp.k <= 48 ? launch_snv_dense_k<48>(a, visit2, stream) : launch_snv_dense_k<(int)KMAX>(a, visit2, stream);
}

uint32_t
presite_launch_count()
{
	return 2 + 2 * DENSE_ROUNDS; // heads, the rounds of the first pass (bookkeeping + sites), second pass
}
#undef NTB_WALK_DISPATCH

// ------------------------------------------------------------------------------------------------------------------
// K5: filter construction.  One thread per strip of INSERT_STRIP positions.
constexpr int INSERT_STRIP = 256;

__device__ __forceinline__ void
bump_counter(uint8_t* data, uint64_t slot)
{
	// saturating 8-bit increment through a CAS on the enclosing 32-bit word
	unsigned int* word = (unsigned int*)(data + (slot & ~3ULL));
	const unsigned sh = 8 * (unsigned)(slot & 3);
	unsigned int old = *word;
	for (;;) {
		const unsigned int c = (old >> sh) & 0xFF;
		if (c == 255) {
			return;
		}
		const unsigned int want = (old & ~(0xFFu << sh)) | ((c + 1) << sh);
		const unsigned int seen = atomicCAS(word, old, want);
		if (seen == old) {
			return;
		}
		old = seen;
	}
}

__global__ void __launch_bounds__(128)
insert_kernel(const uint8_t* text, uint64_t total, uint8_t* data, FilterView f, const __grid_constant__ KParams kp)
{
	const uint64_t g0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * INSERT_STRIP;
	if (g0 >= total) {
		return;
	}
	const uint32_t k = kp.k;
	const uint64_t gend = g0 + INSERT_STRIP < total ? g0 + INSERT_STRIP : total;
	// warm-up over the k-1 bytes in front of the strip (the buffer has SCAN_HALO zero bytes in front of position 0)
	HashState hs;
	hs.fh = hs.rh = 0;
	uint32_t run = 0;
	const int64_t start = (int64_t)g0 - (int64_t)(k - 1);
	for (int64_t p = start; p < (int64_t)gend; p++) {
		const unsigned char cin = text[p];
		const unsigned char cout = (p - (int64_t)k >= start) ? text[p - (int64_t)k] : (unsigned char)0;
		hash_roll(hs, cout, cin, kp);
		run = (base_code(cin) < 4 && (cin | 0x20) != 'u') ? run + 1 : 0; // all-ACGT windows only, as btllib's insert(seq)
		if (p >= (int64_t)g0 && run >= k) {
			const uint64_t b = hash_canonical(hs);
			for (uint32_t i = 0; i < f.hash_num; i++) {
				const uint64_t slot = filter_slot(f, hash_extend(b, k, i));
				if (f.counting) {
					bump_counter(data, slot);
				} else {
					atomicOr((unsigned int*)(data + ((slot >> 5) << 2)), 1u << (slot & 31));
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// K4: set bits (bit filter) or non-zero counters (counting filter)
__global__ void __launch_bounds__(256)
occupancy_kernel(const uint8_t* data, uint64_t bytes, int counting, unsigned long long* out)
{
	unsigned long long local = 0;
	const uint64_t nvec = bytes / 16;
	const uint4* v = (const uint4*)data;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint4 q = v[i];
		const uint32_t w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (counting) {
				// number of non-zero bytes in the word
				uint32_t x = w[j];
				x |= x >> 4;
				x |= x >> 2;
				x |= x >> 1;
				local += __popc(x & 0x01010101u);
			} else {
				local += __popc(w[j]);
			}
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		for (uint64_t i = nvec * 16; i < bytes; i++) {
			local += counting ? (data[i] != 0) : __popc((unsigned)data[i]);
		}
	}
	for (int o = 16; o > 0; o >>= 1) {
		local += __shfl_down_sync(0xFFFFFFFFu, local, o);
	}
	if ((threadIdx.x & 31) == 0 && local) {
		atomicAdd(out, local);
	}
}

} // namespace ntb
