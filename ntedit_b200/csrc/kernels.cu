// sm_100a kernels of the ntEdit hot path -- see kernels.cuh for the map to the reference.
#include "kernels.cuh"

#include <cstdlib>

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
// K2
// K2: persistent warps, one task (contig segment) per warp at a time, tasks handed out through an atomic counter.
// The walker state of every warp lives in shared memory (engine.h: WalkerState); lane 0 is the leader.
// NCAP (capacity of the local rope copy) is 2.5 k + 32 rounded up: 160 serves k <= 48, 352 serves k <= KMAX.
template<int NCAP>
__global__ void __launch_bounds__(WALK_THREADS)
walk_kernel(const uint8_t* text, const uint32_t* visit, FilterView bloom, FilterView rep, const __grid_constant__ KParams kp,
            const Task* tasks, const uint32_t* order, TaskResult* results, uint32_t n_tasks, Event* events, uint32_t ev_cap, Counters* ctr)
{
	extern __shared__ __align__(16) uint8_t walk_smem[];
	WalkerState<NCAP>* states = reinterpret_cast<WalkerState<NCAP>*>(walk_smem);
	WalkerState<NCAP>& S = states[threadIdx.x / NTB_TEAM];
	uint64_t* rot = reinterpret_cast<uint64_t*>(walk_smem + (size_t)WALK_TEAMS * sizeof(WalkerState<NCAP>));
	for (uint32_t q = threadIdx.x; q < ROT_WORDS; q += WALK_THREADS) {
		rot[q] = rot_entry(q);
	}
	__syncthreads();
	const uint32_t lane = lane_id();
	Walker<NCAP> w(S, kp);
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

template<int NCAP>
static cudaError_t
launch_walk_n(const uint8_t* text, const uint32_t* visit, const FilterView& bloom, const FilterView& rep, const KParams& kp, const Task* tasks,
              const uint32_t* order, TaskResult* results, uint32_t n_tasks, Event* events, uint32_t ev_cap, Counters* ctr, int sm_count,
              cudaStream_t stream)
{
	static int blocks_per_sm = 0;
	const size_t smem = (size_t)WALK_TEAMS * sizeof(WalkerState<NCAP>) + ROT_WORDS * sizeof(uint64_t);
	if (blocks_per_sm == 0) {
		cudaError_t e = cudaFuncSetAttribute(walk_kernel<NCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) {
			return e;
		}
		int n = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, walk_kernel<NCAP>, WALK_THREADS, smem);
		if (e != cudaSuccess) {
			return e;
		}
		blocks_per_sm = n > 0 ? n : 1;
		if (const char* cap = std::getenv("NTB_WALK_BLOCKS_PER_SM")) { // tuning aid
			const int c = std::atoi(cap);
			if (c > 0 && c < blocks_per_sm) {
				blocks_per_sm = c;
			}
		}
	}
	const uint64_t want = ((uint64_t)n_tasks + WALK_TEAMS - 1) / WALK_TEAMS;
	const uint64_t cap = (uint64_t)sm_count * (uint64_t)blocks_per_sm;
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	if (grid == 0) {
		return cudaSuccess;
	}
	walk_kernel<NCAP><<<grid, WALK_THREADS, smem, stream>>>(text, visit, bloom, rep, kp, tasks, order, results, n_tasks, events, ev_cap, ctr);
	return cudaGetLastError();
}

cudaError_t
launch_walk(const uint8_t* text, const uint32_t* visit, const FilterView& bloom, const FilterView& rep, const KParams& kp, const Task* tasks,
            uint32_t* order, TaskResult* results, uint32_t n_tasks, Event* events, uint32_t ev_cap, Counters* ctr, int sm_count,
            cudaStream_t stream)
{
	if (order && n_tasks) {
		order_tasks_kernel<<<(n_tasks + 255) / 256, 256, 0, stream>>>(visit, tasks, n_tasks, 24u, order, ctr);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) {
			return e;
		}
	}
	if (kp.k <= 48) {
		return launch_walk_n<160>(text, visit, bloom, rep, kp, tasks, order, results, n_tasks, events, ev_cap, ctr, sm_count, stream);
	}
	return launch_walk_n<352>(text, visit, bloom, rep, kp, tasks, order, results, n_tasks, events, ev_cap, ctr, sm_count, stream);
}

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
