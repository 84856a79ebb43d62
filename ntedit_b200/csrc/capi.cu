// C ABI of libntedit_b200.so (include/ntedit_b200.h): filter objects, batches, the CUDA backend of the polishing driver.
#include "../../include/ntedit_b200.h"
#include "filter_io.hpp"
#include "kernels.cuh"
#include "polish_driver.hpp"
#include "writer.hpp"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <new>

using namespace ntb;

namespace {

thread_local std::string g_error;

int
fail(int code, const std::string& msg)
{
	g_error = msg;
	return code;
}

int
cuda_fail(cudaError_t e, const char* what)
{
	const int code = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? NTB_ENODEV
	                 : (e == cudaErrorMemoryAllocation)                          ? NTB_ENOMEM
	                                                                             : NTB_ECUDA;
	return fail(code, std::string(what) + ": " + cudaGetErrorString(e));
}

#define NTB_CUDA(call)                           \
	do {                                         \
		cudaError_t e_ = (call);                 \
		if (e_ != cudaSuccess) {                 \
			return cuda_fail(e_, #call);         \
		}                                        \
	} while (0)

int
select_device(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		return fail(NTB_ENODEV, "no CUDA device available (ntedit_b200 has no CPU fallback)");
	}
	if (device < 0 || device >= n) {
		return fail(NTB_EINVAL, "device index out of range");
	}
	NTB_CUDA(cudaSetDevice(device));
	return NTB_OK;
}

int
sm_count(int device)
{
	int n = 148;
	cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
	return n > 0 ? n : 148;
}

} // namespace

struct ntb_filter
{
	uint8_t* d = nullptr; // device bytes (padded to 16)
	uint64_t bytes = 0;
	uint32_t k = 0, h = 0;
	int counting = 0;
	int device = 0;
	bool owned = true;
	double fpr = -1.0; // cached; reset by inserts

	FilterView view() const
	{
		FilterView v;
		std::memset(&v, 0, sizeof v);
		v.data = d;
		v.bytes = bytes;
		v.mod = counting ? bytes : bytes * 8;
		v.recip = v.mod ? 0xFFFFFFFFFFFFFFFFULL / v.mod : 0;
		v.mask = (v.mod > 1 && (v.mod & (v.mod - 1)) == 0) ? v.mod - 1 : 0;
		v.hash_num = h;
		v.counting = counting ? 1u : 0u;
		return v;
	}
};

struct ntb_batch
{
	uint8_t* d_alloc = nullptr; // SCAN_HALO zero bytes, the text, zero padding to whole scan tiles
	bool owns_text = true;      // false: d_alloc is the calling workspace's cached text buffer (ntb_polish_batch)
	uint8_t* d_text = nullptr;  // d_alloc + SCAN_HALO
	uint64_t total = 0;         // bytes of text (including the NUL separators)
	uint64_t n_tiles = 0;
	std::vector<uint64_t> offsets;
	int device = 0;
	float ms_h2d = 0;
	// streamed upload (ntb_polish_batch): the text is copied piece by piece on its own stream while the scan of the earlier
	// pieces already runs; piece i is on the device when up_events[i] has fired
	const char* up_src = nullptr;
	cudaStream_t up_stream = nullptr;
	cudaEvent_t up_begin = nullptr, up_end = nullptr;
	std::vector<cudaEvent_t> up_events;
	uint64_t up_issued = 0; // bytes whose copy has been issued
	static constexpr uint64_t UP_PIECE = 64ull << 20;

	// issue the copies of every piece that overlaps [0, end); returns the event that covers byte end-1 (nullptr: nothing streamed)
	cudaError_t upload_until(uint64_t end, cudaEvent_t* covering)
	{
		*covering = nullptr;
		if (!up_src) {
			return cudaSuccess;
		}
		if (end > total) {
			end = total;
		}
		while (up_issued < end) {
			const uint64_t len = std::min<uint64_t>(UP_PIECE, total - up_issued);
			cudaError_t e = cudaMemcpyAsync(d_text + up_issued, up_src + up_issued, len, cudaMemcpyHostToDevice, up_stream);
			if (e != cudaSuccess) {
				return e;
			}
			cudaEvent_t ev = nullptr;
			e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
			if (e != cudaSuccess) {
				return e;
			}
			up_events.push_back(ev);
			e = cudaEventRecord(ev, up_stream);
			if (e != cudaSuccess) {
				return e;
			}
			up_issued += len;
			if (up_issued == total && up_end) {
				cudaEventRecord(up_end, up_stream);
			}
		}
		if (end > 0 && !up_events.empty()) {
			*covering = up_events[std::min<uint64_t>(up_events.size() - 1, (end - 1) / UP_PIECE)];
		}
		return cudaSuccess;
	}
};

struct ntb_result
{
	ResultImpl impl;
};

namespace {

// bytes of device memory a batch of `total` text bytes needs (halo in front, zero padding to whole scan tiles behind)
uint64_t
batch_device_bytes(uint64_t total)
{
	return SCAN_HALO + (total + SCAN_TILE - 1) / SCAN_TILE * SCAN_TILE + 64;
}

// lays the batch out in `mem` (batch_device_bytes(total) bytes) and zeroes halo and padding on `stream`
int
batch_bind(ntb_batch* b, uint64_t total, uint8_t* mem, bool owned, cudaStream_t stream)
{
	b->total = total;
	b->n_tiles = (total + SCAN_TILE - 1) / SCAN_TILE;
	const uint64_t padded = b->n_tiles * SCAN_TILE;
	b->d_alloc = mem;
	b->owns_text = owned;
	b->d_text = b->d_alloc + SCAN_HALO;
	NTB_CUDA(cudaMemsetAsync(b->d_alloc, 0, SCAN_HALO, stream));
	NTB_CUDA(cudaMemsetAsync(b->d_text + total, 0, padded - total + 64, stream));
	return NTB_OK;
}

int
batch_alloc(ntb_batch* b, uint64_t total, cudaStream_t stream = 0)
{
	uint8_t* mem = nullptr;
	NTB_CUDA(cudaMalloc((void**)&mem, batch_device_bytes(total)));
	return batch_bind(b, total, mem, true, stream);
}

int
check_offsets(const uint64_t* offsets, uint64_t n_contigs)
{
	if (!offsets) {
		return fail(NTB_EINVAL, "offsets is NULL");
	}
	if (offsets[0] != 0) {
		return fail(NTB_EINVAL, "offsets[0] must be 0");
	}
	for (uint64_t c = 0; c < n_contigs; c++) {
		if (offsets[c + 1] <= offsets[c]) {
			return fail(NTB_EINVAL, "offsets must be strictly increasing (every contig carries its NUL terminator)");
		}
	}
	return NTB_OK;
}

void
fill_scan_tables(ScanArgs& a, uint32_t k)
{
	const uint64_t seeds[5] = { SEED_A, SEED_C, SEED_G, SEED_T, 0 };
	for (int i = 0; i < 5; i++) {
		a.seed[i] = seeds[i];
		a.rotk[i] = sroln(seeds[i], k);
	}
	for (unsigned i = 0; i < HMAX; i++) {
		a.mult[i] = (uint64_t)i ^ ((uint64_t)k * MULTISEED);
	}
}

// Device + pinned-host scratch of one polishing call.  Grow-only and cached per device in a small pool so that
// back-to-back calls (the reference's OpenMP loop makes one per contig batch) do not pay cudaMalloc / cudaFree /
// cudaHostAlloc again; concurrent calls from different host threads each check out their own workspace.
struct Workspace
{
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t stream2 = nullptr;                      // K1b: the probe kernels run beside the next chunk's bin kernel
	cudaStream_t stream_d2h = nullptr;                   // a round's events and results go to the host beside the next group's scan
	cudaEvent_t ev_compact = nullptr;                    // ... once they are grouped by walker
	cudaEvent_t ev_bin[2] = { nullptr, nullptr };        // K1b: records of buffer i are complete
	cudaEvent_t ev_probe[2] = { nullptr, nullptr };      // K1b: buffer i has been consumed
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	std::vector<cudaEvent_t> ev_scan;                    // K1 / K1b: begin / end of every scanned range of a call
	uint32_t* d_visit = nullptr;
	size_t cap_visit = 0; // words
	uint32_t* d_visit2 = nullptr; // -s 1: the positions whose site does something (K3)
	size_t cap_visit2 = 0;
	Task* d_tasks = nullptr;
	uint32_t* d_order = nullptr;
	TaskResult* d_results = nullptr;
	size_t cap_tasks = 0;
	Event* d_events = nullptr;
	Event* d_events_sorted = nullptr; // the same events, grouped by walker (compact_events_kernel)
	size_t cap_events = 0;
	Counters* d_ctr = nullptr;
	uint8_t* d_text = nullptr;     // text buffer of streamed batches (cudaMalloc / cudaFree of GBs per call can stall for 100s of ms)
	size_t cap_text = 0;
	uint64_t* d_records = nullptr; // K1b: probe records, n_buckets x bucket_cap
	size_t cap_records = 0;
	uint32_t* d_cursor = nullptr;  // K1b: BIN_MAX_BUCKETS record counters + the probe kernel's pacing counter
	SiteRec* d_table = nullptr;    // pre-evaluated sites (open addressing, power-of-two slots)
	size_t cap_table = 0;
	uint2* d_items = nullptr;      // heads of flagged runs (task, position)
	size_t cap_items = 0;
	PendingSite* d_pending = nullptr;
	size_t cap_pending = 0;
	cudaEvent_t ev_pre0 = nullptr, ev_pre1 = nullptr, ev_pre_mid = nullptr;
	Counters* h_ctr_pre = nullptr; // pinned: the counters after the pre-evaluation passes (diagnostics)
	// pinned host mirrors
	Task* h_tasks = nullptr;
	TaskResult* h_results = nullptr;
	size_t cap_h_tasks = 0;
	// pinned event arenas, one per walker round of a call: a round's events never move, so the host may replay one round
	// while the device fills the next (polish_driver.hpp: contig groups)
	std::vector<Event*> h_event_blocks;
	std::vector<size_t> cap_event_blocks;
	Counters* h_ctr = nullptr;

	~Workspace()
	{
		cudaSetDevice(device);
		cudaFree(d_visit);
		cudaFree(d_visit2);
		cudaFree(d_tasks);
		cudaFree(d_order);
		cudaFree(d_results);
		cudaFree(d_events);
		cudaFree(d_events_sorted);
		cudaFree(d_ctr);
		cudaFree(d_records);
		cudaFree(d_cursor);
		cudaFree(d_text);
		cudaFree(d_table);
		cudaFree(d_items);
		cudaFree(d_pending);
		cudaFreeHost(h_ctr_pre);
		if (ev_pre0) {
			cudaEventDestroy(ev_pre0);
		}
		if (ev_pre1) {
			cudaEventDestroy(ev_pre1);
		}
		if (ev_pre_mid) {
			cudaEventDestroy(ev_pre_mid);
		}
		cudaFreeHost(h_tasks);
		cudaFreeHost(h_results);
		for (Event* b : h_event_blocks) {
			cudaFreeHost(b);
		}
		cudaFreeHost(h_ctr);
		if (ev0) {
			cudaEventDestroy(ev0);
		}
		if (ev1) {
			cudaEventDestroy(ev1);
		}
		for (cudaEvent_t ev : ev_scan) {
			cudaEventDestroy(ev);
		}
		for (int i = 0; i < 2; i++) {
			if (ev_bin[i]) {
				cudaEventDestroy(ev_bin[i]);
			}
			if (ev_probe[i]) {
				cudaEventDestroy(ev_probe[i]);
			}
		}
		if (ev_compact) {
			cudaEventDestroy(ev_compact);
		}
		if (stream_d2h) {
			cudaStreamDestroy(stream_d2h);
		}
		if (stream2) {
			cudaStreamDestroy(stream2);
		}
		if (stream) {
			cudaStreamDestroy(stream);
		}
	}
};

std::mutex g_pool_mutex;
std::vector<Workspace*> g_pool;

Workspace*
workspace_acquire(int device)
{
	{
		std::lock_guard<std::mutex> lock(g_pool_mutex);
		for (size_t i = 0; i < g_pool.size(); i++) {
			if (g_pool[i]->device == device) {
				Workspace* w = g_pool[i];
				g_pool.erase(g_pool.begin() + (long)i);
				return w;
			}
		}
	}
	Workspace* w = new (std::nothrow) Workspace();
	if (w) {
		w->device = device;
	}
	return w;
}

void
workspace_release(Workspace* w)
{
	if (!w) {
		return;
	}
	std::lock_guard<std::mutex> lock(g_pool_mutex);
	if (g_pool.size() >= 8) {
		delete w;
	} else {
		g_pool.push_back(w);
	}
}

// CUDA implementation of the Backend concept of polish_driver.hpp
struct CudaBackend
{
	ntb_filter* bloom;
	ntb_filter* rep;
	ntb_batch* batch;
	Workspace* ws = nullptr;
	size_t n_rounds = 0; // walker rounds of this call so far; round r's events live in ws->h_event_blocks[r]
	float ms_scan = 0, ms_walk = 0, ms_d2h = 0, ms_pre = 0;
	bool pre_timed = false;       // ev_pre0 / ev_pre1 bracket this call's pre-evaluation passes
	size_t table_slots = 0;       // slots of ws->d_table that hold this call's records (0: no pre-evaluation)
	bool use_visit2 = false;      // -s 1 with K3: the walkers jump through ws->d_visit2
	uint32_t launches = 0;
	std::string err;
	int rc = NTB_OK;

	const std::string& error() const { return err; }

	const Event* round_events(size_t r) const { return ws->h_event_blocks[r]; }

	// the pinned arena of the round about to run, with room for n events
	int round_block(size_t n, Event** out)
	{
		if (n_rounds >= ws->h_event_blocks.size()) {
			ws->h_event_blocks.push_back(nullptr);
			ws->cap_event_blocks.push_back(0);
		}
		if (n > ws->cap_event_blocks[n_rounds]) {
			cudaFreeHost(ws->h_event_blocks[n_rounds]);
			ws->h_event_blocks[n_rounds] = nullptr;
			ws->cap_event_blocks[n_rounds] = 0;
			const size_t want = n + n / 4 + 4096;
			cudaError_t e = cudaHostAlloc((void**)&ws->h_event_blocks[n_rounds], want * sizeof(Event), cudaHostAllocDefault);
			if (e != cudaSuccess) {
				return cuda_err(e, "cudaHostAlloc(events)");
			}
			ws->cap_event_blocks[n_rounds] = want;
		}
		*out = ws->h_event_blocks[n_rounds];
		return NTB_OK;
	}

	int cuda_err(cudaError_t e, const char* what)
	{
		err = std::string(what) + ": " + cudaGetErrorString(e);
		rc = e == cudaErrorMemoryAllocation ? NTB_ENOMEM : NTB_ECUDA;
		return rc;
	}

#define NTB_BE(call)                          \
	do {                                      \
		cudaError_t e_ = (call);              \
		if (e_ != cudaSuccess) {              \
			return cuda_err(e_, #call);       \
		}                                     \
	} while (0)

	int init()
	{
		ws = workspace_acquire(batch->device);
		if (!ws) {
			err = "out of memory";
			return rc = NTB_ENOMEM;
		}
		if (!ws->stream) {
			NTB_BE(cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking));
			NTB_BE(cudaEventCreate(&ws->ev0));
			NTB_BE(cudaEventCreate(&ws->ev1));
			NTB_BE(cudaMalloc((void**)&ws->d_ctr, sizeof(Counters)));
			NTB_BE(cudaHostAlloc((void**)&ws->h_ctr, sizeof(Counters), cudaHostAllocDefault));
		}
		if (batch->up_src && !batch->d_alloc) {
			// a streamed batch lives in the workspace's cached text buffer
			const uint64_t need = batch_device_bytes(batch->total);
			if (need > ws->cap_text) {
				cudaFree(ws->d_text);
				ws->d_text = nullptr;
				ws->cap_text = 0;
				NTB_BE(cudaMalloc((void**)&ws->d_text, need + need / 16));
				ws->cap_text = need + need / 16;
			}
			if (batch_bind(batch, batch->total, ws->d_text, false, batch->up_stream) != NTB_OK) {
				err = g_error;
				return rc = NTB_ECUDA;
			}
			NTB_BE(cudaEventRecord(batch->up_begin, batch->up_stream));
			if (batch->total == 0) {
				NTB_BE(cudaEventRecord(batch->up_end, batch->up_stream));
			}
		}
		const size_t words = batch->n_tiles * SCAN_BITWORDS + 16;
		if (words > ws->cap_visit) {
			cudaFree(ws->d_visit);
			ws->d_visit = nullptr;
			ws->cap_visit = 0;
			NTB_BE(cudaMalloc((void**)&ws->d_visit, words * 4));
			ws->cap_visit = words;
		}
		// a batch that is not streamed was uploaded on the default stream
		if (!batch->up_src) {
			NTB_BE(cudaStreamSynchronize(0));
		}
		return NTB_OK;
	}

	// make the work stream wait until text bytes [0, end) are on the device (streamed batches only)
	int need_text(uint64_t end)
	{
		cudaEvent_t ev = nullptr;
		NTB_BE(batch->upload_until(end, &ev));
		if (ev) {
			NTB_BE(cudaStreamWaitEvent(ws->stream, ev, 0));
		}
		return NTB_OK;
	}

	~CudaBackend()
	{
		// streamed upload copies may still target the workspace's text buffer (error paths): let them land before another
		// caller can check the workspace out
		if (ws && batch && batch->up_stream && !batch->owns_text) {
			cudaStreamSynchronize(batch->up_stream);
		}
		workspace_release(ws);
	}

	static uint64_t env_u64(const char* name, uint64_t dflt)
	{
		const char* v = std::getenv(name);
		return (v && *v) ? std::strtoull(v, nullptr, 10) : dflt;
	}

	// K1b geometry for this filter: region size (log2 slots) and bucket count; false = use the direct scan
	bool binned_geometry(const KParams& kp, uint32_t& rl, uint32_t& nb) const
	{
		if (kp.snv) {
			return false; // every valid window is a site: nothing is probed
		}
		if (bloom->bytes < env_u64("NTB_BIN_MIN_BYTES", 96ull << 20)) {
			return false; // the filter sits in L2 anyway
		}
		const FilterView fv = bloom->view();
		// 64 MB regions, ONE of them hot at a time (the probe CTAs move from bucket to bucket in lock step, pace_lag = 1): half
		// of the L2.  Fewer, larger buckets mean longer runs out of the bin kernel's per-round sort and fewer grid-wide waits
		// in the probe kernel -- measured per 3 G positions on the 4 GiB filter: 16 MB regions / two hot 71.7 ms (round 1's
		// choice), 16 MB / one hot 74.1, 32 MB / one hot 68.9, 64 MB / one hot 66.8, 128 MB 122.6; 32 MB / two hot 85.3.  The
		// 16 GiB filter (256 buckets of 64 MB): two hot 94 ms per 2.5 G positions, one hot 68.6.
		rl = (uint32_t)env_u64("NTB_BIN_REGION_LOG2", bloom->counting ? 26 : 29);
		if (rl < 3) {
			rl = 3;
		}
		while (rl <= 32 && ((fv.mod + (1ull << rl) - 1) >> rl) > (uint64_t)BIN_MAX_BUCKETS) {
			rl++;
		}
		if (rl > 32) {
			return false;
		}
		nb = (uint32_t)((fv.mod + (1ull << rl) - 1) >> rl);
		return true;
	}

	// ---- K1 / K1b, enqueued range by range: the contig groups of a call are consecutive pieces of the text, and the scan
	// of group g+1 is put on the work stream right behind the kernels of group g (walk(): after the round's copies are
	// enqueued), so a text that is still arriving from the host is uploaded beside ALL device work, not just beside the scan.
	KParams scan_kp;
	bool scan_prepared = false, scan_is_binned = false;
	BinArgs binA;
	uint64_t bin_chunk_tiles = 0, bin_cur_tiles = 0, bin_per_buffer = 0, bin_ctas = BIN_CTAS_PER_SM, bin_chunks = 0;
	int bin_probe_ctas = 3;
	bool bin_overlap = false;
	uint64_t scan_hi = 0;       // tiles [0, scan_hi) have been enqueued
	uint64_t scan_prefetch_p = 0; // scan up to this text position behind the next round's copies (0: nothing)
	size_t scan_ev_used = 0;    // event pairs of this call (ws->ev_scan)

	int scan_prepare(const KParams& kp)
	{
		scan_kp = kp;
		scan_prepared = true;
		scan_hi = 0;
		if (batch->n_tiles == 0) {
			return NTB_OK;
		}
		// every piece of a streamed text is put on the copy stream now, in order; the ranges wait for the pieces they read
		{
			cudaEvent_t ev = nullptr;
			NTB_BE(batch->upload_until(batch->total, &ev));
		}
		uint32_t rl = 0, nb = 0;
		scan_is_binned = binned_geometry(kp, rl, nb);
		if (!scan_is_binned) {
			return NTB_OK;
		}
		// NTB_BIN_OVERLAP=1 (experiment, off): two record buffers, the probe kernel of chunk c on a second stream beside the bin
		// kernel of chunk c+1.  Measured 1.6-2.5x SLOWER, also with the two kernels sized to be co-resident
		// (profiles/r01b_scan_stage_tuning_overlap.jsonl, r02_tuning_bin_probe_overlap*.jsonl): the bin kernel's 8 GB/chunk write
		// stream evicts the filter region the probe kernel needs in L2, so the two kernels of a chunk run back to back.
		const uint64_t H = bloom->h;
		bin_overlap = env_u64("NTB_BIN_OVERLAP", 0) != 0;
		// record scratch: fewer, larger chunks amortise the per-bucket pacing of the probe kernel (16 GB: 75 ms, 8 GB: 77 ms,
		// 4 GB: 94 ms per 3 G positions); never more than a quarter of the memory that is free right now
		uint64_t scratch_mb = env_u64("NTB_BIN_SCRATCH_MB", 0);
		if (scratch_mb == 0) {
			scratch_mb = 16384;
			size_t free_b = 0, total_b = 0;
			if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
				const uint64_t have = ((uint64_t)free_b + (uint64_t)ws->cap_records * 8) >> 20;
				scratch_mb = std::max<uint64_t>(64, std::min<uint64_t>(scratch_mb, have / 4));
			}
		}
		const uint64_t budget_records = (scratch_mb << 20) / 8 / (bin_overlap ? 2 : 1);
		uint64_t chunk_tiles = budget_records / (uint64_t)((double)SCAN_TILE * (double)H * 1.06);
		chunk_tiles = std::max<uint64_t>(1, std::min<uint64_t>(chunk_tiles, batch->n_tiles));
		chunk_tiles = std::min<uint64_t>(chunk_tiles, (0xFFFFFFFFull / SCAN_TILE) - 1); // record positions are 32-bit
		uint64_t cap = (uint64_t)((double)(chunk_tiles * SCAN_TILE * H) / (double)nb * 1.04) + 4096;
		cap = env_u64("NTB_BIN_BUCKET_CAP", cap); // testing aid: small rows force the direct-probe overflow path
		cap = (cap + 31) & ~31ull;
		if (cap > 0xFFFFFFE0ull) {
			cap = 0xFFFFFFE0ull;
		}
		bin_per_buffer = (uint64_t)nb * cap;
		const size_t n_buffers = bin_overlap ? 2 : 1;
		if (n_buffers * bin_per_buffer > ws->cap_records) {
			cudaFree(ws->d_records);
			ws->d_records = nullptr;
			ws->cap_records = 0;
			NTB_BE(cudaMalloc((void**)&ws->d_records, n_buffers * bin_per_buffer * 8));
			ws->cap_records = n_buffers * bin_per_buffer;
		}
		if (!ws->d_cursor) {
			NTB_BE(cudaMalloc((void**)&ws->d_cursor, 2 * (BIN_MAX_BUCKETS + 1) * sizeof(uint32_t)));
		}
		if (!ws->stream2) {
			NTB_BE(cudaStreamCreateWithFlags(&ws->stream2, cudaStreamNonBlocking));
			for (int i = 0; i < 2; i++) {
				NTB_BE(cudaEventCreateWithFlags(&ws->ev_bin[i], cudaEventDisableTiming));
				NTB_BE(cudaEventCreateWithFlags(&ws->ev_probe[i], cudaEventDisableTiming));
			}
		}
		std::memset(&binA, 0, sizeof binA);
		binA.scan.filter = bloom->view();
		binA.scan.k = kp.k;
		binA.scan.min_threshold = kp.min_threshold;
		fill_scan_tables(binA.scan, kp.k);
		binA.bucket_cap = (uint32_t)cap;
		binA.n_buckets = nb;
		binA.region_log2 = rl;
		binA.pace_lag = (uint32_t)std::min<uint64_t>(2, std::max<uint64_t>(1, env_u64("NTB_BIN_PACE_LAG", 1)));
		bin_ctas = std::max<uint64_t>(1, env_u64("NTB_BIN_CTAS_PER_SM", BIN_CTAS_PER_SM));
		bin_probe_ctas = (int)env_u64("NTB_BIN_PROBE_CTAS_PER_SM", 4); // 64 MB regions, one hot: 3 -> 66.8 ms, 4 -> 66.1, 5 -> 66.5
		bin_chunk_tiles = chunk_tiles;
		// A batch whose text is still arriving from the host (ntb_polish_batch) starts with a small chunk and grows from there:
		// the first kernel then waits for 128 MB instead of a whole chunk's upload, and as long as a chunk is at most 1.25 x its
		// predecessor a PCIe 5 upload (1.3 x the scan rate) stays ahead of the scan (NTB_BIN_FIRST_CHUNK_MB: 0 = whole chunks)
		bin_cur_tiles = chunk_tiles;
		if (batch->up_src) {
			const uint64_t first_mb = env_u64("NTB_BIN_FIRST_CHUNK_MB", 128);
			if (first_mb) {
				bin_cur_tiles = std::max<uint64_t>(1, std::min<uint64_t>(chunk_tiles, (first_mb << 20) / SCAN_TILE));
			}
		}
		bin_chunks = 0;
		// the bitmap is OR-ed into: cleared once, for the whole batch
		NTB_BE(cudaMemsetAsync(ws->d_visit, 0, batch->n_tiles * SCAN_BITWORDS * 4, ws->stream));
		return NTB_OK;
	}

	// tiles [t0, t1) through K1b
	int scan_binned_tiles(uint64_t t0, uint64_t t1)
	{
		const int sms = sm_count(batch->device);
		cudaStream_t s_probe = bin_overlap ? ws->stream2 : ws->stream;
		BinArgs& A = binA;
		for (uint64_t nt = 0; t0 < t1; t0 += nt, bin_chunks++, bin_cur_tiles = std::min<uint64_t>(bin_chunk_tiles, bin_cur_tiles + bin_cur_tiles / 4 + 1)) {
			nt = std::min<uint64_t>(bin_cur_tiles, t1 - t0);
			const int buf = bin_overlap ? (int)(bin_chunks & 1) : 0;
			A.scan.text = batch->d_text + t0 * SCAN_TILE;
			A.scan.n_tiles = nt;
			A.scan.visit = ws->d_visit + t0 * SCAN_BITWORDS;
			A.chunk_base = t0 * SCAN_TILE;
			A.records = ws->d_records + (size_t)buf * bin_per_buffer;
			A.cursor = ws->d_cursor + (size_t)buf * (BIN_MAX_BUCKETS + 1);
			if (need_text((t0 + nt) * SCAN_TILE) != NTB_OK) {
				return rc;
			}
			if (bin_overlap && bin_chunks >= 2) {
				NTB_BE(cudaStreamWaitEvent(ws->stream, ws->ev_probe[buf], 0)); // the buffer's previous chunk has been probed
			}
			NTB_BE(cudaMemsetAsync(A.cursor, 0, (BIN_MAX_BUCKETS + 1) * sizeof(uint32_t), ws->stream));
			NTB_BE(launch_bin(A, bloom->counting != 0, (int)std::min<uint64_t>(nt, (uint64_t)sms * bin_ctas), ws->stream));
			NTB_BE(cudaEventRecord(ws->ev_bin[buf], ws->stream));
			NTB_BE(cudaStreamWaitEvent(s_probe, ws->ev_bin[buf], 0));
			NTB_BE(launch_probe_bin(A, bloom->counting != 0, bin_probe_ctas, s_probe));
			NTB_BE(cudaEventRecord(ws->ev_probe[buf], s_probe));
			launches += 2;
		}
		// what follows on the work stream needs every probe done
		for (int i = 0; bin_overlap && i < 2 && (uint64_t)i < bin_chunks; i++) {
			NTB_BE(cudaStreamWaitEvent(ws->stream, ws->ev_probe[i], 0));
		}
		return NTB_OK;
	}

	// tiles [t0, t1) through K1 (filters that fit L2, -s 1)
	int scan_direct_tiles(uint64_t t0, uint64_t t1)
	{
		ScanArgs a;
		std::memset(&a, 0, sizeof a);
		a.text = batch->d_text + t0 * SCAN_TILE;
		a.n_tiles = t1 - t0;
		a.filter = bloom->view();
		a.k = scan_kp.k;
		a.min_threshold = scan_kp.min_threshold;
		a.snv = (uint32_t)scan_kp.snv;
		a.visit = ws->d_visit + t0 * SCAN_BITWORDS;
		fill_scan_tables(a, scan_kp.k);
		const int grid = (int)std::min<uint64_t>(a.n_tiles, (uint64_t)sm_count(batch->device) * 3);
		if (need_text(t1 * SCAN_TILE) != NTB_OK) {
			return rc;
		}
		if (grid > 0) {
			NTB_BE(launch_scan(a, bloom->counting != 0, false, grid, ws->stream));
			launches++;
		}
		return NTB_OK;
	}

	// enqueue the scan of every tile below text position p_end that has not been enqueued yet
	int scan_until_impl(uint64_t p_end)
	{
		const uint64_t t1 = std::min<uint64_t>(batch->n_tiles, (std::min<uint64_t>(p_end, batch->total) + SCAN_TILE - 1) / SCAN_TILE);
		if (t1 <= scan_hi) {
			return NTB_OK;
		}
		while (ws->ev_scan.size() < 2 * (scan_ev_used + 1)) {
			cudaEvent_t ev = nullptr;
			NTB_BE(cudaEventCreate(&ev));
			ws->ev_scan.push_back(ev);
		}
		NTB_BE(cudaEventRecord(ws->ev_scan[2 * scan_ev_used], ws->stream));
		if ((scan_is_binned ? scan_binned_tiles(scan_hi, t1) : scan_direct_tiles(scan_hi, t1)) != NTB_OK) {
			return rc;
		}
		NTB_BE(cudaEventRecord(ws->ev_scan[2 * scan_ev_used + 1], ws->stream));
		scan_ev_used++;
		scan_hi = t1;
		return NTB_OK;
	}

	void scan_begin(const KParams& kp)
	{
		if (rc == NTB_OK) {
			scan_prepare(kp);
		}
	}

	void scan_until(uint64_t p_end)
	{
		if (rc == NTB_OK && scan_prepared) {
			scan_until_impl(p_end);
		}
	}

	void scan_prefetch(uint64_t p_end)
	{
		scan_prefetch_p = p_end;
	}

	bool text_streaming() const { return batch && batch->up_src != nullptr; }

	// the work stream is idle: add up what the ranges took (waits for the text included)
	void scan_end()
	{
		if (rc != NTB_OK || !ws) {
			return;
		}
		if (scan_ev_used && cudaEventSynchronize(ws->ev_scan[2 * scan_ev_used - 1]) != cudaSuccess) {
			cuda_err(cudaGetLastError(), "cudaEventSynchronize(scan)");
			return;
		}
		for (size_t i = 0; i < scan_ev_used; i++) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, ws->ev_scan[2 * i], ws->ev_scan[2 * i + 1]) == cudaSuccess) {
				ms_scan += ms;
			}
			if (std::getenv("NTB_DEBUG_TASKS")) {
				float a = 0;
				cudaEventElapsedTime(&a, ws->ev_scan[0], ws->ev_scan[2 * i]);
				std::fprintf(stderr, "[ntb] device: scan range %zu %.1f - %.1f ms\n", i, a, a + ms);
			}
		}
		scan_ev_used = 0;
	}

	Task* task_buffer(size_t n)
	{
		if (rc != NTB_OK) {
			return nullptr;
		}
		if (n > ws->cap_h_tasks) {
			cudaFreeHost(ws->h_tasks);
			cudaFreeHost(ws->h_results);
			ws->h_tasks = nullptr;
			ws->h_results = nullptr;
			ws->cap_h_tasks = 0;
			const size_t want = n + n / 8 + 1024;
			if (cudaHostAlloc((void**)&ws->h_tasks, want * sizeof(Task), cudaHostAllocDefault) != cudaSuccess ||
			    cudaHostAlloc((void**)&ws->h_results, want * sizeof(TaskResult), cudaHostAllocDefault) != cudaSuccess) {
				err = "cudaHostAlloc(tasks) failed";
				rc = NTB_ENOMEM;
				return nullptr;
			}
			ws->cap_h_tasks = want;
		}
		return ws->h_tasks;
	}

	// Pre-evaluation in front of the first walker round (kernels.cuh: launch_heads / launch_presite): fills ws->d_table.
	// Sized from the batch: a head every ~800 bases at the usual error rates; lists and table that run full only cost speed
	// (the walkers evaluate what has no record).
	int presites(WalkArgs& a, cudaStream_t stream)
	{
		if (env_u64("NTB_NO_PRESITE", 0)) {
			return NTB_OK;
		}
		const uint64_t total = batch->total;
		const size_t want_items = (size_t)std::min<uint64_t>(0x7FFFFFF0ull, std::max<uint64_t>(1u << 14, total / 96));
		const size_t want_pending = (size_t)std::min<uint64_t>(0x7FFFFFF0ull, std::max<uint64_t>(1u << 12, total / 384));
		size_t slots = 1u << 12;
		// -s 1 files a record wherever a variant has support: with a well-filled counting filter that is a large share of all
		// positions, so the table gets a slot per two positions as far as a third of the free memory allows
		uint64_t dflt_slots = total / 64;
		if (a.kp.snv) {
			size_t free_b = 0, total_b = 0;
			uint64_t room = 1ull << 24;
			if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
				room = ((uint64_t)free_b + (uint64_t)ws->cap_table * sizeof(SiteRec)) / 3 / sizeof(SiteRec);
			}
			dflt_slots = std::min<uint64_t>(total / 2, room);
			uint64_t p2 = 1;
			while (p2 * 2 <= dflt_slots) {
				p2 *= 2; // (the loop below rounds up: stay below the memory bound)
			}
			dflt_slots = p2;
		}
		const uint64_t want_slots = std::min<uint64_t>(1ull << 31, env_u64("NTB_SITE_TABLE_SLOTS", dflt_slots));
		while (slots < want_slots) {
			slots <<= 1;
		}
		if (slots > ws->cap_table) {
			cudaFree(ws->d_table);
			ws->d_table = nullptr;
			ws->cap_table = 0;
			NTB_BE(cudaMalloc((void**)&ws->d_table, slots * sizeof(SiteRec)));
			ws->cap_table = slots;
		}
		if (want_items > ws->cap_items) {
			cudaFree(ws->d_items);
			ws->d_items = nullptr;
			ws->cap_items = 0;
			NTB_BE(cudaMalloc((void**)&ws->d_items, want_items * sizeof(uint2)));
			ws->cap_items = want_items;
		}
		if (want_pending > ws->cap_pending) {
			cudaFree(ws->d_pending);
			ws->d_pending = nullptr;
			ws->cap_pending = 0;
			NTB_BE(cudaMalloc((void**)&ws->d_pending, want_pending * sizeof(PendingSite)));
			ws->cap_pending = want_pending;
		}
		if (!ws->ev_pre0) {
			NTB_BE(cudaEventCreate(&ws->ev_pre0));
			NTB_BE(cudaEventCreate(&ws->ev_pre1));
			NTB_BE(cudaEventCreate(&ws->ev_pre_mid));
			NTB_BE(cudaHostAlloc((void**)&ws->h_ctr_pre, sizeof(Counters), cudaHostAllocDefault));
		}
		a.table = ws->d_table;
		a.table_mask = (uint32_t)(slots - 1);
		a.items = ws->d_items;
		a.items_cap = (uint32_t)ws->cap_items;
		a.pending = ws->d_pending;
		a.pending_cap = (uint32_t)ws->cap_pending;
		NTB_BE(cudaEventRecord(ws->ev_pre0, stream));
		if (!table_slots) {
			// (the contig groups of a call share the table)
			NTB_BE(cudaMemsetAsync(ws->d_table, 0, slots * sizeof(SiteRec), stream));
		}
		NTB_BE(cudaMemsetAsync(ws->d_ctr, 0, sizeof(Counters), stream));
		if (a.kp.snv) {
			// K3: every valid position is a site; the walkers get a bitmap of the ones that do something
			const size_t words = batch->n_tiles * SCAN_BITWORDS + 16;
			if (!use_visit2) {
				if (words > ws->cap_visit2) {
					cudaFree(ws->d_visit2);
					ws->d_visit2 = nullptr;
					ws->cap_visit2 = 0;
					NTB_BE(cudaMalloc((void**)&ws->d_visit2, words * 4));
					ws->cap_visit2 = words;
				}
				NTB_BE(cudaMemsetAsync(ws->d_visit2, 0, words * 4, stream)); // (the contig groups of a call share it)
				use_visit2 = true;
			}
			NTB_BE(launch_snv_dense(a, ws->d_visit2, stream));
			NTB_BE(cudaEventRecord(ws->ev_pre_mid, stream));
		} else {
			NTB_BE(launch_heads(a, stream));
			NTB_BE(launch_presite(a, false, stream));
			NTB_BE(cudaEventRecord(ws->ev_pre_mid, stream));
			NTB_BE(launch_presite(a, true, stream));
		}
		NTB_BE(cudaMemcpyAsync(ws->h_ctr_pre, ws->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
		NTB_BE(cudaEventRecord(ws->ev_pre1, stream));
		launches += a.kp.snv ? 1u : presite_launch_count();
		pre_timed = true;
		table_slots = slots;
		return NTB_OK;
	}

	int walk(const KParams& kp, size_t n, bool first_round_of_group, const TaskResult** res_out, const Event** ev_out, size_t* n_ev_out)
	{
		if (rc != NTB_OK) {
			return rc;
		}
		*res_out = ws->h_results;
		*ev_out = nullptr;
		*n_ev_out = 0;
		if (n == 0) {
			Event* none = nullptr;
			if (round_block(0, &none) != NTB_OK) {
				return rc;
			}
			n_rounds++;
			return NTB_OK;
		}
		if (n > 0xFFFFFFF0ULL) {
			err = "too many segments in one batch";
			return rc = NTB_EINVAL;
		}
		cudaStream_t stream = ws->stream;
		if (n > ws->cap_tasks) {
			cudaFree(ws->d_tasks);
			cudaFree(ws->d_results);
			cudaFree(ws->d_order);
			ws->d_tasks = nullptr;
			ws->d_results = nullptr;
			ws->d_order = nullptr;
			ws->cap_tasks = 0;
			const size_t want = n + n / 8 + 1024;
			NTB_BE(cudaMalloc((void**)&ws->d_order, want * sizeof(uint32_t)));
			NTB_BE(cudaMalloc((void**)&ws->d_tasks, want * sizeof(Task)));
			NTB_BE(cudaMalloc((void**)&ws->d_results, want * sizeof(TaskResult)));
			ws->cap_tasks = want;
		}
		if (ws->cap_events == 0) {
			// sized for the recipe's ~1.1e-3 edits per base with headroom; grown on overflow
			const size_t want = std::max<size_t>(1u << 16, (size_t)(batch->total / 256));
			NTB_BE(cudaMalloc((void**)&ws->d_events, want * sizeof(Event)));
			NTB_BE(cudaMalloc((void**)&ws->d_events_sorted, want * sizeof(Event)));
			ws->cap_events = want;
		}
		static_assert(sizeof(Task) % 16 == 0, "tasks are fetched in 16-byte units");
		NTB_BE(launch_fetch_host(ws->d_tasks, ws->h_tasks, n * sizeof(Task), stream));
		launches++;
		const FilterView fb = bloom->view();
		FilterView fr;
		std::memset(&fr, 0, sizeof fr);
		if (rep) {
			fr = rep->view();
		}
		WalkArgs wa;
		std::memset(&wa, 0, sizeof wa);
		wa.text = batch->d_text;
		wa.visit = ws->d_visit;
		wa.bloom = fb;
		wa.rep = fr;
		wa.kp = kp;
		wa.tasks = ws->d_tasks;
		wa.order = ws->d_order;
		wa.results = ws->d_results;
		wa.n_tasks = (uint32_t)n;
		wa.ctr = ws->d_ctr;
		wa.sm_count = sm_count(batch->device);
		if (first_round_of_group) {
			// a group's first round: its tasks cover every contig of the group
			if (presites(wa, stream) != NTB_OK) {
				return rc;
			}
		}
		if (table_slots) {
			// (later rounds look the same records up)
			wa.table = ws->d_table;
			wa.table_mask = (uint32_t)(table_slots - 1);
		}
		if (use_visit2) {
			wa.visit = ws->d_visit2;
		}
		for (;;) {
			wa.events = ws->d_events;
			wa.ev_cap = (uint32_t)std::min<size_t>(ws->cap_events, 0xFFFFFFF0u);
			NTB_BE(cudaMemsetAsync(ws->d_ctr, 0, sizeof(Counters), stream));
			NTB_BE(cudaEventRecord(ws->ev0, stream));
			NTB_BE(launch_walk(wa, stream));
			launches += 2; // order_tasks_kernel + walk_kernel
			NTB_BE(cudaEventRecord(ws->ev1, stream));
			NTB_BE(cudaMemcpyAsync(ws->h_ctr, ws->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
			NTB_BE(cudaStreamSynchronize(stream));
			const Counters ctr = *ws->h_ctr;
			float ms = 0;
			NTB_BE(cudaEventElapsedTime(&ms, ws->ev0, ws->ev1));
			ms_walk += ms;
			if (std::getenv("NTB_DEBUG_TASKS") && !ws->ev_scan.empty()) {
				// device timeline of the call, relative to the first scanned range
				float a = 0, b = 0, c = 0, d = 0;
				cudaEventElapsedTime(&c, ws->ev_scan[0], ws->ev0);
				cudaEventElapsedTime(&d, ws->ev_scan[0], ws->ev1);
				if (pre_timed) {
					cudaEventElapsedTime(&a, ws->ev_scan[0], ws->ev_pre0);
					cudaEventElapsedTime(&b, ws->ev_scan[0], ws->ev_pre1);
				}
				std::fprintf(stderr, "[ntb] device: pre-evaluation %.1f - %.1f ms, walk %.1f - %.1f ms\n", a, b, c, d);
			}
			if (pre_timed) {
				pre_timed = false;
				NTB_BE(cudaEventElapsedTime(&ms, ws->ev_pre0, ws->ev_pre1));
				ms_pre += ms;
				if (std::getenv("NTB_DEBUG_TASKS")) {
					const Counters& pc = *ws->h_ctr_pre;
					float ms1 = 0;
					cudaEventElapsedTime(&ms1, ws->ev_pre0, ws->ev_pre_mid);
					std::fprintf(stderr, "[ntb] pre-evaluation: first pass %.2f ms, second pass %.2f ms\n", ms1, ms - ms1);
					std::fprintf(stderr, "[ntb] pre-evaluation: %.2f ms, %u heads (cap %zu), %u pending (cap %zu), %u records dropped, table %zu slots\n", ms,
					             pc.n_items, ws->cap_items, pc.n_pending, ws->cap_pending, pc.n_dropped, table_slots);
				}
			}
			if (ctr.overflow) {
				cudaFree(ws->d_events);
				cudaFree(ws->d_events_sorted);
				ws->d_events = nullptr;
				ws->d_events_sorted = nullptr;
				const size_t want = ws->cap_events * 4;
				ws->cap_events = 0;
				NTB_BE(cudaMalloc((void**)&ws->d_events, want * sizeof(Event)));
				NTB_BE(cudaMalloc((void**)&ws->d_events_sorted, want * sizeof(Event)));
				ws->cap_events = want;
				continue;
			}
			Event* h_round = nullptr;
			if (round_block(ctr.n_events, &h_round) != NTB_OK) {
				return rc;
			}
			if (!ws->stream_d2h) {
				NTB_BE(cudaStreamCreateWithFlags(&ws->stream_d2h, cudaStreamNonBlocking));
				NTB_BE(cudaEventCreateWithFlags(&ws->ev_compact, cudaEventDisableTiming));
			}
			NTB_BE(cudaEventRecord(ws->ev0, stream));
			if (ctr.n_events) {
				// group the events by walker on the device, then bring them over
				NTB_BE(launch_compact_events(ws->d_events, ws->d_events_sorted, ws->d_results, (uint32_t)n, ws->d_ctr, stream));
				launches++;
			}
			// the copies run on their own stream (100 MB of events per 2.5 Gbp: 2 ms): the work stream goes straight on with the
			// next group's scan.  Nothing on the work stream touches these buffers before the host has seen the copies land.
			NTB_BE(cudaEventRecord(ws->ev_compact, stream));
			NTB_BE(cudaStreamWaitEvent(ws->stream_d2h, ws->ev_compact, 0));
			if (ctr.n_events) {
				NTB_BE(cudaMemcpyAsync(h_round, ws->d_events_sorted, (size_t)ctr.n_events * sizeof(Event),
				                       cudaMemcpyDeviceToHost, ws->stream_d2h));
			}
			NTB_BE(cudaMemcpyAsync(ws->h_results, ws->d_results, n * sizeof(TaskResult), cudaMemcpyDeviceToHost, ws->stream_d2h));
			NTB_BE(cudaEventRecord(ws->ev1, ws->stream_d2h));
			if (scan_prefetch_p) {
				// the next contig group's scan goes right behind this round's copies: the device works on it while the host
				// stitches, and a text that is still being uploaded had the whole group's device phase to arrive
				const uint64_t p = scan_prefetch_p;
				scan_prefetch_p = 0;
				if (scan_until_impl(p) != NTB_OK) {
					return rc;
				}
			}
			NTB_BE(cudaEventSynchronize(ws->ev1));
			NTB_BE(cudaEventElapsedTime(&ms, ws->ev0, ws->ev1));
			ms_d2h += ms;
			*res_out = ws->h_results;
			*ev_out = h_round;
			*n_ev_out = ctr.n_events;
			n_rounds++;
			if (std::getenv("NTB_DEBUG_TASKS")) {
				debug_tasks(n, ctr);
			}
			return NTB_OK;
		}
	}

	// diagnostics: the slowest walkers of this launch and the cycles by number of sites
	void debug_tasks(size_t n, const Counters& ctr)
	{
		if (std::getenv("NTB_DEBUG_TASKS")[0] == '2') {
			return; // timeline only
		}
		const TaskResult* results = ws->h_results;
		const Task* tasks = ws->h_tasks;
		std::vector<size_t> idx(n);
		for (size_t i = 0; i < n; i++) {
			idx[i] = i;
		}
		const size_t top = std::min<size_t>(6, n);
		std::partial_sort(idx.begin(), idx.begin() + (long)top, idx.end(),
		                  [&](size_t a, size_t b) { return results[a].kcycles > results[b].kcycles; });
		unsigned long long tot = 0;
		for (size_t i = 0; i < n; i++) {
			tot += results[i].kcycles;
		}
		std::fprintf(stderr, "[ntb] walk launch: %zu tasks, %.2f ms, sum %.1f Mcycles\n", n, ms_walk, tot / 1024.0);
		static const char* names[16] = { "loop_head", "next_visit", "fill_cache", "seed", "lookahead", "dirty_misc", "evaluate_site", "advance",
			                             "site_begin+linearise", "compute_plain", "phase1", "candidates", "try_indels", "commit", "-", "-" };
		for (int q = 0; q < 14; q++) {
			if (ctr.prof[q]) {
				std::fprintf(stderr, "[ntb]   phase %-22s %10.1f Mcycles\n", names[q], ctr.prof[q] / 1048576.0);
			}
		}
		if (ctr.prof[14]) {
			std::fprintf(stderr, "[ntb]   filter probes issued %llu, tryIndels calls %llu\n", ctr.prof[14], ctr.prof[15]);
		}
		const unsigned edges[7] = { 0, 1, 4, 8, 16, 32, 1u << 30 };
		for (int b = 0; b < 6; b++) {
			unsigned long long cyc = 0, cnt = 0, sites = 0, evs = 0;
			for (size_t i = 0; i < n; i++) {
				if (results[i].n_sites >= edges[b] && results[i].n_sites < edges[b + 1]) {
					cyc += results[i].kcycles;
					cnt++;
					sites += results[i].n_sites;
					evs += results[i].n_events;
				}
			}
			std::fprintf(stderr, "[ntb]   sites in [%u,%u): %llu tasks, %llu sites, %llu events, %.1f Mcycles\n", edges[b], edges[b + 1], cnt, sites, evs,
			             cyc / 1024.0);
		}
		for (size_t q = 0; q < top; q++) {
			const TaskResult& r = results[idx[q]];
			const Task& t = tasks[idx[q]];
			std::fprintf(stderr, "[ntb]   task %zu contig %u [%u,%u) kcycles %u sites %u events %u end %u status %u\n", idx[q], t.contig, t.start, t.end,
			             r.kcycles, r.n_sites, r.n_events, r.end_pos, r.status);
		}
	}
#undef NTB_BE
};

int
polish_common(ntb_filter* bloom, ntb_filter* rep, const ntb_params* p, ntb_batch* batch, char* host_bases, ntb_result** out)
{
	if (!bloom || !p || !batch || !out) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	if (bloom->device != batch->device || (rep && rep->device != batch->device)) {
		return fail(NTB_EINVAL, "filter and batch live on different devices");
	}
	if (rep && rep->k != bloom->k) { // ntedit.cpp:2580-2585
		return fail(NTB_EINVAL, "secondary Bloom filter k size (" + std::to_string(rep->k) + ") is different than main Bloom filter k size (" +
		                            std::to_string(bloom->k) + ")");
	}
	int rc = select_device(batch->device);
	if (rc != NTB_OK) {
		return rc;
	}
	KParams kp;
	std::string err;
	rc = make_kparams(*p, bloom->k, bloom->h, rep ? rep->h : 0, bloom->counting != 0, kp, err);
	if (rc != NTB_OK) {
		return fail(rc, err);
	}
	ntb_result* res = new (std::nothrow) ntb_result();
	if (!res) {
		return fail(NTB_ENOMEM, "out of memory");
	}
	CudaBackend be;
	be.bloom = bloom;
	be.rep = rep;
	be.batch = batch;
	rc = be.init();
	if (rc == NTB_OK) {
		rc = polish_run(be, kp, *p, host_bases, batch->offsets.data(), batch->offsets.size() - 1, res->impl, err);
		if (rc == NTB_OK && be.rc != NTB_OK) {
			rc = be.rc;
			err = be.err;
		}
	} else {
		err = be.err;
	}
	if (rc != NTB_OK) {
		delete res;
		return fail(rc, err);
	}
	res->impl.stats.ms_scan = be.ms_scan;
	res->impl.stats.ms_walk = be.ms_walk;
	res->impl.stats.ms_pre = be.ms_pre;
	res->impl.stats.ms_d2h = be.ms_d2h;
	if (batch->up_src && batch->up_end && batch->up_issued == batch->total && cudaEventSynchronize(batch->up_end) == cudaSuccess) {
		cudaEventElapsedTime(&batch->ms_h2d, batch->up_begin, batch->up_end); // includes the waits between pieces
	}
	res->impl.stats.ms_h2d = batch->ms_h2d;
	res->impl.stats.kernel_launches = be.launches;
	*out = res;
	return NTB_OK;
}

void
strbuf_append(ntb_strbuf* b, const std::string& s)
{
	if (!b || s.empty()) {
		return;
	}
	if (b->len + s.size() + 1 > b->cap) {
		size_t c = b->cap ? b->cap : 4096;
		while (c < b->len + s.size() + 1) {
			c *= 2;
		}
		b->data = (char*)std::realloc(b->data, c);
		b->cap = c;
	}
	std::memcpy(b->data + b->len, s.data(), s.size());
	b->len += s.size();
	b->data[b->len] = 0;
}

} // namespace

extern "C" {

const char*
ntb_last_error(void)
{
	return g_error.c_str();
}

const char*
ntb_version(void)
{
	return "ntedit_b200 0.1 (ntEdit v2.1.1 hot path, sm_100a)";
}

int
ntb_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		return 0;
	}
	return n;
}

void*
ntb_host_alloc(size_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

void
ntb_host_free(void* p)
{
	if (p) {
		cudaFreeHost(p);
	}
}

void
ntb_params_init(ntb_params* p)
{
	if (!p) {
		return;
	}
	std::memset(p, 0, sizeof *p);
	p->jump = 3;
	p->max_insertions = 5;
	p->max_deletions = 5;
	p->edit_threshold = 9.0f;
	p->missing_threshold = 5.0f;
	p->edit_ratio = 0.5f;
	p->missing_ratio = 0.5f;
	p->min_threshold = 1;
	p->max_threshold = 255;
	p->min_contig_len = 100;
}

// ---------------------------------------------------------------- filters
int
ntb_filter_create(uint64_t bytes, uint32_t k, uint32_t hash_num, int counting, int device, ntb_filter** out)
{
	if (!out || bytes == 0 || k == 0 || hash_num == 0 || hash_num > HMAX) {
		return fail(NTB_EINVAL, "bad filter geometry");
	}
	int rc = select_device(device);
	if (rc != NTB_OK) {
		return rc;
	}
	ntb_filter* f = new (std::nothrow) ntb_filter();
	if (!f) {
		return fail(NTB_ENOMEM, "out of memory");
	}
	f->bytes = bytes;
	f->k = k;
	f->h = hash_num;
	f->counting = counting ? 1 : 0;
	f->device = device;
	const uint64_t padded = (bytes + 15) / 16 * 16 + 16;
	cudaError_t e = cudaMalloc((void**)&f->d, padded);
	if (e == cudaSuccess) {
		e = cudaMemset(f->d, 0, padded);
	}
	if (e != cudaSuccess) {
		cudaFree(f->d);
		delete f;
		return cuda_fail(e, "cudaMalloc(filter)");
	}
	*out = f;
	return NTB_OK;
}

int
ntb_filter_wrap_device(void* dev_bytes, uint64_t bytes, uint32_t k, uint32_t hash_num, int counting, int device, ntb_filter** out)
{
	if (!out || !dev_bytes || bytes == 0 || k == 0 || hash_num == 0 || hash_num > HMAX) {
		return fail(NTB_EINVAL, "bad filter geometry");
	}
	if (((uintptr_t)dev_bytes & 15u) != 0) {
		return fail(NTB_EINVAL, "wrapped filter bytes must be 16-byte aligned");
	}
	int rc = select_device(device);
	if (rc != NTB_OK) {
		return rc;
	}
	ntb_filter* f = new (std::nothrow) ntb_filter();
	if (!f) {
		return fail(NTB_ENOMEM, "out of memory");
	}
	f->d = (uint8_t*)dev_bytes;
	f->bytes = bytes;
	f->k = k;
	f->h = hash_num;
	f->counting = counting ? 1 : 0;
	f->device = device;
	f->owned = false;
	*out = f;
	return NTB_OK;
}

int
ntb_filter_replicate(ntb_filter* src, int device, ntb_filter** out)
{
	if (!src || !out) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	ntb_filter* f = nullptr;
	int rc = ntb_filter_create(src->bytes, src->k, src->h, src->counting, device, &f); // selects `device`
	if (rc != NTB_OK) {
		return rc;
	}
	if (device != src->device) {
		int can = 0;
		if (cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can) {
			const cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0); // direct NVLink path; without it the copy is staged
			if (pe != cudaSuccess) {
				cudaGetLastError(); // already enabled (or refused): cudaMemcpyPeer works either way
			}
		}
	}
	const cudaError_t e = cudaMemcpyPeer(f->d, device, src->d, src->device, src->bytes);
	if (e != cudaSuccess) {
		ntb_filter_free(f);
		return cuda_fail(e, "cudaMemcpyPeer(filter)");
	}
	f->fpr = src->fpr;
	*out = f;
	return NTB_OK;
}

int
ntb_filter_load(const char* path, int device, ntb_filter** out)
{
	if (!path || !out) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	FilterHeader hdr;
	std::string err;
	if (!read_filter_header(path, hdr, err)) {
		return fail(NTB_EIO, err);
	}
	ntb_filter* f = nullptr;
	int rc = ntb_filter_create(hdr.bytes, hdr.k, hdr.hash_num, hdr.counting ? 1 : 0, device, &f);
	if (rc != NTB_OK) {
		return rc;
	}
	// stream the payload through two pinned staging buffers: the file read of one piece overlaps the upload of the other
	const size_t chunk = (size_t)64 << 20;
	char* stage[2] = { nullptr, nullptr };
	cudaStream_t cs = nullptr;
	cudaEvent_t done_ev[2] = { nullptr, nullptr };
	cudaError_t e = cudaMallocHost((void**)&stage[0], chunk);
	if (e == cudaSuccess) {
		e = cudaMallocHost((void**)&stage[1], chunk);
	}
	if (e == cudaSuccess) {
		e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
	}
	for (int i = 0; i < 2 && e == cudaSuccess; i++) {
		e = cudaEventCreateWithFlags(&done_ev[i], cudaEventDisableTiming);
	}
	auto cleanup = [&]() {
		if (cs) {
			cudaStreamSynchronize(cs);
			cudaStreamDestroy(cs);
		}
		for (int i = 0; i < 2; i++) {
			if (done_ev[i]) {
				cudaEventDestroy(done_ev[i]);
			}
			cudaFreeHost(stage[i]);
		}
	};
	if (e != cudaSuccess) {
		cleanup();
		ntb_filter_free(f);
		return cuda_fail(e, "cudaMallocHost");
	}
	FILE* fp = std::fopen(path, "rb");
	bool ok = fp && fseeko(fp, (off_t)hdr.data_offset, SEEK_SET) == 0;
	uint64_t done = 0;
	for (int i = 0; ok && e == cudaSuccess && done < hdr.bytes; i ^= 1) {
		e = cudaEventSynchronize(done_ev[i]); // the buffer's previous upload (a fresh event is complete)
		if (e != cudaSuccess) {
			break;
		}
		const size_t want = (size_t)std::min<uint64_t>(chunk, hdr.bytes - done);
		if (std::fread(stage[i], 1, want, fp) != want) {
			ok = false;
			break;
		}
		e = cudaMemcpyAsync(f->d + done, stage[i], want, cudaMemcpyHostToDevice, cs);
		if (e == cudaSuccess) {
			e = cudaEventRecord(done_ev[i], cs);
		}
		done += want;
	}
	if (fp) {
		std::fclose(fp);
	}
	if (e == cudaSuccess) {
		e = cudaStreamSynchronize(cs);
	}
	cleanup();
	if (e != cudaSuccess) {
		ntb_filter_free(f);
		return cuda_fail(e, "cudaMemcpy(filter)");
	}
	if (!ok) {
		ntb_filter_free(f);
		return fail(NTB_EIO, std::string("truncated Bloom filter file ") + path);
	}
	*out = f;
	return NTB_OK;
}

int
ntb_filter_get_info(ntb_filter* f, ntb_filter_info* info)
{
	if (!f || !info) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	int rc = select_device(f->device);
	if (rc != NTB_OK) {
		return rc;
	}
	if (f->fpr < 0) {
		unsigned long long* d_cnt = nullptr;
		NTB_CUDA(cudaMalloc((void**)&d_cnt, 8));
		NTB_CUDA(cudaMemset(d_cnt, 0, 8));
		occupancy_kernel<<<sm_count(f->device) * 8, 256>>>(f->d, f->bytes, f->counting, d_cnt);
		unsigned long long cnt = 0;
		cudaError_t e = cudaMemcpy(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost);
		cudaFree(d_cnt);
		if (e != cudaSuccess) {
			return cuda_fail(e, "occupancy_kernel");
		}
		const double denom = f->counting ? (double)f->bytes : (double)f->bytes * 8.0;
		f->fpr = std::pow((double)cnt / denom, (double)f->h);
	}
	info->bytes = f->bytes;
	info->k = f->k;
	info->hash_num = f->h;
	info->counting = f->counting;
	info->device = f->device;
	info->fpr = f->fpr;
	return NTB_OK;
}

void*
ntb_filter_device_ptr(ntb_filter* f)
{
	return f ? f->d : nullptr;
}

int
ntb_filter_insert_batch(ntb_filter* f, const ntb_batch* b)
{
	if (!f || !b) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	if (f->device != b->device) {
		return fail(NTB_EINVAL, "filter and batch live on different devices");
	}
	if (f->k < 2 || f->k > KMAX) {
		return fail(NTB_EINVAL, "unsupported k");
	}
	if (!f->owned && (f->bytes & 3u) != 0) {
		return fail(NTB_EINVAL, "inserting into a wrapped filter needs a size that is a multiple of 4 bytes");
	}
	int rc = select_device(f->device);
	if (rc != NTB_OK) {
		return rc;
	}
	ntb_params up;
	ntb_params_init(&up);
	KParams kp;
	std::memset(&kp, 0, sizeof kp);
	kp.k = f->k;
	kp.h = f->h;
	const uint64_t seeds[4] = { SEED_A, SEED_C, SEED_G, SEED_T };
	for (int c = 0; c < 4; c++) {
		kp.seed_rot_k[c] = sroln(seeds[c], f->k);
		kp.seed_rot_k1[c] = sroln(seeds[c], f->k - 1);
	}
	const uint64_t strips = (b->total + 255) / 256;
	const unsigned block = 128;
	const uint64_t grid = (strips + block - 1) / block;
	if (grid > 0x7FFFFFFFULL) {
		return fail(NTB_EINVAL, "batch too large");
	}
	if (grid > 0) {
		insert_kernel<<<(unsigned)grid, block>>>(b->d_text, b->total, f->d, f->view(), kp);
		NTB_CUDA(cudaGetLastError());
		NTB_CUDA(cudaDeviceSynchronize());
	}
	f->fpr = -1.0;
	return NTB_OK;
}

int
ntb_filter_insert(ntb_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_contigs)
{
	if (!f) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	ntb_batch* b = nullptr;
	int rc = ntb_batch_upload(bases, offsets, n_contigs, f->device, &b);
	if (rc != NTB_OK) {
		return rc;
	}
	rc = ntb_filter_insert_batch(f, b);
	ntb_batch_free(b);
	return rc;
}

int
ntb_filter_download(ntb_filter* f, void* host_dst, uint64_t bytes)
{
	if (!f || !host_dst || bytes > f->bytes) {
		return fail(NTB_EINVAL, "bad argument");
	}
	int rc = select_device(f->device);
	if (rc != NTB_OK) {
		return rc;
	}
	NTB_CUDA(cudaMemcpy(host_dst, f->d, bytes, cudaMemcpyDeviceToHost));
	return NTB_OK;
}

int
ntb_filter_save(ntb_filter* f, const char* path)
{
	if (!f || !path) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	int rc = select_device(f->device);
	if (rc != NTB_OK) {
		return rc;
	}
	FILE* fp = std::fopen(path, "wb");
	if (!fp) {
		return fail(NTB_EIO, std::string("cannot open ") + path + " for writing");
	}
	const std::string hdr = format_filter_header(f->bytes, f->k, f->h, f->counting != 0);
	bool ok = std::fwrite(hdr.data(), 1, hdr.size(), fp) == hdr.size();
	const size_t chunk = (size_t)64 << 20;
	std::vector<char> stage(std::min<uint64_t>(chunk, f->bytes));
	uint64_t done = 0;
	cudaError_t e = cudaSuccess;
	while (ok && done < f->bytes) {
		const size_t want = (size_t)std::min<uint64_t>(chunk, f->bytes - done);
		e = cudaMemcpy(stage.data(), f->d + done, want, cudaMemcpyDeviceToHost);
		if (e != cudaSuccess) {
			break;
		}
		ok = std::fwrite(stage.data(), 1, want, fp) == want;
		done += want;
	}
	std::fclose(fp);
	if (e != cudaSuccess) {
		return cuda_fail(e, "cudaMemcpy(filter)");
	}
	return ok ? NTB_OK : fail(NTB_EIO, std::string("short write to ") + path);
}

void
ntb_filter_free(ntb_filter* f)
{
	if (!f) {
		return;
	}
	if (f->owned && f->d) {
		cudaSetDevice(f->device);
		cudaFree(f->d);
	}
	delete f;
}

// ---------------------------------------------------------------- batches
int
ntb_batch_upload(const char* bases, const uint64_t* offsets, uint64_t n_contigs, int device, ntb_batch** out)
{
	if (!bases || !out) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	int rc = check_offsets(offsets, n_contigs);
	if (rc != NTB_OK) {
		return rc;
	}
	for (uint64_t c = 0; c < n_contigs; c++) {
		if (bases[offsets[c + 1] - 1] != 0) {
			return fail(NTB_EINVAL, "contig " + std::to_string(c) + " is not NUL-terminated inside the batch buffer");
		}
	}
	rc = select_device(device);
	if (rc != NTB_OK) {
		return rc;
	}
	ntb_batch* b = new (std::nothrow) ntb_batch();
	if (!b) {
		return fail(NTB_ENOMEM, "out of memory");
	}
	b->device = device;
	b->offsets.assign(offsets, offsets + n_contigs + 1);
	rc = batch_alloc(b, offsets[n_contigs]);
	if (rc != NTB_OK) {
		ntb_batch_free(b);
		return rc;
	}
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventRecord(e0, 0);
	cudaError_t e = cudaMemcpyAsync(b->d_text, bases, b->total, cudaMemcpyHostToDevice, 0);
	cudaEventRecord(e1, 0);
	if (e == cudaSuccess) {
		e = cudaEventSynchronize(e1);
	}
	if (e == cudaSuccess) {
		cudaEventElapsedTime(&b->ms_h2d, e0, e1);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	if (e != cudaSuccess) {
		ntb_batch_free(b);
		return cuda_fail(e, "cudaMemcpy(batch)");
	}
	*out = b;
	return NTB_OK;
}

// ntb_polish_batch's upload: allocation now, the copies as the scan asks for them (ntb_batch::upload_until)
static int
batch_upload_streamed(const char* bases, const uint64_t* offsets, uint64_t n_contigs, int device, ntb_batch** out)
{
	if (!bases || !out) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	int rc = check_offsets(offsets, n_contigs);
	if (rc != NTB_OK) {
		return rc;
	}
	for (uint64_t c = 0; c < n_contigs; c++) {
		if (bases[offsets[c + 1] - 1] != 0) {
			return fail(NTB_EINVAL, "contig " + std::to_string(c) + " is not NUL-terminated inside the batch buffer");
		}
	}
	rc = select_device(device);
	if (rc != NTB_OK) {
		return rc;
	}
	ntb_batch* b = new (std::nothrow) ntb_batch();
	if (!b) {
		return fail(NTB_ENOMEM, "out of memory");
	}
	b->device = device;
	b->offsets.assign(offsets, offsets + n_contigs + 1);
	cudaError_t e = cudaStreamCreateWithFlags(&b->up_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) {
		e = cudaEventCreate(&b->up_begin);
	}
	if (e == cudaSuccess) {
		e = cudaEventCreate(&b->up_end);
	}
	if (e != cudaSuccess) {
		ntb_batch_free(b);
		return cuda_fail(e, "cudaStreamCreate(upload)");
	}
	// the device buffer is bound when the polishing call checks its workspace out (CudaBackend::init)
	b->total = offsets[n_contigs];
	b->n_tiles = (b->total + SCAN_TILE - 1) / SCAN_TILE;
	b->up_src = bases;
	*out = b;
	return NTB_OK;
}

int
ntb_batch_wrap_device(void* dev_bases, const uint64_t* host_offsets, uint64_t n_contigs, int device, ntb_batch** out)
{
	if (!dev_bases || !out) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	int rc = check_offsets(host_offsets, n_contigs);
	if (rc != NTB_OK) {
		return rc;
	}
	rc = select_device(device);
	if (rc != NTB_OK) {
		return rc;
	}
	ntb_batch* b = new (std::nothrow) ntb_batch();
	if (!b) {
		return fail(NTB_ENOMEM, "out of memory");
	}
	b->device = device;
	b->offsets.assign(host_offsets, host_offsets + n_contigs + 1);
	// the kernels want SCAN_HALO readable bytes in front and whole-tile zero padding behind: take a padded device copy
	rc = batch_alloc(b, host_offsets[n_contigs]);
	if (rc != NTB_OK) {
		ntb_batch_free(b);
		return rc;
	}
	cudaError_t e = cudaMemcpy(b->d_text, dev_bases, b->total, cudaMemcpyDeviceToDevice);
	if (e != cudaSuccess) {
		ntb_batch_free(b);
		return cuda_fail(e, "cudaMemcpy(batch, device to device)");
	}
	*out = b;
	return NTB_OK;
}

uint64_t
ntb_batch_total_bases(const ntb_batch* b)
{
	if (!b || b->offsets.empty()) {
		return 0;
	}
	return b->total - (b->offsets.size() - 1);
}

void
ntb_batch_free(ntb_batch* b)
{
	if (!b) {
		return;
	}
	cudaSetDevice(b->device);
	if (b->up_stream) {
		cudaStreamSynchronize(b->up_stream);
	}
	for (cudaEvent_t ev : b->up_events) {
		cudaEventDestroy(ev);
	}
	if (b->up_begin) {
		cudaEventDestroy(b->up_begin);
	}
	if (b->up_end) {
		cudaEventDestroy(b->up_end);
	}
	if (b->up_stream) {
		cudaStreamDestroy(b->up_stream);
	}
	if (b->d_alloc && b->owns_text) {
		cudaFree(b->d_alloc);
	}
	delete b;
}

// ---------------------------------------------------------------- K1 through the ABI
int
ntb_scan(ntb_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_contigs, uint8_t* counts, uint32_t* valid_bits)
{
	if (!f) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	if (f->k < 2 || f->k > KMAX) {
		return fail(NTB_EINVAL, "unsupported k");
	}
	ntb_batch* b = nullptr;
	int rc = ntb_batch_upload(bases, offsets, n_contigs, f->device, &b);
	if (rc != NTB_OK) {
		return rc;
	}
	const uint64_t words = b->n_tiles * SCAN_BITWORDS;
	const uint64_t padded = b->n_tiles * SCAN_TILE;
	uint32_t *d_visit = nullptr, *d_valid = nullptr;
	uint8_t* d_counts = nullptr;
	cudaError_t e = cudaMalloc((void**)&d_visit, words * 4 + 64);
	if (e == cudaSuccess) {
		e = cudaMalloc((void**)&d_valid, words * 4 + 64);
	}
	if (e == cudaSuccess) {
		e = cudaMalloc((void**)&d_counts, padded + 64);
	}
	if (e == cudaSuccess) {
		ScanArgs a;
		std::memset(&a, 0, sizeof a);
		a.text = b->d_text;
		a.n_tiles = b->n_tiles;
		a.filter = f->view();
		a.k = f->k;
		a.min_threshold = 1;
		a.snv = 0;
		a.visit = d_visit;
		a.valid = d_valid;
		a.counts = d_counts;
		fill_scan_tables(a, f->k);
		const int grid = (int)std::min<uint64_t>(b->n_tiles, (uint64_t)sm_count(f->device) * 2);
		if (grid > 0) {
			e = launch_scan(a, f->counting != 0, true, grid, 0);
		}
	}
	if (e == cudaSuccess) {
		e = cudaDeviceSynchronize();
	}
	if (e == cudaSuccess && counts) {
		e = cudaMemcpy(counts, d_counts, b->total, cudaMemcpyDeviceToHost);
	}
	if (e == cudaSuccess && valid_bits) {
		e = cudaMemcpy(valid_bits, d_valid, (b->total + 31) / 32 * 4, cudaMemcpyDeviceToHost);
	}
	cudaFree(d_visit);
	cudaFree(d_valid);
	cudaFree(d_counts);
	ntb_batch_free(b);
	if (e != cudaSuccess) {
		return cuda_fail(e, "scan");
	}
	return NTB_OK;
}

// ---------------------------------------------------------------- polishing
int
ntb_polish_device(ntb_filter* bloom, ntb_filter* rep, const ntb_params* p, ntb_batch* batch, char* host_bases, ntb_result** out)
{
	return polish_common(bloom, rep, p, batch, host_bases, out);
}

int
ntb_polish_batch(ntb_filter* bloom, ntb_filter* rep, const ntb_params* p, char* bases, const uint64_t* offsets, uint64_t n_contigs,
                 ntb_result** out)
{
	if (!bloom) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	ntb_batch* b = nullptr;
	int rc = batch_upload_streamed(bases, offsets, n_contigs, bloom->device, &b);
	if (rc != NTB_OK) {
		return rc;
	}
	rc = polish_common(bloom, rep, p, b, bases, out);
	ntb_batch_free(b);
	return rc;
}

int
ntb_result_contig(const ntb_result* r, uint64_t contig, int* polished, const ntb_node** nodes, uint64_t* n_nodes, const ntb_srec** srecs,
                  uint64_t* n_srecs)
{
	if (!r || contig >= r->impl.contigs.size()) {
		return fail(NTB_EINVAL, "bad contig index");
	}
	const ContigResult& c = r->impl.contigs[contig];
	if (polished) {
		*polished = c.polished ? 1 : 0;
	}
	if (nodes) {
		*nodes = c.nodes.data();
	}
	if (n_nodes) {
		*n_nodes = c.nodes.size();
	}
	if (srecs) {
		*srecs = c.srecs.data();
	}
	if (n_srecs) {
		*n_srecs = c.srecs.size();
	}
	return NTB_OK;
}

int
ntb_result_stats(const ntb_result* r, ntb_stats* st)
{
	if (!r || !st) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	*st = r->impl.stats;
	return NTB_OK;
}

void
ntb_result_free(ntb_result* r)
{
	delete r;
}

// ---------------------------------------------------------------- writer
int
ntb_format_contig(const char* header, const char* seq, const ntb_node* nodes, uint64_t n_nodes, const ntb_srec* srecs, uint64_t n_srecs,
                  int snv, ntb_strbuf* fa, ntb_strbuf* tsv, ntb_strbuf* vcf)
{
	if (!header || !seq || !nodes) {
		return fail(NTB_EINVAL, "NULL argument");
	}
	std::string sfa, stsv, svcf;
	format_contig(header, seq, nodes, (size_t)n_nodes, srecs, (size_t)n_srecs, snv != 0, nullptr, fa ? &sfa : nullptr, tsv ? &stsv : nullptr,
	              vcf ? &svcf : nullptr);
	strbuf_append(fa, sfa);
	strbuf_append(tsv, stsv);
	strbuf_append(vcf, svcf);
	return NTB_OK;
}

int
ntb_format_tsv_header(uint32_t k, uint32_t jump, int counting, ntb_strbuf* tsv)
{
	if (!tsv || jump == 0) {
		return fail(NTB_EINVAL, "bad argument");
	}
	strbuf_append(tsv, tsv_header(k, jump, counting != 0));
	return NTB_OK;
}

void
ntb_strbuf_free(ntb_strbuf* b)
{
	if (b) {
		std::free(b->data);
		b->data = nullptr;
		b->len = b->cap = 0;
	}
}
}
