// sm_100a kernels of the ntEdit hot path.
//   K1  scan_kernel       : ntHash roll over every base + h-way filter probe -> visit bitmap (and, on request, counts / validity)
//                           replaces the main-loop test + roll of ntedit.cpp:1806-1807, 2118-2138; filters that fit L2, -s 1, ntb_scan
//   K1b bin_kernel + probe_bin_kernel : the same bitmap for filters larger than L2 -- probes become records bucketed by filter
//                           region, then are served bucket by bucket from the L2-resident region (a direct probe costs a 128-byte
//                           DRAM line on B200)
//   K2p heads_kernel + presite_dense_kernel + presite_kernel : the evaluation of every site the walk can reach with a clean
//                           window, ahead of the walk, into a table of site records (ntb_common.h: SiteRec): first pass one
//                           THREAD per site (site_dense.h: check-missing, gates, substitution trials), second pass one warp
//                           per site that needs tryIndels
//   K3  snv_dense_kernel  : -s 1 -- every valid position is a site: one thread per position evaluates it from the text and
//                           marks the few that do something; the walk jumps through those
//   K2  walk_kernel       : one warp per contig segment replays the edit state machine (engine.h) at the flagged positions:
//                           commits pre-evaluated sites, evaluates the others (dirty windows) one candidate k-mer series per lane
//                           replaces ntedit.cpp:1808-2116 (check-missing, substitutions, tryIndels, tryDeletion, makeEdit decisions)
//       order_tasks_kernel, compact_events_kernel : work-queue order in front of K2, per-walker grouping of its events behind it
//   K4  occupancy_kernel  : popcount / non-zero count of the filter (btllib get_fpr, printed by ntedit.cpp:387-395)
//   K5  insert_kernel     : filter construction (src/ntedit_make_genome_bf.cpp:151-156)
#pragma once
#include "engine.h"
#include "site_dense.h"

#include <cuda_runtime.h>

namespace ntb {

// ------------------------------------------------------------------------------------------------------------------
// K1 geometry: a CTA of SCAN_THREADS threads owns a tile of SCAN_TILE consecutive buffer positions; thread i owns the strip
// [i*SCAN_STRIP, (i+1)*SCAN_STRIP).  SCAN_STRIP is 4 (mod 128) bytes so that the 32 lanes of a warp, each reading the same
// word index of its own strip, hit 32 different shared-memory banks.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_STRIP = 132;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_STRIP;   // 33792 positions = 1056 bitmap words
constexpr int SCAN_HALO = 128;                          // bytes staged in front of a tile (>= KMAX+3, keeps 16-byte alignment)
constexpr int SCAN_STAGE_BYTES = SCAN_HALO + SCAN_TILE; // one staged tile
constexpr int SCAN_STAGES = 2;
constexpr int SCAN_BITWORDS = SCAN_TILE / 32;
// dynamic shared memory of one scan CTA: staged tiles, 1 (or 2 with the optional outputs) tile bitmaps, class table,
// seed tables, mbarriers
inline size_t
scan_smem_bytes(bool extra)
{
	return (size_t)SCAN_STAGES * SCAN_STAGE_BYTES + (size_t)SCAN_BITWORDS * 4 * (extra ? 2 : 1) + 256 + 16 * 8 + SCAN_STAGES * 8;
}

struct ScanArgs
{
	const uint8_t* text;   // buffer position 0; SCAN_HALO readable bytes precede it and the buffer is zero-padded to whole tiles
	uint64_t n_tiles;
	FilterView filter;
	uint32_t k;
	uint32_t min_threshold;
	uint32_t snv;          // visit = valid (no probes)
	uint32_t* visit;       // bit per position: window valid and the main loop would enter its edit block
	uint32_t* valid;       // optional: bit per position: window valid
	uint8_t* counts;       // optional: byte per position: 0/1 (bit filter) or min counter (counting filter); 0 when invalid
	uint64_t mult[HMAX];   // i ^ (k * MULTISEED)
	uint64_t seed[5];      // A C G T none
	uint64_t rotk[5];      // srol^k of the same
};

// launches scan_kernel<hash_num, counting, extra> on `grid` persistent CTAs
cudaError_t launch_scan(const ScanArgs& a, bool counting, bool extra, int grid, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------------------------
// K1b: binned scan.  B200's L2 fetches a whole 128-byte line from DRAM for every missed sector, so a random 1-bit probe
// into a multi-GB filter costs 128 bytes of HBM traffic (profiles/r01_*).  The binned scan therefore does not load the
// filter while it hashes: bin_kernel turns every probe into an 8-byte record
//     (position within the text chunk : 32 | slot within the filter region : 32)
// appended to the bucket of the filter REGION (2^region_log2 slots, sized to sit in L2) the probe falls into; records are
// counting-sorted per round in shared memory and leave the SM as contiguous runs.  probe_bin_kernel then walks the buckets
// one after the other, every CTA on the same bucket, so that the probes of a bucket hit the one region that is L2-resident;
// a failed probe sets the visit bit of its position.  Same visit bitmap as scan_kernel, bit for bit.
constexpr int BIN_POS_PER_ROUND = 4;                             // positions per thread and round (same step as K1)
constexpr int BIN_MAX_BUCKETS = SCAN_THREADS;                    // one thread per bucket in the per-round scan
constexpr int BIN_ROUND_RECORDS = SCAN_THREADS * BIN_POS_PER_ROUND * (int)HMAX; // upper bound for the template maximum
#ifndef NTB_BIN_STAGES
#define NTB_BIN_STAGES 1
#endif
constexpr int BIN_STAGES = NTB_BIN_STAGES;                       // staged text tiles per CTA: 1 leaves room for a third CTA per SM
constexpr int BIN_CTAS_PER_SM = BIN_STAGES == 1 ? 3 : 2;
inline size_t
bin_smem_bytes(int hash_num)
{
	const size_t round_records = (size_t)SCAN_THREADS * BIN_POS_PER_ROUND * (size_t)hash_num;
	return (size_t)BIN_STAGES * SCAN_STAGE_BYTES + round_records * 8 + BIN_MAX_BUCKETS * 8 + 2 * (BIN_MAX_BUCKETS + 32) * 4 + 16 * 4 + 256 + 16 * 8 +
	       SCAN_STAGES * 8 + 64;
}

struct BinArgs
{
	ScanArgs scan;          // text, filter, tables, visit (n_tiles = tiles of this chunk, text = chunk start)
	uint64_t chunk_base;    // text position of the chunk's first tile (multiple of SCAN_TILE)
	uint64_t* records;      // n_buckets x bucket_cap
	uint32_t* cursor;       // records appended per bucket (may exceed bucket_cap: the excess was probed directly)
	uint32_t bucket_cap;
	uint32_t n_buckets;
	uint32_t region_log2;   // slots per region = 1 << region_log2
	uint32_t pace_lag;      // probe kernel: a CTA starts bucket b when every CTA has finished bucket b - pace_lag (1 or 2)
};

// bin_kernel<hash_num, counting> on `grid` persistent CTAs
cudaError_t launch_bin(const BinArgs& a, bool counting, int grid, cudaStream_t stream);
// probe_bin_kernel<counting>, launched cooperatively on min(ctas_per_sm, occupancy) CTAs per SM; `cursor` has
// BIN_MAX_BUCKETS + 1 entries zeroed before the bin kernel (the last one paces the probe CTAs)
cudaError_t launch_probe_bin(const BinArgs& a, bool counting, int ctas_per_sm, cudaStream_t stream);

// K2 geometry: WALK_TEAMS walkers per CTA, each with its own WalkerState in dynamic shared memory
#ifndef NTB_WALK_WARPS
#define NTB_WALK_WARPS 4
#endif
#ifndef NTB_WALK_MIN_CTAS
#define NTB_WALK_MIN_CTAS 5   // CTAs per SM the walker's shared-memory footprint allows: keep the register count below that bound too
#endif
constexpr int WALK_WARPS = NTB_WALK_WARPS;
constexpr int WALK_THREADS = WALK_WARPS * 32;
constexpr int WALK_TEAMS = WALK_THREADS / NTB_TEAM;   // walkers per CTA (engine.h: a team of NTB_TEAM lanes runs one walker)

// arguments of the walker-layout kernels (K2 and the pre-evaluation passes in front of it)
struct WalkArgs
{
	const uint8_t* text;
	const uint32_t* visit;
	FilterView bloom, rep;
	KParams kp;
	const Task* tasks;
	uint32_t* order;          // n_tasks entries, may be NULL = queue order
	TaskResult* results;
	uint32_t n_tasks;
	Event* events;
	uint32_t ev_cap;
	Counters* ctr;
	int sm_count;
	// pre-evaluated sites (ntb_common.h: SiteRec); table == NULL: none
	SiteRec* table;
	uint32_t table_mask;
	uint2* items;             // heads listed by launch_heads: (task, tail position)
	uint32_t items_cap;
	PendingSite* pending;
	uint32_t pending_cap;
};

// orders the tasks (dense ones first) into `order` and launches the walker kernel on a persistent grid sized from its
// occupancy; 2 launches
cudaError_t launch_walk(const WalkArgs& a, cudaStream_t stream);

// pre-evaluation in front of the first walker round: launch_heads lists the heads of every task's nominal range
// (ctr->n_items), launch_presite(second = false) evaluates them and the chains behind them into the table, leaving the
// sites that reach tryIndels in `pending` (ctr->n_pending), launch_presite(second = true) completes those.  The counters
// must be zero before launch_heads.
cudaError_t launch_heads(const WalkArgs& a, cudaStream_t stream);
cudaError_t launch_presite(const WalkArgs& a, bool second, cudaStream_t stream);
uint32_t presite_launch_count(); // kernels launch_heads + both passes launch
// K3 (-s 1): evaluates every valid position of the tasks' nominal ranges, files the records of the sites that do something
// and marks those in `visit2` (zeroed by the caller, same geometry as a.visit); the walkers then take visit2 as their bitmap
cudaError_t launch_snv_dense(const WalkArgs& a, uint32_t* visit2, cudaStream_t stream);

// lays every walker's events out contiguously in `out`, in emission order; results[i].last_event becomes the index of the first
// dst[0, bytes) = pinned host memory src_host[0, bytes), read by a kernel (no copy engine); whole 16-byte units
cudaError_t launch_fetch_host(void* dst, const void* src_host, size_t bytes, cudaStream_t stream);
cudaError_t launch_compact_events(const Event* in, Event* out, TaskResult* results, uint32_t n_tasks, Counters* ctr, cudaStream_t stream);

__global__ void insert_kernel(const uint8_t* text, uint64_t total, uint8_t* data, FilterView f, const __grid_constant__ KParams kp);

__global__ void occupancy_kernel(const uint8_t* data, uint64_t bytes, int counting, unsigned long long* out);

} // namespace ntb
