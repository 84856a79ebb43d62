// Internal types shared by the CUDA kernels, the device engine and the host-side stitcher.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NTB_HD __host__ __device__ __forceinline__
#define NTB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define NTB_HD inline
#define NTB_HD_NOINLINE
#endif

namespace ntb {

constexpr unsigned KMAX = 96;        // largest supported k
constexpr unsigned KMIN = 12;        // k must exceed max_deletions (10) -- see DESIGN.md
constexpr unsigned HMAX = 8;         // largest supported hash_num
constexpr uint32_t NONE32 = 0xFFFFFFFFu;

// Read-only view of a filter resident in device (or, for the test-only host simulator, host) memory.
struct FilterView
{
	const uint8_t* data;
	uint64_t bytes;
	uint64_t mod;    // bit filter: bytes*8 ; counting filter: bytes    (btllib: hash % array_bits / % array_size)
	uint64_t recip;  // floor((2^64-1)/mod), for the multiply-high remainder
	uint64_t mask;   // mod-1 when mod is a power of two, else 0
	uint32_t hash_num;
	uint32_t counting;
};

// Parameters of one polishing run, thresholds pre-reduced to integers on the host with the reference's
// float arithmetic (ntedit.cpp:1531-1535, 1659-1663, 1865-1873, 1892-1897, 1992-1997).
struct KParams
{
	uint32_t k, h;           // from the primary filter
	uint32_t h_rep;          // hash_num of the -e filter (0 = absent)
	uint32_t jump;
	int32_t mode, snv, mask;
	uint32_t max_ins_tries;  // num_tries[max_insertions], ntedit.cpp:172
	uint32_t max_deletions;
	uint32_t thr_missing;    // smallest count c with float(c) >= missing threshold (NONE32 = never)
	uint32_t thr_edit;       // ... edit threshold
	uint32_t thr_edit_del;   // ... tryDeletion's variant, ntedit.cpp:1531-1535
	uint32_t insertion_cap;
	uint32_t min_threshold, max_threshold;
	uint32_t counting;
	uint32_t boundary_lim;   // first-round tasks may move their borders forward by up to this many positions (0 = never), see
	                         // engine.h: safe_boundary()
	// pre-rotated seeds: srol^(k)(S[c]) and srol^(k-1)(S[c]) for the 4 bases (index: A C G T)
	uint64_t seed_rot_k[4];
	uint64_t seed_rot_k1[4];
};

// One unit of device work: walk contig `contig` from tail position `start` (window known to be clean, i.e.
// made of unedited consecutive bases) until tail position >= `end` with a clean window again.
struct Task
{
	uint64_t text_off;   // offset of the contig in the batch buffer
	uint32_t len;        // contig length
	uint32_t start;      // first tail position this task is responsible for
	uint32_t end;        // first tail position of the next task (== len for the last one)
	uint32_t contig;
	uint32_t flags;      // bit0: contig start (run findFirstAcceptedKmer, ntedit.cpp:1773)
	uint32_t pad_;
};
constexpr uint32_t TASK_CONTIG_START = 1u;
constexpr uint32_t TASK_ADJUST_START = 2u; // `start` is a nominal border: begin at safe_boundary(start)
constexpr uint32_t TASK_ADJUST_END = 4u;   // `end` is a nominal border: the successor begins at safe_boundary(end)

// What a walker reports besides its events.
struct TaskResult
{
	uint32_t end_pos;      // tail position at which the walker stopped with a clean window (>= task.end), or len
	uint32_t first_touch;  // tail position of the first site it evaluated (NONE32 if none)
	uint32_t last_event;   // as the walker writes it: index of its last event in the arena (NONE32 if none), events chain
	                       // through `prev`.  As Backend::walk() hands it to the host: index of its FIRST event -- the
	                       // backend lays every walker's n_events events out contiguously, in emission order
	uint32_t n_events;
	uint32_t n_sites;
	uint32_t status;
	uint8_t stale[4];      // values of the reference's uninitialised locals after the walker's last site (see STALE_REF)
	uint32_t kcycles;      // SM clock cycles / 1024 the walker ran for (diagnostics)
	uint32_t start_pos;    // the tail position the walker actually began at (== task.start unless TASK_ADJUST_START moved it)
};
constexpr uint32_t ST_DONE = 1u;          // finished normally
constexpr uint32_t ST_CONTIG_END = 2u;    // the reference's main loop ended (roll failed / guard): nothing after this walker counts
constexpr uint32_t ST_EV_OVERFLOW = 4u;   // event arena exhausted: re-run with a larger arena
constexpr uint32_t ST_ROPE_OVERFLOW = 8u; // local rope capacity exceeded (unsupported input)

// makeEdit-level event (ntedit.cpp:1250-1448) in the order the reference would perform them.
struct Event
{
	uint32_t prev;        // previous event of the same walker
	uint32_t t_pos;       // t_seq_i at the time of the call
	uint32_t advance;     // tail increments (roll calls, ntedit.cpp:2121) since the previous event; NONE32 = anchored:
	                      // the walker's window was clean in between, the tail sits on the final position node at t_pos
	uint16_t support;     // best_num_support
	uint16_t altsupp[3];
	uint8_t kind;         // best_edit_type: 0 none, 1 substitution, 2 insertion, 3 deletion
	uint8_t flags;
	uint8_t draft;        // draft_char
	uint8_t base;         // best_sub_base
	uint8_t altbase[3];
	uint8_t indel_len;    // insertion: number of bases in indel[]; deletion: number of deleted bases
	char indel[5];        // inserted bases
	uint8_t pad_;
};
static_assert(sizeof(Event) == 36, "Event layout");
// The reference declares best_sub_base / altbase1..3 without initialisers inside its loop body (ntedit.cpp:1881-1885);
// the compiled reference keeps them in fixed slots, so a site can report a base that an EARLIER site left behind
// (observable in mode 2 through tryIndels' altsupp1 side channel).  A walker that starts mid-contig does not know what
// its predecessors left, so such bytes travel symbolically: STALE_REF|j means "slot j (0 best_sub, 1..3 altbase1..3) as
// it was when this walker started"; the host resolves them while replaying walkers in contig order.
constexpr uint8_t STALE_REF = 0x80;
constexpr uint8_t EV_TOUCHED = 1;  // a substitution trial patched and reverted the tail char: it is now draft (upper case), ntedit.cpp:1975-1981

// Pre-evaluated sites.  A site the main loop reaches with a CLEAN window (k unedited, consecutive draft bases; the rope a
// single position node) is a pure function of the contig text right of the window and of the filter: its whole
// evaluation (ntedit.cpp:1808-2116 up to, not including, makeEdit) can run ahead of the sequential walk, for all such
// candidate sites at once.  The pre-evaluation kernels (presite_kernel) do that for the first position of every run of
// flagged positions -- and, while a site ends without an edit, for the flagged position right behind it -- and leave one
// record per site in an open-addressing table keyed by the text position; a walker that arrives at a flagged position
// with a clean window looks the record up and only commits it.  A missing record is not an error: the walker then
// evaluates the site itself.
struct SiteRec
{
	uint64_t key;         // text position of the site's tail + 1; 0 = empty slot
	uint16_t support;     // best_num_support
	uint16_t altsupp[3];
	uint8_t state;        // SITE_*
	uint8_t best_type;    // best_edit_type: 0 none, 1 substitution, 2 insertion, 3 deletion
	uint8_t best_sub;     // may be STALE_REF | j, like the alternates (resolved by the walker that commits the record)
	uint8_t altbase[3];
	uint8_t indel_len;
	char indel[5];
	uint8_t flags;        // SITE_FL_*
	uint8_t draft;        // draft_char (upper case)
	uint8_t pad_[2];
};
static_assert(sizeof(SiteRec) == 32, "SiteRec layout");
constexpr uint8_t SITE_NONE = 1;    // no attempt (do_not_fix / too few missing k-mers): nothing is committed
constexpr uint8_t SITE_DONE = 2;    // the decision is in the record
constexpr uint8_t SITE_PENDING = 3; // stopped in front of tryIndels; completed by the second pre-evaluation pass (else ignored)
constexpr uint8_t SITE_FL_TOUCHED = 1; // makeEdit's `touched && raw != draft` (EV_TOUCHED of the event)
constexpr uint8_t SITE_FL_QUIET = 2;   // accepted substitution whose k-1 following windows are no sites: the walker jumps k
constexpr uint8_t SITE_FL_SECOND = 4;  // completed by the second pass (diagnostics)
// A record that made no edit may say that the chain's next sites make none either (filed by the first pass's chain rounds,
// which evaluate DENSE_GROUP consecutive sites of a chain side by side): the walker then goes straight to the last of
// them.  The fields an edit would use carry it: indel_len = sites to jump over (1 .. DENSE_GROUP - 1), indel[4] |
// pad_[0] << 8 = distance to the last of them, indel[0..3] = what those sites leave in the four stale slots, each byte a
// value or STALE_REF | j = "slot j as it is behind THIS record's site".  Only in the plain modes (no -a masking, no -s 1),
// and only over sites that emit nothing (no SITE_FL_TOUCHED).
constexpr uint8_t SITE_FL_SKIP = 8;
constexpr int DENSE_GROUP = 8;         // sites of a chain one round of the first pass evaluates side by side
constexpr uint32_t SITE_TABLE_PROBES = 64;  // linear probing gives up after this many slots (insert: the record is dropped)
constexpr uint32_t SITE_CHAIN_MAX = 64;     // flagged positions one pre-evaluation item follows behind a failed site

// ---- skip information of no-edit records (ntb_common.h: SITE_FL_SKIP)
// may the walker jump over this site?  It makes no edit and emits nothing.
NTB_HD bool
dense_skippable(uint32_t st, uint32_t best_type, uint32_t flags)
{
	return (st == SITE_NONE || (st == SITE_DONE && best_type == 0)) && !(flags & SITE_FL_TOUCHED);
}

// T = what the four stale slots hold behind one more site (state st, bytes b = best_sub | altbase1..3 << 8..24), given T
// before it; both packed one byte per slot, a byte being a value or STALE_REF | j
constexpr uint32_t SKIP_IDENTITY = (uint32_t)(STALE_REF | 0) | ((uint32_t)(STALE_REF | 1) << 8) | ((uint32_t)(STALE_REF | 2) << 16) |
                                   ((uint32_t)(STALE_REF | 3) << 24);

NTB_HD uint32_t
dense_skip_compose(uint32_t T, uint32_t st, uint32_t b)
{
	if (st != SITE_DONE) {
		return T; // no attempt: the slots keep their values
	}
	uint32_t n = 0;
	for (int j = 0; j < 4; j++) {
		uint32_t v = (b >> (8 * j)) & 0xFFu;
		if (v & STALE_REF) {
			v = (T >> (8 * (v & 3u))) & 0xFFu;
		}
		n |= v << (8 * j);
	}
	return n;
}

NTB_HD uint32_t
dense_pack_bases(const SiteRec& r)
{
	return (uint32_t)r.best_sub | ((uint32_t)r.altbase[0] << 8) | ((uint32_t)r.altbase[1] << 16) | ((uint32_t)r.altbase[2] << 24);
}

NTB_HD void
dense_skip_store(SiteRec& r, uint32_t n, uint32_t dist, uint32_t T)
{
	r.flags |= SITE_FL_SKIP;
	r.indel_len = (uint8_t)n;
	for (int j = 0; j < 4; j++) {
		r.indel[j] = (char)((T >> (8 * j)) & 0xFFu);
	}
	r.indel[4] = (char)(dist & 0xFFu);
	r.pad_[0] = (uint8_t)(dist >> 8);
}

// a site whose pre-evaluation stopped in front of tryIndels
struct PendingSite
{
	uint32_t task;  // index of the task (contig segment) the position lies in
	uint32_t pos;   // tail position inside the contig
	uint32_t slot;  // the record's slot in the table
};

// Per-launch device counters.
struct Counters
{
	uint32_t n_events;   // arena fill
	uint32_t overflow;
	uint32_t next_task;  // work queue of the persistent walker warps
	uint32_t n_front;    // order_tasks_kernel: dense tasks placed so far (from the front of the queue)
	uint32_t n_back;     // ... the others (from the back)
	uint32_t n_compact;  // compact_events_kernel: events placed so far
	uint32_t n_items;    // pre-evaluation: items listed so far (heads, then the chain items each round of the first pass adds)
	uint32_t item_lo, item_hi; // ... the items of the current round of the first pass
	uint32_t next_item;  // ... work queue of the first pass (warp form)
	uint32_t n_pending;  // ... sites waiting for the second pass (tryIndels)
	uint32_t next_pending;
	uint32_t n_dropped;  // ... records that found no slot / list entry (the walkers evaluate those sites themselves)
	uint32_t n_rec_used; // walkers: sites committed from a record (diagnostics)
	uint32_t n_rec_used2; // ... from a record of the second pass
	uint32_t n_skipped;  // ... no-edit sites of a chain jumped over (SITE_FL_SKIP); host builds and -DNTB_PHASE_PROF
	uint32_t pad_;
	unsigned long long prof[16]; // -DNTB_PHASE_PROF: leader cycles per phase of the walker
};

} // namespace ntb
