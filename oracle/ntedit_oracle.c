/* ORACLE / TEST INFRASTRUCTURE ONLY -- see ntedit_oracle.h.  Plain-C CPU restatement of ntEdit's
 * hot path; every function cites the reference file:line (under /root/reference) it follows.
 * Never linked into or called from the product path (ntedit_b200/). */
#define _GNU_SOURCE
#include "ntedit_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* =====================================================================================
 * ntHash2 as used through btllib::hashing_internals (call sites
 * ntedit.cpp:412-415, 428-431, 444-451)
 * ===================================================================================== */
#define ORC_MULTISEED 0x90b45d39fb6da1faULL
#define ORC_MULTISHIFT 27
#define ORC_CP_OFF 0x07

static const uint64_t ORC_SEED_A = 0x3c8bfbb395c60474ULL;
static const uint64_t ORC_SEED_C = 0x3193c18562a02b4cULL;
static const uint64_t ORC_SEED_G = 0x20323ed082572324ULL;
static const uint64_t ORC_SEED_T = 0x295549f54be24456ULL;

uint64_t
orc_seed(unsigned char c)
{
	switch (c) {
	case 'A': case 'a': case 4: case 5: return ORC_SEED_A; /* slots 4,5 = complement of 'T','U' by low 3 bits */
	case 'C': case 'c': case 7: return ORC_SEED_C;         /* slot 7 = complement of 'G' */
	case 'G': case 'g': case 3: return ORC_SEED_G;         /* slot 3 = complement of 'C' */
	case 'T': case 't': case 'U': case 'u': case 1: return ORC_SEED_T; /* slot 1 = complement of 'A' */
	default: return 0;
	}
}

/* split rotate: low 33 bits and high 31 bits rotate independently by one */
uint64_t
orc_srol(uint64_t x)
{
	uint64_t carry = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
	return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | carry;
}

uint64_t
orc_sror(uint64_t x)
{
	uint64_t carry = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
	return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | carry;
}

uint64_t
orc_srol_n(uint64_t x, unsigned d)
{
	while (d--) {
		x = orc_srol(x);
	}
	return x;
}

uint64_t
orc_base_forward_hash(const char* s, unsigned k)
{
	uint64_t h = 0;
	for (unsigned i = 0; i < k; i++) {
		h = orc_srol(h) ^ orc_seed((unsigned char)s[i]);
	}
	return h;
}

uint64_t
orc_base_reverse_hash(const char* s, unsigned k)
{
	uint64_t h = 0;
	for (unsigned i = 0; i < k; i++) {
		h = orc_srol(h) ^ orc_seed((unsigned char)s[k - 1 - i] & ORC_CP_OFF);
	}
	return h;
}

uint64_t
orc_next_forward_hash(uint64_t fh, unsigned k, unsigned char out, unsigned char in)
{
	return orc_srol(fh) ^ orc_seed(in) ^ orc_srol_n(orc_seed(out), k);
}

uint64_t
orc_next_reverse_hash(uint64_t rh, unsigned k, unsigned char out, unsigned char in)
{
	return orc_sror(rh ^ orc_srol_n(orc_seed(in & ORC_CP_OFF), k) ^ orc_seed(out & ORC_CP_OFF));
}

void
orc_extend_hashes(uint64_t base, unsigned k, unsigned h, uint64_t* out)
{
	out[0] = base;
	for (unsigned i = 1; i < h; i++) {
		uint64_t t = base * ((uint64_t)i ^ ((uint64_t)k * ORC_MULTISEED));
		t ^= t >> ORC_MULTISHIFT;
		out[i] = t;
	}
}

/* ntedit.cpp:403-416 */
void
orc_ntmc64_seed(const char* s, unsigned k, unsigned h, uint64_t* fh, uint64_t* rh, uint64_t* hv)
{
	*fh = orc_base_forward_hash(s, k);
	*rh = orc_base_reverse_hash(s, k);
	orc_extend_hashes(*fh + *rh, k, h, hv);
}

/* ntedit.cpp:418-432 */
void
orc_ntmc64_roll(unsigned char out, unsigned char in, unsigned k, unsigned h, uint64_t* fh, uint64_t* rh, uint64_t* hv)
{
	*fh = orc_next_forward_hash(*fh, k, out, in);
	*rh = orc_next_reverse_hash(*rh, k, out, in);
	orc_extend_hashes(*fh + *rh, k, h, hv);
}

/* ntedit.cpp:434-452 */
void
orc_ntmc64_changelast(unsigned char out, unsigned char in, unsigned k, unsigned h, uint64_t* fh, uint64_t* rh, uint64_t* hv)
{
	*fh ^= orc_seed(out) ^ orc_seed(in);
	*rh ^= orc_srol_n(orc_seed(out & ORC_CP_OFF), k - 1) ^ orc_srol_n(orc_seed(in & ORC_CP_OFF), k - 1);
	orc_extend_hashes(*fh + *rh, k, h, hv);
}

/* =====================================================================================
 * Filters: btllib KmerBloomFilter / KmerCountingBloomFilter8 (SURVEY.md Appendix B;
 * used through BFWrapper, ntedit.cpp:350-401)
 * ===================================================================================== */
orc_filter*
orc_filter_new(uint64_t bytes, unsigned k, unsigned h, int counting)
{
	orc_filter* f = (orc_filter*)calloc(1, sizeof(*f));
	if (!f) {
		return NULL;
	}
	f->data = (uint8_t*)calloc(bytes ? bytes : 1, 1);
	if (!f->data) {
		free(f);
		return NULL;
	}
	f->bytes = bytes;
	f->k = k;
	f->h = h;
	f->counting = counting;
	return f;
}

void
orc_filter_free(orc_filter* f)
{
	if (f) {
		free(f->data);
		free(f);
	}
}

unsigned
orc_filter_count(const orc_filter* f, const uint64_t* hv)
{
	if (!f->counting) {
		return 1; /* BFWrapper::get_count, ntedit.cpp:373-376 */
	}
	unsigned m = 255;
	for (unsigned i = 0; i < f->h; i++) {
		unsigned c = f->data[hv[i] % f->bytes];
		if (c < m) {
			m = c;
		}
	}
	return m;
}

int
orc_filter_contains(const orc_filter* f, const uint64_t* hv)
{
	if (f->counting) {
		return orc_filter_count(f, hv) > 0; /* ntedit.cpp:368-371 */
	}
	const uint64_t bits = f->bytes * 8;
	for (unsigned i = 0; i < f->h; i++) {
		uint64_t n = hv[i] % bits;
		if (!(f->data[n >> 3] & (1u << (n & 7)))) {
			return 0;
		}
	}
	return 1;
}

void
orc_filter_insert_hashes(orc_filter* f, const uint64_t* hv)
{
	if (f->counting) {
		/* our own builder semantics (not on the ntEdit path; ntStat builds real .cbf files): every one of
		 * the h counters is incremented, saturating at 255 -- order independent, so a parallel GPU build
		 * gives the identical array. */
		for (unsigned i = 0; i < f->h; i++) {
			uint8_t* c = &f->data[hv[i] % f->bytes];
			if (*c != 255) {
				(*c)++;
			}
		}
		return;
	}
	const uint64_t bits = f->bytes * 8;
	for (unsigned i = 0; i < f->h; i++) {
		uint64_t n = hv[i] % bits;
		f->data[n >> 3] |= (uint8_t)(1u << (n & 7));
	}
}

static int
orc_is_acgt_any_case(unsigned char c)
{
	c = (unsigned char)toupper(c);
	return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

void
orc_filter_insert_seq(orc_filter* f, const char* seq, size_t len)
{
	const unsigned k = f->k;
	uint64_t hv[64];
	if (len < k || f->h > 64) {
		return;
	}
	size_t run = 0; /* number of consecutive ACGT bases ending at i */
	uint64_t fh = 0, rh = 0;
	for (size_t i = 0; i < len; i++) {
		if (!orc_is_acgt_any_case((unsigned char)seq[i])) {
			run = 0;
			continue;
		}
		run++;
		if (run < k) {
			continue;
		}
		if (run == k) {
			orc_ntmc64_seed(seq + i + 1 - k, k, f->h, &fh, &rh, hv);
		} else {
			orc_ntmc64_roll((unsigned char)seq[i - k], (unsigned char)seq[i], k, f->h, &fh, &rh, hv);
		}
		orc_filter_insert_hashes(f, hv);
	}
}

int
orc_filter_save(const orc_filter* f, const char* path)
{
	FILE* fp = fopen(path, "wb");
	if (!fp) {
		return -1;
	}
	if (f->counting) {
		fprintf(fp, "[BTLKmerCountingBloomFilter_v5]\nbytes = %llu\ncounter_bits = 8\nhash_fn = \"ntHash_v2\"\nhash_num = %u\nk = %u\n[HeaderEnd]\n",
		        (unsigned long long)f->bytes, f->h, f->k);
	} else {
		fprintf(fp, "[BTLKmerBloomFilter_v7]\nbytes = %llu\nhash_fn = \"ntHash_v2\"\nhash_num = %u\nk = %u\n[HeaderEnd]\n",
		        (unsigned long long)f->bytes, f->h, f->k);
	}
	size_t w = fwrite(f->data, 1, f->bytes, fp);
	fclose(fp);
	return w == f->bytes ? 0 : -1;
}

orc_filter*
orc_filter_load(const char* path)
{
	FILE* fp = fopen(path, "rb");
	if (!fp) {
		return NULL;
	}
	char line[512];
	uint64_t bytes = 0;
	unsigned k = 0, h = 0;
	int counting = 0, first = 1, ok = 0;
	while (fgets(line, sizeof line, fp)) {
		if (first) {
			counting = strstr(line, "Counting") != NULL;
			first = 0;
			continue;
		}
		if (!strncmp(line, "[HeaderEnd]", 11)) {
			ok = 1;
			break;
		}
		unsigned long long v;
		if (sscanf(line, " bytes = %llu", &v) == 1) {
			bytes = v;
		} else if (sscanf(line, " hash_num = %llu", &v) == 1) {
			h = (unsigned)v;
		} else if (sscanf(line, " k = %llu", &v) == 1) {
			k = (unsigned)v;
		}
	}
	if (!ok) {
		fclose(fp);
		return NULL;
	}
	orc_filter* f = orc_filter_new(bytes, k, h, counting);
	if (f && fread(f->data, 1, bytes, fp) != bytes) {
		orc_filter_free(f);
		f = NULL;
	}
	fclose(fp);
	return f;
}

double
orc_filter_fpr(const orc_filter* f)
{
	uint64_t n = 0;
	for (uint64_t i = 0; i < f->bytes; i++) {
		n += f->counting ? (f->data[i] != 0) : (uint64_t)__builtin_popcount(f->data[i]);
	}
	double occ = (double)n / (double)(f->counting ? f->bytes : f->bytes * 8);
	return pow(occ, (double)f->h);
}

/* ntedit.cpp:486-499 */
static int
is_atgc(unsigned char c)
{
	return c == 'A' || c == 'T' || c == 'G' || c == 'C';
}

static int
is_accepted(unsigned char c)
{
	switch (c) {
	case 'A': case 'T': case 'G': case 'C': case 'R': case 'Y': case 'S':
	case 'W': case 'K': case 'M': case 'B': case 'D': case 'H': case 'V':
		return 1;
	default:
		return 0;
	}
}

void
orc_scan_counts(const orc_filter* f, const char* seq, size_t len, uint8_t* out)
{
	const unsigned k = f->k;
	uint64_t hv[64];
	size_t run = 0;
	for (size_t t = 0; t < len; t++) {
		run = is_accepted((unsigned char)toupper((unsigned char)seq[t])) ? run + 1 : 0;
		if (run < k) {
			out[t] = 0xFF;
			continue;
		}
		uint64_t fh, rh;
		orc_ntmc64_seed(seq + t + 1 - k, k, f->h, &fh, &rh, hv);
		out[t] = f->counting ? (uint8_t)orc_filter_count(f, hv) : (uint8_t)orc_filter_contains(f, hv);
	}
}

/* =====================================================================================
 * Engine restatement (ntedit.cpp:524-2151)
 * ===================================================================================== */
void
orc_params_default(orc_params* p, unsigned k, unsigned h)
{
	memset(p, 0, sizeof *p);
	p->k = k;
	p->h = h;
	p->jump = 3;
	p->mode = 0;
	p->max_insertions = 5;
	p->max_deletions = 5;
	p->edit_threshold = 9.0f;
	p->missing_threshold = 5.0f;
	p->edit_ratio = 0.5f;
	p->missing_ratio = 0.5f;
	p->insertion_cap = (unsigned)((float)k * 1.5f);
	p->min_threshold = 1;
	p->max_threshold = 255;
}

typedef struct nodevec {
	orc_node* v;
	size_t n, cap;
} nodevec;

static void
nv_reserve(nodevec* nv, size_t want)
{
	if (want <= nv->cap) {
		return;
	}
	size_t c = nv->cap ? nv->cap * 2 : 64;
	while (c < want) {
		c *= 2;
	}
	nv->v = (orc_node*)realloc(nv->v, c * sizeof(orc_node));
	nv->cap = c;
}

static void
nv_push(nodevec* nv, orc_node x)
{
	nv_reserve(nv, nv->n + 1);
	nv->v[nv->n++] = x;
}

/* store x at index i, appending when i is one past the end (the reference's "if (i < size) v[i]=x else push_back") */
static void
nv_put(nodevec* nv, size_t i, orc_node x)
{
	if (i < nv->n) {
		nv->v[i] = x;
	} else {
		nv_push(nv, x);
	}
}

typedef struct srecvec {
	orc_srec* v;
	size_t n, cap;
} srecvec;

static void
sv_push(srecvec* sv, orc_srec x)
{
	if (sv->n == sv->cap) {
		sv->cap = sv->cap ? sv->cap * 2 : 64;
		sv->v = (orc_srec*)realloc(sv->v, sv->cap * sizeof(orc_srec));
	}
	sv->v[sv->n++] = x;
}

typedef struct cursor {
	uint32_t pos; /* h_seq_i / t_seq_i */
	uint32_t ni;  /* h_node_index / t_node_index */
} cursor;

typedef struct engine {
	char* seq;
	uint32_t len;
	const orc_filter* bloom;
	const orc_filter* rep;
	const orc_params* p;
	nodevec rope;
	srecvec recs;
	uint64_t hv[64];
} engine;

/* ntedit.cpp:812-823 */
static unsigned char
node_char(const engine* e, uint32_t pos, const orc_node* nd)
{
	if (nd->node_type == 0) {
		return pos < e->len ? (unsigned char)e->seq[pos] : 0;
	}
	if (nd->node_type == 1) {
		return nd->c;
	}
	return 0;
}

static unsigned char
cursor_char(const engine* e, const cursor* c)
{
	static const orc_node dead = { -1, 0, 0, 0, 0 };
	return node_char(e, c->pos, c->ni < e->rope.n ? &e->rope.v[c->ni] : &dead);
}

/* ntedit.cpp:826-844 */
static void
step(const engine* e, cursor* c)
{
	const orc_node nd = e->rope.v[c->ni];
	if (nd.node_type == 0) {
		c->pos++;
		if (c->pos > nd.e_pos) {
			c->ni++;
			if (c->ni < e->rope.n && e->rope.v[c->ni].node_type == 0) {
				c->pos = e->rope.v[c->ni].s_pos;
			}
		}
	} else if (nd.node_type == 1) {
		c->ni++;
		if (c->ni < e->rope.n && e->rope.v[c->ni].node_type == 0) {
			c->pos = e->rope.v[c->ni].s_pos;
		}
	}
}

/* ntedit.cpp:1216-1247 */
static int
roll_cursors(const engine* e, cursor* h, cursor* t, unsigned char* out, unsigned char* in)
{
	if (h->pos >= e->len || h->ni >= e->rope.n) {
		return 0;
	}
	*out = cursor_char(e, h);
	step(e, h);
	if (t->pos >= e->len || t->ni >= e->rope.n) {
		return 0;
	}
	step(e, t);
	if (t->pos >= e->len || t->ni >= e->rope.n) {
		return 0;
	}
	*in = cursor_char(e, t);
	return 1;
}

/* ntedit.cpp:524-545 */
static uint32_t
first_accepted_kmer(const engine* e, uint32_t from)
{
	const uint32_t k = e->p->k;
	for (uint32_t i = from; (uint64_t)i + k < e->len;) {
		if (is_accepted((unsigned char)toupper((unsigned char)e->seq[i]))) {
			int good = 1;
			for (uint32_t j = i + 1; j < i + k; j++) {
				if (!is_accepted((unsigned char)toupper((unsigned char)e->seq[j]))) {
					good = 0;
					i = j + 1;
					break;
				}
			}
			if (good) {
				return i;
			}
		} else {
			i++;
		}
	}
	return e->len - 1;
}

/* ntedit.cpp:501-520 */
static char
revcomp_base(unsigned char c)
{
	switch (c) {
	case 'A': case 'a': return 'T';
	case 'T': case 't': return 'A';
	case 'G': case 'g': return 'C';
	case 'C': case 'c': return 'G';
	default: return 'N';
	}
}

/* ntedit.cpp:561-596: true when s is a whole number of repeats of a shorter word (KMP failure function) */
static int
is_repeat(const char* s, int n)
{
	if (n <= 0) {
		return 0;
	}
	int* lps = (int*)malloc(sizeof(int) * (size_t)n);
	int l = 0, i = 1;
	lps[0] = 0;
	while (i < n) {
		if (s[i] == s[l]) {
			lps[i++] = ++l;
		} else if (l != 0) {
			l = lps[l - 1];
		} else {
			lps[i++] = 0;
		}
	}
	int last = lps[n - 1];
	free(lps);
	return last > 0 && n % (n - last) == 0;
}

/* ntedit.cpp:465-473 */
static int
kmer_solid(const engine* e)
{
	int ok_rep = !e->p->secbf || !orc_filter_contains(e->rep, e->hv);
	int ok_cnt = !e->bloom->counting || (orc_filter_count(e->bloom, e->hv) <= e->p->max_threshold &&
	                                     orc_filter_count(e->bloom, e->hv) >= e->p->min_threshold);
	return ok_rep && ok_cnt;
}

static int
present_solid(const engine* e)
{
	return orc_filter_contains(e->bloom, e->hv) && kmer_solid(e);
}

/* ntedit.cpp:455-463 */
static int
cmp_u8(const void* a, const void* b)
{
	return (int)*(const uint8_t*)a - (int)*(const uint8_t*)b;
}

static unsigned
median_u8(uint8_t* v, size_t n)
{
	if (n == 0) {
		return 0;
	}
	qsort(v, n, 1, cmp_u8);
	return v[n / 2];
}

/* ntedit.cpp:625-714 */
static void
rope_insert(engine* e, uint32_t* t_ni, uint32_t insert_pos, const char* bases, size_t nb, unsigned support)
{
	nodevec* r = &e->rope;
	const orc_node orig = r->v[*t_ni];
	orc_node ins[8];
	for (size_t i = 0; i < nb; i++) {
		ins[i].node_type = 1;
		ins[i].s_pos = ins[i].e_pos = 0;
		ins[i].c = (uint8_t)bases[i];
		ins[i].num_support = support;
	}
	if (orig.node_type == 0 && insert_pos > orig.s_pos) {
		/* split the position node around the insertion */
		orc_node after = { 0, insert_pos, orig.e_pos, 0, 0 };
		r->v[*t_ni].e_pos = insert_pos - 1;
		for (size_t i = 0; i < nb; i++) {
			nv_put(r, *t_ni + i + 1, ins[i]);
		}
		nv_put(r, *t_ni + nb + 1, after);
		(*t_ni)++;
		return;
	}
	if (orig.node_type == 0 || orig.node_type == 1) {
		/* insert in front of the tail node: lift the live run that starts there, then put it back */
		size_t i = *t_ni, nlift = 0;
		while (i < r->n && r->v[i].node_type != -1) {
			nlift++;
			i++;
		}
		orc_node* lift = (orc_node*)malloc(sizeof(orc_node) * (nlift ? nlift : 1));
		for (size_t q = 0; q < nlift; q++) {
			lift[q] = r->v[*t_ni + q];
			r->v[*t_ni + q].node_type = -1;
		}
		for (size_t q = 0; q < nb; q++) {
			nv_put(r, *t_ni + q, ins[q]);
		}
		for (size_t q = 0; q < nlift; q++) {
			nv_put(r, *t_ni + nb + q, lift[q]);
		}
		free(lift);
	}
}

/* ntedit.cpp:719-809 */
static void
rope_delete(engine* e, uint32_t* t_ni, uint32_t* pos, unsigned num_del, unsigned support)
{
	nodevec* r = &e->rope;
	const orc_node orig = r->v[*t_ni];
	if (orig.node_type == 0) {
		unsigned leftover = 0;
		if (*pos <= orig.s_pos) {
			if (*pos + num_del <= orig.e_pos) {
				r->v[*t_ni].s_pos = *pos + num_del;
				r->v[*t_ni].num_support = support;
				*pos = r->v[*t_ni].s_pos;
				return;
			}
			leftover = *pos + num_del - orig.e_pos;
			*pos = orig.e_pos + 1;
			size_t i = (size_t)*t_ni + 1;
			while (i < r->n && r->v[i].node_type != -1) {
				r->v[i - 1] = r->v[i];
				r->v[i].node_type = -1;
				i++;
			}
		} else {
			if (*pos + num_del <= orig.e_pos) {
				orc_node split = { 0, *pos + num_del, orig.e_pos, support, 0 };
				r->v[*t_ni].e_pos = *pos - 1;
				*pos = split.s_pos;
				(*t_ni)++;
				nv_put(r, *t_ni, split);
				return;
			}
			leftover = *pos + num_del - orig.e_pos;
			r->v[*t_ni].e_pos = *pos - 1;
			*pos = orig.e_pos + 1;
			(*t_ni)++;
		}
		if (leftover > 0 && *t_ni < r->n && r->v[*t_ni].node_type != -1) {
			if (r->v[*t_ni].node_type == 0) {
				*pos = r->v[*t_ni].s_pos;
			}
			rope_delete(e, t_ni, pos, leftover, support);
		}
	} else if (orig.node_type == 1) {
		size_t i = *t_ni;
		unsigned leftover = num_del;
		while (i < r->n && r->v[i].node_type == 1 && leftover > 0) {
			r->v[i].node_type = -1;
			leftover--;
			i++;
		}
		size_t j = *t_ni;
		while (i < r->n && r->v[i].node_type != -1) {
			r->v[j] = r->v[i];
			r->v[i].node_type = -1;
			i++;
			j++;
		}
		if (leftover > 0 && *t_ni < r->n && r->v[*t_ni].node_type != -1) {
			if (r->v[*t_ni].node_type == 0) {
				*pos = r->v[*t_ni].s_pos;
			}
			rope_delete(e, t_ni, pos, leftover, support);
		}
	}
}

/* ntedit.cpp:848-903.  Writes the k-mer into kmer (k+1 bytes); returns its length (k on success, 0 on failure) */
static unsigned
find_accepted_kmer(engine* e, cursor* h, cursor* t, char* kmer)
{
	nodevec* r = &e->rope;
	const unsigned k = e->p->k;
	orc_node cur = r->v[t->ni];
	uint32_t tni = t->ni, hni = 0;
	uint32_t i = t->pos;
	while (i < e->len && tni < r->n && r->v[tni].node_type != -1) {
		unsigned char c = node_char(e, i, &cur);
		if (is_accepted((unsigned char)toupper(c))) {
			unsigned n = 0;
			kmer[n++] = (char)c;
			hni = tni;
			cursor jc = { i, tni };
			step(e, &jc);
			tni = jc.ni;
			uint32_t j = jc.pos;
			while (j < e->len && tni < r->n && r->v[tni].node_type != -1) {
				cur = r->v[tni];
				c = node_char(e, j, &cur);
				if (!is_accepted((unsigned char)toupper(c))) {
					i = j;
					break;
				}
				kmer[n++] = (char)c;
				if (n == k) {
					break;
				}
				cursor jj = { j, tni };
				step(e, &jj);
				j = jj.pos;
				tni = jj.ni;
			}
			if (n == k) {
				h->pos = i;
				t->pos = j;
				h->ni = hni;
				t->ni = tni;
				kmer[n] = 0;
				return n;
			}
		}
		if (tni < r->n) {
			cursor ii = { i, tni };
			step(e, &ii);
			i = ii.pos;
			tni = ii.ni;
		}
	}
	h->pos = e->len;
	t->pos = e->len;
	kmer[0] = 0;
	return 0;
}

/* ntedit.cpp:907-922.  Returns the length written into out (reverse-complemented run of inserted chars) */
static size_t
prev_insertion(const engine* e, uint32_t t_pos, uint32_t t_ni, char* out, size_t cap)
{
	const nodevec* r = &e->rope;
	size_t n = 0;
	if ((t_ni < r->n && r->v[t_ni].node_type == 0 && t_pos == r->v[t_ni].s_pos) || r->v[t_ni].node_type == 1) {
		t_ni--;
	}
	while (t_ni < r->n && r->v[t_ni].node_type == 1 && n + 1 < cap) {
		out[n++] = revcomp_base(r->v[t_ni].c);
		t_ni--;
	}
	out[n] = 0;
	return n;
}

typedef struct site {
	unsigned best_type; /* 0 none, 1 substitution, 2 insertion, 3 deletion */
	unsigned char best_sub;
	char best_indel[16];
	size_t best_indel_len;
	unsigned best_support;
	unsigned char altbase1, altbase2, altbase3;
	unsigned altsupp1, altsupp2, altsupp3;
} site;

/* the removal loop shared by both guard branches of makeEdit (ntedit.cpp:1321-1334 / 1352-1366) */
static void
guard_remove(engine* e, const cursor* t, size_t count)
{
	nodevec* r = &e->rope;
	unsigned j = 1;
	if (r->v[t->ni].node_type == 0 && t->pos == r->v[t->ni].s_pos) {
		j = 0;
	}
	for (size_t i = count; i > 0; i--) {
		if ((size_t)t->ni + j < r->n && r->v[t->ni + j].node_type != -1) {
			r->v[t->ni - i] = r->v[t->ni + j];
			r->v[t->ni + j].node_type = -1;
			j++;
		} else {
			r->v[t->ni - i].node_type = -1;
		}
	}
}

/* ntedit.cpp:1250-1448 */
static void
apply_edit(engine* e, unsigned char draft, site* s, cursor* h, cursor* t, uint64_t* fh, uint64_t* rh)
{
	const orc_params* p = e->p;
	const orc_node tnode = e->rope.v[t->ni];
	switch (s->best_type) {
	case 1:
		if (tnode.node_type == 0) {
			e->seq[t->pos] = (char)s->best_sub;
			orc_srec r;
			memset(&r, 0, sizeof r);
			r.draft_char = draft;
			r.pos = t->pos;
			r.sub_base = s->best_sub;
			r.num_support = s->best_support;
			if (s->altsupp1 && s->altbase1 != s->best_sub) {
				r.altbase1 = s->altbase1;
				r.altsupp1 = s->altsupp1;
			}
			if (s->altsupp2 && s->altbase2 != s->altbase1) {
				r.altbase2 = s->altbase2;
				r.altsupp2 = s->altsupp2;
			}
			if (s->altsupp3 && s->altbase3 != s->altbase2) {
				r.altbase3 = s->altbase3;
				r.altsupp3 = s->altsupp3;
			}
			sv_push(&e->recs, r);
		} else if (tnode.node_type == 1) {
			e->rope.v[t->ni].c = s->best_sub;
		}
		orc_ntmc64_changelast(draft, s->best_sub, p->k, p->h, fh, rh, e->hv);
		break;
	case 2: {
		int skipped = 0;
		char prev[1024];
		size_t np = prev_insertion(e, t->pos, t->ni, prev, sizeof prev - 16);
		char kmer[512];
		if (np + s->best_indel_len >= p->k) {
			if (is_repeat(prev, (int)np) || np + s->best_indel_len >= p->insertion_cap) {
				guard_remove(e, t, np);
				unsigned n = find_accepted_kmer(e, h, t, kmer);
				(void)n;
				/* the reference re-seeds from the (possibly empty) k-mer; with an empty k-mer it reads past
				 * the string.  The hash is never consulted again when the search failed (the next roll
				 * stops the contig), so only the successful case is restated. */
				if (n == p->k) {
					orc_ntmc64_seed(kmer, p->k, p->h, fh, rh, e->hv);
				}
				skipped = 1;
			} else {
				for (size_t w = 0; w < s->best_indel_len; w++) {
					memmove(prev + 1, prev, np + 1);
					prev[0] = revcomp_base((unsigned char)s->best_indel[w]);
					np++;
					if (is_repeat(prev, (int)np)) {
						guard_remove(e, t, np - w);
						unsigned n = find_accepted_kmer(e, h, t, kmer);
						if (n == p->k) {
							orc_ntmc64_seed(kmer, p->k, p->h, fh, rh, e->hv);
						}
						skipped = 1;
					}
				}
			}
		}
		if (!skipped) {
			rope_insert(e, &t->ni, t->pos, s->best_indel, s->best_indel_len, s->best_support);
			orc_ntmc64_changelast(draft, (unsigned char)s->best_indel[0], p->k, p->h, fh, rh, e->hv);
		}
		break;
	}
	case 3:
		rope_delete(e, &t->ni, &t->pos, (unsigned)s->best_indel_len, s->best_support);
		orc_ntmc64_changelast(draft, cursor_char(e, t), p->k, p->h, fh, rh, e->hv);
		break;
	case 0:
		if (p->mask) {
			if (tnode.node_type == 0) {
				e->seq[t->pos] = (char)tolower(draft);
			} else if (tnode.node_type == 1) {
				e->rope.v[t->ni].c = (uint8_t)tolower(draft);
			}
			orc_ntmc64_changelast(draft, (unsigned char)tolower(draft), p->k, p->h, fh, rh, e->hv);
		}
		if (p->snv) {
			orc_srec r;
			memset(&r, 0, sizeof r);
			r.draft_char = draft;
			r.pos = t->pos;
			r.sub_base = draft;
			r.num_support = s->best_support;
			r.altbase1 = s->altbase1;
			r.altsupp1 = s->altsupp1;
			r.altbase2 = s->altbase2;
			r.altsupp2 = s->altsupp2;
			r.altbase3 = s->altbase3;
			r.altsupp3 = s->altsupp3;
			if (s->altsupp1) {
				sv_push(&e->recs, r);
			}
		}
		break;
	default:
		break;
	}
}

static int
meets_edit_threshold(const orc_params* p, unsigned count)
{
	/* ntedit.cpp:1659-1663, 1892-1897, 1992-1997 */
	if (!p->use_ratio) {
		return (float)count >= ((float)p->k / p->edit_threshold);
	}
	return (float)count >= ((float)p->k / p->jump) * p->edit_ratio;
}

/* ntedit.cpp:1451-1545 */
static int
try_deletion(engine* e, unsigned char draft, unsigned num_del, const cursor* h0, const cursor* t0, uint64_t fh, uint64_t rh,
             char* deleted, size_t* ndeleted)
{
	const orc_params* p = e->p;
	cursor h = *h0, t = *t0;
	unsigned char out = 0, in = 0;
	for (unsigned i = 0; i < num_del; i++) {
		deleted[(*ndeleted)++] = (char)cursor_char(e, &t);
		step(e, &t);
	}
	orc_ntmc64_changelast(draft, cursor_char(e, &t), p->k, p->h, &fh, &rh, e->hv);
	unsigned present = 0;
	if (present_solid(e)) {
		present++;
	}
	for (unsigned q = 1; q <= p->k - 2 && h.pos < e->len; q++) {
		if (roll_cursors(e, &h, &t, &out, &in)) {
			orc_ntmc64_roll(out, in, p->k, p->h, &fh, &rh, e->hv);
			if (q % p->jump == 0 && present_solid(e)) {
				present++;
			}
		}
	}
	int ok;
	if (!p->use_ratio) {
		ok = (float)present >= ((float)p->k / p->edit_threshold);
	} else {
		ok = (float)present >= (1 + ((float)p->k / p->jump)) * p->edit_ratio;
	}
	return ok ? (int)present : 0;
}

/* the q-th (0-based) string of ntedit.cpp:203-348 for a given first base: all strings of length 1..5
 * over ACGT starting with `first`, ordered by (length, lexicographic A<C<G<T). */
static size_t
indel_string(unsigned char first, unsigned q, char* out)
{
	static const unsigned start[6] = { 0, 1, 5, 21, 85, 341 };
	static const char alpha[4] = { 'A', 'C', 'G', 'T' };
	unsigned len = 1;
	while (q >= start[len]) {
		len++;
	}
	unsigned r = q - start[len - 1];
	out[0] = (char)first;
	for (unsigned i = len - 1; i >= 1; i--) {
		out[i] = alpha[r & 3];
		r >>= 2;
	}
	out[len] = 0;
	return len;
}

/* ntedit.cpp:1548-1744 */
static int
try_indels(engine* e, unsigned char draft, unsigned char index_char, unsigned* num_deletions, const cursor* h0,
           const cursor* t0, uint64_t fh0, uint64_t rh0, site* s, unsigned* alt_support_out)
{
	static const unsigned num_tries[6] = { 0, 1, 5, 21, 85, 341 };
	const orc_params* p = e->p;
	unsigned tb_support = 0, ta_support = 0, tb_type = 0;
	char tb_indel[16] = { 0 };
	size_t tb_len = 0;
	unsigned char out = 0, in = 0;

	for (unsigned i = 0; i < num_tries[p->max_insertions]; i++) {
		char ins[16];
		size_t nins = indel_string(index_char, i, ins);
		ins[nins++] = (char)draft;
		ins[nins] = 0;

		uint64_t fh = fh0, rh = rh0;
		cursor h = *h0, t = *t0;
		orc_ntmc64_changelast(draft, index_char, p->k, p->h, &fh, &rh, e->hv);
		unsigned present = 0;
		unsigned q = 0;
		for (; q < nins - 1 && h.pos < e->len; q++) {
			orc_ntmc64_roll(cursor_char(e, &h), (unsigned char)ins[q + 1], p->k, p->h, &fh, &rh, e->hv);
			step(e, &h);
			if (q % p->jump == 0 && present_solid(e)) {
				present++;
			}
		}
		for (; q < p->k - 1 && h.pos < e->len; q++) {
			if (roll_cursors(e, &h, &t, &out, &in)) {
				orc_ntmc64_roll(out, in, p->k, p->h, &fh, &rh, e->hv);
				if (q % p->jump == 0 && present_solid(e)) {
					present++;
				}
			}
		}
		nins--; /* drop the draft char again */
		ins[nins] = 0;
		if (meets_edit_threshold(p, present)) {
			if (p->mode == 0) {
				s->best_type = 2;
				memcpy(s->best_indel, ins, nins + 1);
				s->best_indel_len = nins;
				s->best_support = present;
				return 1;
			}
			if (present >= tb_support) {
				if (tb_support) {
					ta_support = tb_support;
				}
				tb_type = 2;
				memcpy(tb_indel, ins, nins + 1);
				tb_len = nins;
				tb_support = present;
			}
		}

		if (*num_deletions <= p->max_deletions) {
			char deleted[16];
			size_t nd = 0;
			unsigned del_support = (unsigned)try_deletion(e, draft, *num_deletions, h0, t0, fh0, rh0, deleted, &nd);
			deleted[nd] = 0;
			if (del_support > 0) {
				if (p->mode == 0) {
					s->best_type = 3;
					memcpy(s->best_indel, deleted, nd + 1);
					s->best_indel_len = nd;
					s->best_support = del_support;
					return 1;
				}
				if (del_support >= tb_support) {
					if (tb_support) {
						ta_support = tb_support;
					}
					tb_type = 3;
					memcpy(tb_indel, deleted, nd + 1);
					tb_len = nd;
					tb_support = del_support;
				}
			}
			(*num_deletions)++;
		}
	}

	if (tb_support > 0) {
		if ((p->mode == 2 && tb_support > s->best_support) || p->mode == 1) {
			s->best_type = tb_type;
			memcpy(s->best_indel, tb_indel, tb_len + 1);
			s->best_indel_len = tb_len;
			s->best_support = tb_support;
			*alt_support_out = ta_support;
		}
		return 1;
	}
	return 0;
}

/* substitution candidates, ntedit.cpp:178-199 */
static const char*
candidate_bases(int snv, unsigned char draft)
{
	if (snv) {
		switch (draft) {
		case 'A': return "TCG";
		case 'T': return "ACG";
		case 'C': return "ATG";
		case 'G': return "ATC";
		case 'R': case 'Y': case 'S': case 'W': case 'K': case 'M':
		case 'B': case 'D': case 'H': case 'V': case 'N':
			return "ATCG";
		default: return "";
		}
	}
	switch (draft) {
	case 'A': return "TCG";
	case 'T': return "ACG";
	case 'C': return "ATG";
	case 'G': return "ATC";
	case 'R': return "TC";
	case 'Y': return "AG";
	case 'S': return "AT";
	case 'W': return "CG";
	case 'K': return "AC";
	case 'M': return "TG";
	case 'B': return "A";
	case 'D': return "C";
	case 'H': return "G";
	case 'V': return "T";
	case 'N': return "ATCG";
	default: return "";
	}
}

/* ntedit.cpp:1747-2151 (without the writer) */
int
orc_polish_contig(char* seq, uint32_t len, const orc_filter* bloom, const orc_filter* bloomrep, const orc_params* p,
                  orc_result* res)
{
	engine E;
	memset(&E, 0, sizeof E);
	E.seq = seq;
	E.len = len;
	E.bloom = bloom;
	E.rep = bloomrep;
	E.p = p;
	engine* e = &E;
	const unsigned k = p->k;
	if (p->h > 64 || len == 0) {
		return -1;
	}

	uint64_t fh = 0, rh = 0;
	unsigned char char_in = 0, char_out = 0;

	cursor h = { first_accepted_kmer(e, 0), 0 };
	cursor t = { h.pos + k - 1, 0 };
	if ((uint64_t)h.pos + k - 1 < len) {
		orc_ntmc64_seed(seq + h.pos, k, p->h, &fh, &rh, e->hv);
		char_in = (unsigned char)seq[t.pos];
	}
	orc_node root = { 0, 0, len - 1, 0, 0 };
	nv_push(&e->rope, root);

	int keep_going = 1;
	uint8_t medbuf[1024];
	/* ntedit.cpp:1881-1885 declares best_sub_base / altbase1..3 WITHOUT initialisers inside the loop
	 * body; the compiled reference keeps them in fixed slots, so a site that reads one before writing
	 * it (possible in mode 2 via tryIndels' altsupp1 side channel) sees the value the previous site left
	 * behind.  Restated as state that persists across sites of one contig, starting at 0. */
	unsigned char stale_best_sub = 0, stale_alt1 = 0, stale_alt2 = 0, stale_alt3 = 0;
	do {
		if ((uint64_t)h.pos + k - 1 >= len) {
			break;
		}
		if (p->snv || !orc_filter_contains(bloom, e->hv) ||
		    (bloom->counting && orc_filter_count(bloom, e->hv) < p->min_threshold)) {
			uint64_t tfh = fh, trh = rh;
			cursor th = h, tt = t;
			const unsigned char draft = (unsigned char)toupper(char_in);

			unsigned missing = 0, there = 0, there_median = 0;
			size_t nmed = 0;
			int do_not_fix = 0;
			for (unsigned q = 0; q < k && th.pos < len; q++) {
				if (roll_cursors(e, &th, &tt, &char_out, &char_in)) {
					orc_ntmc64_roll(char_out, char_in, k, p->h, &tfh, &trh, e->hv);
					if (!is_accepted((unsigned char)toupper(char_in))) {
						do_not_fix = 1;
						break;
					}
					if (q % p->jump == 0 && !orc_filter_contains(bloom, e->hv)) {
						missing++;
					} else if (is_atgc(draft) && q % p->jump == 0 && orc_filter_contains(bloom, e->hv) &&
					           (!bloom->counting || orc_filter_count(bloom, e->hv) >= p->min_threshold)) {
						there++;
						if (bloom->counting && nmed < sizeof medbuf) {
							medbuf[nmed++] = (uint8_t)orc_filter_count(bloom, e->hv);
						}
					}
				} else {
					do_not_fix = 1;
					break;
				}
			}
			if (bloom->counting) {
				there_median = median_u8(medbuf, nmed);
			}
			int attempt = p->snv ||
			              (!do_not_fix &&
			               ((!p->use_ratio && (float)missing >= ((float)k / p->missing_threshold)) ||
			                (p->use_ratio && (float)missing >= (((float)k / p->jump) * p->missing_ratio)) ||
			                (bloom->counting && there_median < p->min_threshold)));
			if (attempt) {
				unsigned num_deletions = 1;
				site s;
				memset(&s, 0, sizeof s);
				s.best_sub = stale_best_sub;
				s.altbase1 = stale_alt1;
				s.altbase2 = stale_alt2;
				s.altbase3 = stale_alt3;
				if (p->snv && meets_edit_threshold(p, there)) {
					s.best_sub = draft;
					s.best_support = bloom->counting ? there_median : there;
				}
				const char* cands = candidate_bases(p->snv, draft);
				for (const char* cp = cands; *cp; cp++) {
					const unsigned char sub = (unsigned char)*cp;
					tfh = fh;
					trh = rh;
					orc_ntmc64_changelast(draft, sub, k, p->h, &tfh, &trh, e->hv);
					if (!(present_solid(e) || p->mode == 2)) {
						continue;
					}
					th = h;
					tt = t;
					/* patch the candidate base in, ntedit.cpp:1936-1940 */
					if (e->rope.v[t.ni].node_type == 0) {
						seq[tt.pos] = (char)sub;
					} else if (e->rope.v[t.ni].node_type == 1) {
						e->rope.v[t.ni].c = sub;
					}
					unsigned present = 0;
					for (unsigned q = 0; q < k && th.pos < len && tt.pos < len; q++) {
						if (!roll_cursors(e, &th, &tt, &char_out, &char_in)) {
							break;
						}
						orc_ntmc64_roll(char_out, char_in, k, p->h, &tfh, &trh, e->hv);
						if (q % p->jump == 0 && present_solid(e)) {
							present++;
						}
					}
					/* revert with the UPPER-CASED draft char, ntedit.cpp:1975-1981 */
					if (e->rope.v[t.ni].node_type == 0) {
						seq[t.pos] = (char)draft;
					} else if (e->rope.v[t.ni].node_type == 1) {
						e->rope.v[t.ni].c = draft;
					}
					if (meets_edit_threshold(p, present)) {
						if (present >= s.best_support) {
							if (s.altsupp2) {
								s.altbase3 = s.altbase2;
								s.altsupp3 = s.altsupp2;
							}
							if (s.altsupp1) {
								s.altbase2 = s.altbase1;
								s.altsupp2 = s.altsupp1;
							}
							if (s.best_support) {
								s.altsupp1 = s.best_support;
								s.altbase1 = s.best_sub;
							}
							s.best_type = 1;
							s.best_sub = sub;
							s.best_support = present;
						} else if (!s.altsupp1) {
							s.altbase1 = sub;
							s.altsupp1 = present;
						} else if (!s.altsupp2) {
							if (present < s.altsupp1) {
								s.altbase2 = sub;
								s.altsupp2 = present;
							} else {
								s.altbase2 = s.altbase1;
								s.altsupp2 = s.altsupp1;
								s.altbase1 = sub;
								s.altsupp1 = present;
							}
						} else if (!s.altsupp3) {
							if (present < s.altsupp2) {
								s.altbase3 = sub;
								s.altsupp3 = present;
							} else if (present < s.altsupp1) {
								s.altbase3 = s.altbase2;
								s.altsupp3 = s.altsupp2;
								s.altbase2 = sub;
								s.altsupp2 = present;
							} else {
								s.altbase3 = s.altbase2;
								s.altsupp3 = s.altsupp2;
								s.altbase2 = s.altbase1;
								s.altsupp2 = s.altsupp1;
								s.altbase1 = sub;
								s.altsupp1 = present;
							}
						}
						if (p->mode == 0 || p->mode == 1) {
							continue;
						}
					}
					if (p->mode == 2 || s.best_type != 1) {
						if (try_indels(e, draft, sub, &num_deletions, &h, &t, fh, rh, &s, &s.altsupp1)) {
							if (p->mode == 0 || p->mode == 1) {
								break;
							}
						}
					}
				}
				apply_edit(e, draft, &s, &h, &t, &fh, &rh);
				stale_best_sub = s.best_sub;
				stale_alt1 = s.altbase1;
				stale_alt2 = s.altbase2;
				stale_alt3 = s.altbase3;
			}
		}
		/* advance, skipping k positions after every non-accepted incoming base, ntedit.cpp:2118-2138 */
		int target = -1;
		do {
			if (roll_cursors(e, &h, &t, &char_out, &char_in)) {
				if (!is_accepted((unsigned char)toupper(char_in))) {
					target = (int)t.pos + (int)k;
				}
				orc_ntmc64_roll(char_out, char_in, k, p->h, &fh, &rh, e->hv);
			} else {
				keep_going = 0;
				break;
			}
		} while (target >= 0 && (int)t.pos != target);
	} while (keep_going);

	res->nodes = e->rope.v;
	res->n_nodes = e->rope.n;
	res->srecs = e->recs.v;
	res->n_srecs = e->recs.n;
	return 0;
}

void
orc_result_free(orc_result* r)
{
	free(r->nodes);
	free(r->srecs);
	memset(r, 0, sizeof *r);
}

/* =====================================================================================
 * Writer restatement (ntedit.cpp:925-1213, header :2175-2188)
 * ===================================================================================== */
typedef struct sbuf {
	char* s;
	size_t n, cap;
} sbuf;

static void
sb_add(sbuf* b, const char* s, size_t n)
{
	if (b->n + n + 1 > b->cap) {
		size_t c = b->cap ? b->cap : 256;
		while (c < b->n + n + 1) {
			c *= 2;
		}
		b->s = (char*)realloc(b->s, c);
		b->cap = c;
	}
	memcpy(b->s + b->n, s, n);
	b->n += n;
	b->s[b->n] = 0;
}

static void
sb_str(sbuf* b, const char* s)
{
	sb_add(b, s, strlen(s));
}

static void
sb_chr(sbuf* b, char c)
{
	sb_add(b, &c, 1);
}

static void
sb_u(sbuf* b, unsigned long long v)
{
	char t[32];
	int n = snprintf(t, sizeof t, "%llu", v);
	sb_add(b, t, (size_t)n);
}

int
orc_tsv_header(const orc_params* p, int counting, char* buf, size_t buflen)
{
	const char* evi = counting ? "Coverage" : "Support";
	char col[128];
	if (counting) {
		snprintf(col, sizeof col, "Coverage (max 255)");
	} else {
		/* ostream << double prints ceil(k/j) with %g formatting */
		snprintf(col, sizeof col, "Support %u-mer (out of %g)", p->k, ceil((double)p->k / (double)p->jump));
	}
	int n = snprintf(buf, buflen, "ID\tbpPosition+1\tOriginalBase\tNewBase\t%s\tAlt.Base1\tAlt.%s1\tAlt.Base2\tAlt.%s2\tAlt.Base3\tAlt.%s3\n",
	                 col, evi, evi, evi);
	return n > 0 && (size_t)n < buflen ? n : -1;
}

static void
vcf_sub_row(sbuf* vcf, const char* hdr, const orc_srec* r, const orc_params* p)
{
	/* ntedit.cpp:986-1162 with an empty ClinVar map: every lookup yields "^NA".  Bases are handled as
	 * (bytes, length) because a stale alt base can be NUL (see the note in orc_polish_contig). */
	int edit_row = !(p->snv && r->draft_char == r->sub_base);
	char altb[3];
	unsigned alts[3];
	int na = 0;
	if (r->altsupp1 > 0) { altb[na] = (char)r->altbase1; alts[na++] = r->altsupp1; }
	if (r->altsupp2 > 0) { altb[na] = (char)r->altbase2; alts[na++] = r->altsupp2; }
	if (r->altsupp3 > 0) { altb[na] = (char)r->altbase3; alts[na++] = r->altsupp3; }
	char base[4];
	size_t nbase = 1;
	char support[64];
	int nclin = 1 + (edit_row ? 1 : 0);
	const char* gt;
	base[0] = (char)r->sub_base;
	snprintf(support, sizeof support, "%u", r->num_support);
	unsigned best_alt_supp = 0;
	char best_alt = '1';
	if (na > 0) {
		if (p->snv) {
			if (!edit_row) {
				for (int i = 0; i < na; i++) {
					if (alts[i] > best_alt_supp) { best_alt_supp = alts[i]; best_alt = altb[i]; }
				}
				base[0] = best_alt;
				nclin++;
				snprintf(support, sizeof support, "%u,%u", r->num_support, best_alt_supp);
				gt = "0/1";
			} else {
				int ref = 0;
				for (int i = 0; i < na; i++) {
					if (r->draft_char == (unsigned char)altb[i]) { best_alt_supp = alts[i]; ref = 1; break; }
					if (alts[i] > best_alt_supp) { best_alt_supp = alts[i]; best_alt = altb[i]; }
				}
				if (ref) {
					snprintf(support, sizeof support, "%u,%u", best_alt_supp, r->num_support);
					gt = "0/1";
				} else {
					gt = "1/2";
					snprintf(support, sizeof support, "%u,%u", r->num_support, best_alt_supp);
					base[1] = ',';
					base[2] = best_alt;
					nbase = 3;
					nclin++;
				}
			}
		} else {
			for (int i = 0; i < na; i++) {
				if (r->draft_char == (unsigned char)altb[i]) { continue; }
				if (alts[i] > best_alt_supp) { best_alt_supp = alts[i]; best_alt = altb[i]; }
			}
			gt = "1/2";
			snprintf(support, sizeof support, "%u,%u", r->num_support, best_alt_supp);
			base[1] = ',';
			base[2] = best_alt;
			nbase = 3;
			nclin++;
		}
	} else {
		gt = "1/1";
	}
	sb_str(vcf, hdr); sb_chr(vcf, '\t'); sb_u(vcf, (unsigned long long)r->pos + 1); sb_str(vcf, "\t.\t");
	sb_chr(vcf, (char)r->draft_char); sb_chr(vcf, '\t'); sb_add(vcf, base, nbase); sb_str(vcf, "\t.\tPASS\tAD=");
	sb_str(vcf, support);
	for (int i = 0; i < nclin; i++) {
		sb_str(vcf, "^NA");
	}
	sb_str(vcf, "\tGT\t"); sb_str(vcf, gt); sb_chr(vcf, '\n');
}

int
orc_write_contig(const char* hdr, const char* seq, uint32_t len, const orc_result* r, const orc_params* p, char** fa_out,
                 size_t* fa_len, char** tsv_out, size_t* tsv_len, char** vcf_out, size_t* vcf_len)
{
	sbuf fa = { 0, 0, 0 }, tsv = { 0, 0, 0 }, vcf = { 0, 0, 0 };
	sb_str(&fa, ">");
	sb_str(&fa, hdr);
	sb_str(&fa, "\n");
	sb_str(&tsv, "");
	sb_str(&vcf, "");
	size_t ni = 0, si = 0;
	char ins[4096];
	size_t nins = 0;
	long ins_support = -1;
	uint32_t pos = 0;
	(void)len;
	if (r->n_nodes == 0) {
		return -1;
	}
	orc_node cur = r->nodes[0];
	while (ni < r->n_nodes && cur.node_type != -1) {
		if (cur.node_type == 0) {
			if (nins > 0) {
				char draft = seq[cur.s_pos - nins];
				sb_str(&tsv, hdr); sb_chr(&tsv, '\t'); sb_u(&tsv, pos); sb_chr(&tsv, '\t'); sb_chr(&tsv, draft);
				sb_str(&tsv, "\t+"); sb_add(&tsv, ins, nins); sb_chr(&tsv, '\t');
				{
					char t[32];
					int n = snprintf(t, sizeof t, "%ld", ins_support);
					sb_add(&tsv, t, (size_t)n);
				}
				sb_chr(&tsv, '\n');
				sb_str(&vcf, hdr); sb_chr(&vcf, '\t'); sb_u(&vcf, pos); sb_str(&vcf, "\t.\t"); sb_chr(&vcf, draft);
				sb_chr(&vcf, '\t'); sb_chr(&vcf, draft); sb_add(&vcf, ins, nins); sb_str(&vcf, "\t.\tPASS\tAD=");
				{
					char t[32];
					int n = snprintf(t, sizeof t, "%ld", ins_support);
					sb_add(&vcf, t, (size_t)n);
				}
				sb_str(&vcf, "^NA\tGT\t1/1\n");
				nins = 0;
				ins_support = -1;
			}
			while (si < r->n_srecs && r->srecs[si].pos <= cur.e_pos) {
				const orc_srec* s = &r->srecs[si];
				int edit_row = !(p->snv && s->draft_char == s->sub_base);
				if (edit_row) {
					sb_str(&tsv, hdr); sb_chr(&tsv, '\t'); sb_u(&tsv, (unsigned long long)s->pos + 1); sb_chr(&tsv, '\t');
					sb_chr(&tsv, (char)s->draft_char); sb_chr(&tsv, '\t'); sb_chr(&tsv, (char)s->sub_base); sb_chr(&tsv, '\t');
					sb_u(&tsv, s->num_support);
					if (s->altsupp1 > 0) { sb_chr(&tsv, '\t'); sb_chr(&tsv, (char)s->altbase1); sb_chr(&tsv, '\t'); sb_u(&tsv, s->altsupp1); }
					if (s->altsupp2 > 0) { sb_chr(&tsv, '\t'); sb_chr(&tsv, (char)s->altbase2); sb_chr(&tsv, '\t'); sb_u(&tsv, s->altsupp2); }
					if (s->altsupp3 > 0) { sb_chr(&tsv, '\t'); sb_chr(&tsv, (char)s->altbase3); sb_chr(&tsv, '\t'); sb_u(&tsv, s->altsupp3); }
					sb_chr(&tsv, '\n');
				}
				vcf_sub_row(&vcf, hdr, s, p);
				si++;
			}
			sb_add(&fa, seq + cur.s_pos, (size_t)cur.e_pos - cur.s_pos + 1);
			pos = cur.e_pos + 1;
		} else if (cur.node_type == 1) {
			if (nins + 1 < sizeof ins) {
				ins[nins++] = (char)cur.c;
			}
			if (ins_support == -1) {
				ins_support = (long)cur.num_support;
			}
			sb_chr(&fa, (char)cur.c);
		}
		ni++;
		if (ni < r->n_nodes) {
			cur = r->nodes[ni];
			if (cur.node_type == 0 && cur.s_pos != pos) {
				sb_str(&tsv, hdr); sb_chr(&tsv, '\t'); sb_u(&tsv, pos); sb_chr(&tsv, '\t'); sb_chr(&tsv, seq[pos]);
				sb_str(&tsv, "\t-"); sb_add(&tsv, seq + pos, (size_t)cur.s_pos - pos); sb_chr(&tsv, '\t');
				sb_u(&tsv, cur.num_support); sb_chr(&tsv, '\n');
				sb_str(&vcf, hdr); sb_chr(&vcf, '\t'); sb_u(&vcf, pos); sb_str(&vcf, "\t.\t");
				sb_add(&vcf, seq + pos - 1, (size_t)cur.s_pos - pos + 1); sb_chr(&vcf, '\t'); sb_chr(&vcf, seq[pos - 1]);
				sb_str(&vcf, "\t.\tPASS\tAD="); sb_u(&vcf, cur.num_support); sb_str(&vcf, "^NA\tGT\t1/1\n");
			}
		}
	}
	sb_str(&fa, "\n");
	*fa_out = fa.s; *fa_len = fa.n;
	*tsv_out = tsv.s; *tsv_len = tsv.n;
	*vcf_out = vcf.s; *vcf_len = vcf.n;
	return 0;
}
