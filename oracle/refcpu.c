/*
 * refcpu.c -- CPU oracle: plain-C restatement of AchilleBailly/zstd-decompressor.
 * TEST INFRASTRUCTURE ONLY (see refcpu.h).  Parity: pinned by the reference's test vectors and
 * by libzstd on the reference fixtures; XXH64 boundary unpinned by the reference (refcpu.h).
 *
 * The structure deliberately keeps the reference's algorithmic shape so that it can also stand in
 * as "the reference CPU path" when timed: bit-at-a-time Huffman tree walk (huffman.rs:205-218),
 * reversed-copy backward reader (parsing.rs:208), linear code-table search (sequence.rs:36),
 * O(symbols x table) FSE grouping (fse.rs:169-189), byte-wise match copy
 * (decoding_context.rs:95-98), eager parse of every block before decode (frame.rs:198-230).
 */
#include "refcpu.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAX_WIN_SIZE ((uint64_t)8 << 20) /* frame.rs:44 */
#define MAX_AL 9                         /* fse.rs:13 */
#define MAX_SYMBOL 256                   /* fse.rs:14 */

static int fail(rc_error *e, int code, uint64_t a, uint64_t b) {
    if (e) { e->code = code; e->a = a; e->b = b; }
    return code;
}
#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

/* growable byte vector == Vec<u8> */
typedef struct { uint8_t *p; size_t len, cap; } vec8;
static void v8_reserve(vec8 *v, size_t extra) {
    if (v->len + extra <= v->cap) return;
    size_t nc = v->cap ? v->cap : 256;
    while (nc < v->len + extra) nc *= 2;
    v->p = (uint8_t *)realloc(v->p, nc);
    v->cap = nc;
}
static void v8_push(vec8 *v, uint8_t b) { v8_reserve(v, 1); v->p[v->len++] = b; }
static void v8_extend(vec8 *v, const uint8_t *s, size_t n) {
    if (!n) return;
    v8_reserve(v, n); memcpy(v->p + v->len, s, n); v->len += n;
}

/* ===== utils.rs ======================================================== */
/* min_bits_required utils.rs:25-31 ; discrete_log2 utils.rs:33-40 (asserts value > 0) */
static unsigned min_bits_required(uint64_t v) {
    if (v == 0) return 1;
    unsigned n = 0; while (v) { n++; v >>= 1; } return n;
}
/* returns -1 where the reference's assert! would panic */
static int discrete_log2(uint64_t v) { return v == 0 ? -1 : (int)min_bits_required(v) - 1; }

/* ===== parsing.rs : ForwardByteParser (parsing.rs:9,29-112) ============ */
typedef struct { const uint8_t *p; size_t n; } bytep;
static int bp_u8(bytep *s, uint8_t *o, rc_error *e) {                   /* :39-50 */
    if (s->n == 0) return fail(e, RC_NotEnoughBytes, 1, 0);
    *o = *s->p; s->p++; s->n--; return 0;
}
static int bp_slice(bytep *s, size_t len, const uint8_t **o, rc_error *e) { /* :63-79 */
    if (len == 0) return fail(e, RC_EmptySliceError, 0, 0);
    if (s->n < len) return fail(e, RC_NotEnoughBytes, len, s->n);
    *o = s->p; s->p += len; s->n -= len; return 0;
}
static int bp_le_u32(bytep *s, uint32_t *o, rc_error *e) {               /* :82-95 */
    if (s->n < 4) return fail(e, RC_NotEnoughBytes, 4, s->n);
    *o = (uint32_t)s->p[0] | (uint32_t)s->p[1] << 8 | (uint32_t)s->p[2] << 16 | (uint32_t)s->p[3] << 24;
    s->p += 4; s->n -= 4; return 0;
}
static int bp_le_u16(bytep *s, uint16_t *o, rc_error *e) {               /* :98-111 */
    if (s->n < 2) return fail(e, RC_NotEnoughBytes, 2, s->n);
    *o = (uint16_t)(s->p[0] | s->p[1] << 8);
    s->p += 2; s->n -= 2; return 0;
}

/* ===== parsing.rs : ForwardBitParser (parsing.rs:114-189) ============== */
/* bitbuffer 0.10.9 LittleEndian read_int(pos,len): bit i of the stream is bit (i&7) of byte i>>3,
 * the value's bit 0 is the first bit read.  Pinned by tests/parsing.rs. */
struct rc_fwd_bits { const uint8_t *d; size_t bit_len, readable, pos; };
static int fwd_init(rc_fwd_bits *f, const uint8_t *d, size_t n, rc_error *e) { /* :131-140 */
    if (n == 0) return fail(e, RC_EmptyInputData, 0, 0);
    f->d = d; f->bit_len = n * 8; f->readable = n * 8; f->pos = 0; return 0;
}
static uint64_t le_bits(const uint8_t *d, size_t pos, size_t len) {
    uint64_t r = 0;
    for (size_t i = 0; i < len; i++) { size_t b = pos + i; r |= (uint64_t)((d[b >> 3] >> (b & 7)) & 1) << i; }
    return r;
}
static int fwd_take(rc_fwd_bits *f, size_t len, uint64_t *o, rc_error *e) {   /* :152-170 */
    if (f->bit_len - f->pos < len) return fail(e, RC_NotEnoughBits, len, f->readable);
    if (len > 64) return fail(e, RC_MaximumReadableBitsExceeded, len, 0);
    *o = le_bits(f->d, f->pos, len); f->readable -= len; f->pos += len; return 0;
}
static int fwd_peek(const rc_fwd_bits *f, size_t len, uint64_t *o, rc_error *e) { /* :173-188 */
    if (f->bit_len - f->pos < len) return fail(e, RC_NotEnoughBits, len, f->bit_len);
    if (len > 64) return fail(e, RC_MaximumReadableBitsExceeded, len, 0);
    *o = le_bits(f->d, f->pos, len); return 0;
}
static size_t fwd_bytes_read(const rc_fwd_bits *f) { return f->pos / 8 + (f->pos % 8 > 0); } /* :122-126 */

/* ===== parsing.rs : BackwardBitParser (parsing.rs:191-259) ============= */
/* The reference reverses a copy of the stream and reads it big-endian: bit i of the reversed buffer
 * is bit (7 - (i&7)) of reversed byte i>>3; the value's MSB is the first bit read. */
struct rc_bwd_bits { uint8_t *d; size_t n, readable, pos; };
static int bwd_init(rc_bwd_bits *b, const uint8_t *data, size_t n, rc_error *e) { /* :200-220 */
    b->d = NULL;
    if (n == 0) return fail(e, RC_EmptyInputData, 0, 0);
    if (data[n - 1] == 0) return fail(e, RC_NullByte, 0, 0);
    b->d = (uint8_t *)malloc(n);
    for (size_t i = 0; i < n; i++) b->d[i] = data[n - 1 - i];   /* .iter().rev().copied().collect() */
    size_t i = 1;
    while ((b->d[0] & (1u << (8 - i))) == 0) i++;
    b->n = n; b->readable = n * 8 - i; b->pos = i; return 0;
}
static void bwd_drop(rc_bwd_bits *b) { free(b->d); b->d = NULL; }
static int bwd_take(rc_bwd_bits *b, size_t len, uint64_t *o, rc_error *e) {     /* :228-254 */
    if (b->n * 8 - b->pos < len) return fail(e, RC_NotEnoughBits, len, b->readable);
    if (len > 64) return fail(e, RC_MaximumReadableBitsExceeded, len, 0);
    if (len == 0) { *o = 0; return 0; }
    uint64_t r = 0;
    for (size_t i = 0; i < len; i++) { size_t p = b->pos + i; r = (r << 1) | ((b->d[p >> 3] >> (7 - (p & 7))) & 1); }
    b->readable -= len; b->pos += len; *o = r; return 0;
}

/* public handles for tests/parsing.rs vectors */
rc_fwd_bits *rc_fwd_new(const uint8_t *data, size_t n, rc_error *e) {
    rc_fwd_bits *f = (rc_fwd_bits *)calloc(1, sizeof *f);
    if (fwd_init(f, data, n, e)) { free(f); return NULL; } return f;
}
void rc_fwd_free(rc_fwd_bits *f) { free(f); }
uint64_t rc_fwd_take(rc_fwd_bits *f, size_t len, rc_error *e) { uint64_t v = 0; e->code = 0; fwd_take(f, len, &v, e); return v; }
uint64_t rc_fwd_peek(rc_fwd_bits *f, size_t len, rc_error *e) { uint64_t v = 0; e->code = 0; fwd_peek(f, len, &v, e); return v; }
size_t rc_fwd_len(const rc_fwd_bits *f) { return f->readable; }
size_t rc_fwd_bytes_read(const rc_fwd_bits *f) { return fwd_bytes_read(f); }
rc_bwd_bits *rc_bwd_new(const uint8_t *data, size_t n, rc_error *e) {
    rc_bwd_bits *b = (rc_bwd_bits *)calloc(1, sizeof *b);
    if (bwd_init(b, data, n, e)) { free(b); return NULL; } return b;
}
void rc_bwd_free(rc_bwd_bits *b) { if (b) { bwd_drop(b); free(b); } }
uint64_t rc_bwd_take(rc_bwd_bits *b, size_t len, rc_error *e) { uint64_t v = 0; e->code = 0; bwd_take(b, len, &v, e); return v; }
size_t rc_bwd_len(const rc_bwd_bits *b) { return b->readable; }

/* ===== decoders/fse.rs ================================================= */
typedef struct { uint16_t output, baseline, bits_to_read; } fse_state;   /* fse.rs:72-76 */
typedef struct { fse_state *t; size_t size; uint8_t al; } fse_table;     /* fse.rs:86-89 */
#define DIST_CAP 600

/* parse_fse_table fse.rs:16-69 */
static int parse_fse_table(rc_fwd_bits *in, uint8_t *al_out, int16_t *dist, size_t *ndist, rc_error *e) {
    uint64_t v;
    TRY(fwd_take(in, 4, &v, e));
    unsigned al = (unsigned)v + 5;
    if (al > MAX_AL) return fail(e, RC_LargeAccuracyLog, al, 0);
    size_t nd = 0; int32_t remaining = 1 << al; size_t n_sym = 0;
    while (remaining > 0 && n_sym < MAX_SYMBOL) {
        size_t bits_to_read = (size_t)discrete_log2((uint64_t)(remaining + 1)) + 1;
        uint64_t pk; TRY(fwd_peek(in, bits_to_read, &pk, e));
        uint16_t peeked = (uint16_t)pk;
        uint16_t lower_mask = (uint16_t)((1u << (bits_to_read - 1)) - 1);
        uint16_t threshold = (uint16_t)((1u << bits_to_read) - 1 - ((uint16_t)remaining + 1));
        int16_t decoded;
        if ((peeked & lower_mask) < threshold) { TRY(fwd_take(in, bits_to_read - 1, &v, e)); decoded = (int16_t)v; }
        else if (peeked > lower_mask) { TRY(fwd_take(in, bits_to_read, &v, e)); decoded = (int16_t)((int16_t)v - (int16_t)threshold); }
        else { TRY(fwd_take(in, bits_to_read, &v, e)); decoded = (int16_t)v; }
        int16_t proba = (int16_t)(decoded - 1);
        remaining -= proba < 0 ? -proba : proba;
        dist[nd++] = proba; n_sym++;
        if (proba == 0) {
            for (;;) {
                TRY(fwd_take(in, 2, &v, e));
                for (uint64_t z = 0; z < v; z++) dist[nd++] = 0;
                n_sym += (size_t)v;
                if (v != 3) break;
            }
        }
    }
    if (remaining != 0 || n_sym >= MAX_SYMBOL) return fail(e, RC_CorruptedTable, 0, 0);
    *al_out = (uint8_t)al; *ndist = nd; return 0;
}

/* FseTable::from_distribution fse.rs:110-202 */
static int fse_from_distribution(uint8_t al, const int16_t *dist, size_t nsym, fse_table *out, rc_error *e) {
    if (al > MAX_AL) return fail(e, RC_LargeAccuracyLog, al, 0);
    size_t N = (size_t)1 << al;
    fse_state *t = (fse_state *)calloc(N, sizeof *t);
    uint8_t *filled = (uint8_t *)calloc(N, 1);
    size_t zero_pos = N;
    for (size_t s = 0; s < nsym; s++)                      /* :120-133 "less than one" symbols from the end */
        if (dist[s] == -1) {
            if (zero_pos == 0) { free(t); free(filled); return fail(e, RC_Panic, 10, 0); } /* usize underflow */
            zero_pos--; t[zero_pos].output = (uint16_t)s; t[zero_pos].baseline = 0; t[zero_pos].bits_to_read = al; filled[zero_pos] = 2;
        }
    size_t position = 0, step = (N >> 1) + (N >> 3) + 3, mask = N - 1;   /* :136-157 */
    for (size_t s = 0; s < nsym; s++) {
        if (dist[s] <= 0) continue;
        for (int k = 0; k < dist[s]; k++) {
            t[position].output = (uint16_t)s; filled[position] = 1;
            position = (position + step) & mask;
            size_t guard = 0;
            while (position >= zero_pos) { position = (position + step) & mask; if (++guard > N) { free(t); free(filled); return fail(e, RC_Panic, 11, 0); } }
        }
    }
    for (size_t i = 0; i < N; i++)                         /* :160-166 */
        if (!filled[i]) { free(t); free(filled); return fail(e, RC_CorruptedTable, 0, 0); }
    /* :169-189 per-symbol grouping in state order (O(symbols x N) like the reference) */
    size_t *grouped = (size_t *)malloc(N * sizeof(size_t));
    for (size_t symbol = 0; symbol < nsym; symbol++) {
        size_t num_states = 0;
        for (size_t i = 0; i < N; i++) if (t[i].output == symbol) grouped[num_states++] = i;
        if (num_states == 0) continue;       /* parts=1, loop 1..1 is empty */
        size_t parts = 1; while (parts < num_states) parts <<= 1;   /* 1 << ceil(log2(n)) */
        size_t base_width = N / parts;
        int lg = discrete_log2(base_width);
        if (lg < 0) { free(grouped); free(t); free(filled); return fail(e, RC_Panic, 12, 0); }
        unsigned base_nb = (unsigned)lg; uint16_t baseline = 0;
        for (size_t i = parts - num_states; i < parts; i++) {
            size_t new_i = i % num_states;
            unsigned add = new_i != i ? 1 : 0, mult = new_i != i ? 2 : 1;
            t[grouped[new_i]].bits_to_read = (uint16_t)(base_nb + add);
            t[grouped[new_i]].baseline = baseline;
            baseline = (uint16_t)(baseline + (uint16_t)base_width * mult);
        }
    }
    free(grouped); free(filled);
    out->t = t; out->size = N; out->al = al; return 0;
}
static int fse_table_parse(rc_fwd_bits *in, fse_table *out, rc_error *e) {          /* fse.rs:204-208 */
    int16_t dist[DIST_CAP]; size_t nd; uint8_t al;
    TRY(parse_fse_table(in, &al, dist, &nd, e));
    return fse_from_distribution(al, dist, nd, out, e);
}
static fse_table fse_clone(const fse_table *s) {
    fse_table c = *s; c.t = (fse_state *)malloc(s->size * sizeof(fse_state));
    memcpy(c.t, s->t, s->size * sizeof(fse_state)); return c;
}

/* trait BitDecoder<u16> decoders/mod.rs:28-64, implemented by FseDecoder (fse.rs:279-323) and
 * RLEDecoder (rle.rs:6-34).  kind 0 = FSE, 1 = RLE. */
typedef struct { int kind; fse_table table; size_t cur_state; int has_sym; uint16_t next_symbol; uint8_t byte; } bitdec;
static int bd_initialize(bitdec *d, rc_bwd_bits *bs, rc_error *e) {
    if (d->kind == 1) return 0;
    uint64_t st; TRY(bwd_take(bs, d->table.al, &st, e));          /* fse.rs:280-288 */
    if (st >= d->table.size) return fail(e, RC_Panic, 20, 0);
    d->next_symbol = d->table.t[st].output; d->has_sym = 1; d->cur_state = (size_t)st; return 0;
}
static size_t bd_expected_bits(const bitdec *d) { return d->kind == 1 ? 0 : d->table.t[d->cur_state].bits_to_read; }
static int bd_symbol(bitdec *d, uint16_t *o, rc_error *e) {
    if (d->kind == 1) { *o = d->byte; return 0; }
    if (!d->has_sym) return fail(e, RC_Panic, 21, 0);              /* fse.rs:295-297 */
    *o = d->next_symbol; d->has_sym = 0; return 0;
}
static int bd_update_bits(bitdec *d, rc_bwd_bits *bs, rc_error *e) {
    if (d->kind == 1) return 0;
    if (d->has_sym) return fail(e, RC_Panic, 22, 0);               /* fse.rs:306-308 */
    uint64_t v; TRY(bwd_take(bs, bd_expected_bits(d), &v, e));
    size_t ns = (size_t)v + d->table.t[d->cur_state].baseline;     /* fse.rs:310-311 */
    if (ns >= d->table.size) return fail(e, RC_Panic, 23, 0);      /* index out of bounds */
    d->next_symbol = d->table.t[ns].output; d->has_sym = 1; d->cur_state = ns; return 0;
}

/* AlternatingDecoder alternating.rs:8-69 */
typedef struct { bitdec first, second; int last_updated_is_first, last_read_is_first; } altdec;
static int alt_initialize(altdec *a, rc_bwd_bits *bs, rc_error *e) {
    TRY(bd_initialize(&a->first, bs, e)); TRY(bd_initialize(&a->second, bs, e));
    a->last_updated_is_first = 0; a->last_read_is_first = 0; return 0;
}
static size_t alt_expected_bits(const altdec *a) { return a->last_updated_is_first ? bd_expected_bits(&a->second) : bd_expected_bits(&a->first); }
static int alt_symbol(altdec *a, uint16_t *o, rc_error *e) {
    if (a->last_read_is_first) { a->last_read_is_first = 0; return bd_symbol(&a->second, o, e); }
    a->last_read_is_first = 1; return bd_symbol(&a->first, o, e);
}
static int alt_update_bits(altdec *a, rc_bwd_bits *bs, rc_error *e) {
    if (a->last_updated_is_first) { a->last_updated_is_first = 0; return bd_update_bits(&a->second, bs, e); }
    a->last_updated_is_first = 1; return bd_update_bits(&a->first, bs, e);
}

int rc_parse_fse_table(const uint8_t *data, size_t n, uint8_t *al, int16_t *dist, size_t *ndist,
                       size_t *bits_left, size_t *bytes_read, rc_error *e) {
    rc_fwd_bits f; e->code = 0;
    TRY(fwd_init(&f, data, n, e));
    int rc = parse_fse_table(&f, al, dist, ndist, e);
    *bits_left = f.readable; *bytes_read = fwd_bytes_read(&f); return rc;
}
int rc_fse_from_distribution(uint8_t al, const int16_t *dist, size_t ndist, uint16_t *out, rc_error *e) {
    fse_table t; e->code = 0;
    TRY(fse_from_distribution(al, dist, ndist, &t, e));
    for (size_t i = 0; i < t.size; i++) { out[3 * i] = t.t[i].output; out[3 * i + 1] = t.t[i].baseline; out[3 * i + 2] = t.t[i].bits_to_read; }
    free(t.t); return 0;
}
int rc_fse_run(const uint16_t *table, uint8_t al, int alternating, const uint8_t *stream, size_t n,
               size_t count, uint16_t *out, size_t *nout, size_t *bits_left, rc_error *e) {
    fse_table t; t.al = al; t.size = (size_t)1 << al; t.t = (fse_state *)malloc(t.size * sizeof(fse_state));
    for (size_t i = 0; i < t.size; i++) { t.t[i].output = table[3 * i]; t.t[i].baseline = table[3 * i + 1]; t.t[i].bits_to_read = table[3 * i + 2]; }
    rc_bwd_bits bs; e->code = 0; *nout = 0; *bits_left = 0;
    int rc = bwd_init(&bs, stream, n, e);
    if (rc) { free(t.t); return rc; }
    bitdec d; altdec a; memset(&d, 0, sizeof d); memset(&a, 0, sizeof a);
    d.table = t; a.first.table = t; a.second.table = t;
    rc = alternating ? alt_initialize(&a, &bs, e) : bd_initialize(&d, &bs, e);
    for (size_t i = 0; !rc && i < count; i++) {
        rc = alternating ? alt_symbol(&a, &out[i], e) : bd_symbol(&d, &out[i], e);
        if (rc) break;
        *nout = i + 1;
        rc = alternating ? alt_update_bits(&a, &bs, e) : bd_update_bits(&d, &bs, e);
    }
    *bits_left = bs.readable; bwd_drop(&bs); free(t.t); return rc;
}

/* ===== decoders/huffman.rs ============================================= */
/* enum HuffmanDecoder { Absent, Symbol{payload}, Tree{left,right} } huffman.rs:12-22 ; nodes in an arena */
enum { HN_ABSENT = 0, HN_SYMBOL = 1, HN_TREE = 2 };
typedef struct { uint8_t kind, payload; int32_t left, right; } hnode;
typedef struct { hnode *n; int32_t count, cap; } htree;   /* node 0 is the root */
static int32_t ht_new(htree *t) {
    if (t->count == t->cap) { t->cap = t->cap ? t->cap * 2 : 64; t->n = (hnode *)realloc(t->n, (size_t)t->cap * sizeof(hnode)); }
    hnode *x = &t->n[t->count]; x->kind = HN_ABSENT; x->payload = 0; x->left = x->right = -1; return t->count++;
}
static void ht_free(htree *t) { free(t->n); t->n = NULL; t->count = t->cap = 0; }
/* insert huffman.rs:132-159. returns 1 true / 0 false / -1 panic("Trying to inster a symbol into another") */
static int ht_insert(htree *t, int32_t node, uint8_t symbol, unsigned width) {
    if (width == 0) {
        if (t->n[node].kind == HN_ABSENT) { t->n[node].kind = HN_SYMBOL; t->n[node].payload = symbol; return 1; }
        return 0;
    }
    if (t->n[node].kind == HN_ABSENT) {
        int32_t l = ht_new(t), r = ht_new(t);
        t->n[node].kind = HN_TREE; t->n[node].left = l; t->n[node].right = r;
    }
    if (t->n[node].kind == HN_SYMBOL) return -1;
    int a = ht_insert(t, t->n[node].left, symbol, width - 1);
    if (a != 0) return a;
    return ht_insert(t, t->n[node].right, symbol, width - 1);
}
/* from_number_of_bits huffman.rs:161-175: width desc, then symbol asc; leftmost-first insertion */
static int ht_from_number_of_bits(const uint8_t *widths, size_t n, htree *t, rc_error *e) {
    t->n = NULL; t->count = t->cap = 0; ht_new(t);
    unsigned maxw = 0; for (size_t i = 0; i < n; i++) if (widths[i] > maxw) maxw = widths[i];
    for (unsigned w = maxw; w >= 1; w--)
        for (size_t i = 0; i < n; i++)
            if (widths[i] == w) { if (ht_insert(t, 0, (uint8_t)i, w) < 0) { ht_free(t); return fail(e, RC_Panic, 30, 0); } }
    return 0;
}
/* from_weights huffman.rs:177-203 (u8 / u32 arithmetic; debug-build overflow checks => panic) */
static int ht_from_weights(const uint8_t *weights, size_t n, htree *t, uint8_t *widths_out, size_t *nwidths, rc_error *e) {
    uint32_t sum = 0;
    for (size_t i = 0; i < n; i++) if (weights[i] != 0) {
        if (weights[i] - 1 >= 32) return fail(e, RC_Panic, 31, 0);          /* 1 << (poid-1) overflow */
        uint64_t s2 = (uint64_t)sum + ((uint64_t)1 << (weights[i] - 1));
        if (s2 > 0xFFFFFFFFull) return fail(e, RC_Panic, 31, 0);
        sum = (uint32_t)s2;
    }
    int p = discrete_log2(sum);
    if (p < 0) return fail(e, RC_Panic, 32, 0);                               /* huffman.rs:184 assert */
    unsigned puissance = (unsigned)p;
    if (((uint32_t)1 << puissance) < sum) puissance += 1;
    uint8_t rest = (uint8_t)((((uint32_t)1 << puissance) - sum) & 0xFF);      /* `as u8` truncation, huffman.rs:190 */
    int m = discrete_log2(rest);
    if (m < 0) return fail(e, RC_Panic, 33, 0);
    unsigned manquant = (unsigned)m + 1;
    uint8_t widths[MAX_SYMBOL + 1]; size_t k = 0;
    if (n > MAX_SYMBOL) return fail(e, RC_Panic, 34, 0);
    for (size_t i = 0; i < n; i++) {
        if (weights[i] != 0) {
            if (weights[i] > puissance + 1) return fail(e, RC_Panic, 35, 0);  /* u8 underflow */
            widths[k++] = (uint8_t)(puissance + 1 - weights[i]);
        } else widths[k++] = 0;
    }
    widths[k++] = (uint8_t)(puissance + 1 - manquant);
    if (widths_out) { memcpy(widths_out, widths, k); *nwidths = k; }
    /* symbols are `i as u8` (huffman.rs:165): index 256 would wrap; n<=255 weights keeps k<=256 */
    return ht_from_number_of_bits(widths, k, t, e);
}
/* parse_direct huffman.rs:92-106 */
static int huf_parse_direct(bytep *in, size_t num_weights, uint8_t *w, size_t *nw, rc_error *e) {
    const uint8_t *d; size_t nb = num_weights / 2 + num_weights % 2;
    TRY(bp_slice(in, nb, &d, e));
    size_t k = 0;
    for (size_t i = 0; i < nb; i++) { uint8_t lo = d[i] & 0xF, hi = d[i] >> 4; w[k++] = hi; w[k++] = lo; } /* take(4)=low nibble pushed second */
    *nw = num_weights; return 0;   /* truncate */
}
/* parse_fse huffman.rs:108-130 */
static int huf_parse_fse(bytep *in, uint8_t compressed_size, uint8_t *w, size_t *nw, rc_error *e) {
    const uint8_t *d; TRY(bp_slice(in, compressed_size, &d, e));
    rc_fwd_bits fp; if (fwd_init(&fp, d, compressed_size, e)) return fail(e, RC_Panic, 36, 0);
    fse_table tab; TRY(fse_table_parse(&fp, &tab, e));
    size_t br = fwd_bytes_read(&fp);
    rc_bwd_bits bs; int rc = bwd_init(&bs, d + br, compressed_size - br, e);
    if (rc) { free(tab.t); return rc; }
    altdec a; memset(&a, 0, sizeof a); a.first.table = tab; a.second.table = tab;
    size_t k = 0; uint16_t s;
    rc = alt_initialize(&a, &bs, e);
    while (!rc && alt_expected_bits(&a) <= bs.readable) {
        rc = alt_symbol(&a, &s, e); if (rc) break;
        if (k >= 4096) { rc = fail(e, RC_Panic, 37, 0); break; }
        w[k++] = (uint8_t)s;
        rc = alt_update_bits(&a, &bs, e);
    }
    if (!rc) { rc = alt_symbol(&a, &s, e); if (!rc) w[k++] = (uint8_t)s; }
    if (!rc) { rc = alt_symbol(&a, &s, e); if (!rc) w[k++] = (uint8_t)s; }
    bwd_drop(&bs); free(tab.t);
    *nw = k; return rc;
}
/* HuffmanDecoder::parse huffman.rs:80-90 */
static int huf_parse(bytep *in, htree *t, uint8_t *weights_out, size_t *nweights, rc_error *e) {
    uint8_t header; TRY(bp_u8(in, &header, e));
    uint8_t w[4100]; size_t nw = 0;
    if (header < 128) TRY(huf_parse_fse(in, header, w, &nw, e));
    else TRY(huf_parse_direct(in, (size_t)header - 127, w, &nw, e));
    if (weights_out) { memcpy(weights_out, w, nw < 4096 ? nw : 4096); *nweights = nw; }
    return ht_from_weights(w, nw, t, NULL, NULL, e);
}
/* HuffmanDecoder::decode huffman.rs:205-218: one take(1) per tree level */
static int ht_decode(const htree *t, rc_bwd_bits *bs, uint8_t *o, rc_error *e) {
    int32_t node = 0;
    for (;;) {
        const hnode *x = &t->n[node];
        if (x->kind == HN_SYMBOL) { *o = x->payload; return 0; }
        if (x->kind == HN_ABSENT) return fail(e, RC_Panic, 38, 0);
        uint64_t bit; TRY(bwd_take(bs, 1, &bit, e));
        node = bit ? x->right : x->left;
    }
}
static void ht_collect(const htree *t, int32_t node, unsigned depth, uint32_t code, uint8_t *lens, uint32_t *codes) {
    const hnode *x = &t->n[node];
    if (x->kind == HN_SYMBOL) { lens[x->payload] = (uint8_t)depth; codes[x->payload] = code; return; }
    if (x->kind == HN_TREE) { ht_collect(t, x->left, depth + 1, code << 1, lens, codes); ht_collect(t, x->right, depth + 1, (code << 1) | 1, lens, codes); }
}
static int ht_from_lens_codes(const uint8_t *lens, const uint32_t *codes, htree *t) {
    t->n = NULL; t->count = t->cap = 0; ht_new(t);
    for (int s = 0; s < 256; s++) {
        if (!lens[s]) continue;
        int32_t node = 0;
        for (int b = lens[s] - 1; b >= 0; b--) {
            if (t->n[node].kind == HN_ABSENT) { int32_t l = ht_new(t), r = ht_new(t); t->n[node].kind = HN_TREE; t->n[node].left = l; t->n[node].right = r; }
            node = ((codes[s] >> b) & 1) ? t->n[node].right : t->n[node].left;
        }
        t->n[node].kind = HN_SYMBOL; t->n[node].payload = (uint8_t)s;
    }
    return 0;
}
int rc_huffman_from_weights(const uint8_t *weights, size_t n, uint8_t *lens, uint32_t *codes, rc_error *e) {
    htree t; e->code = 0; memset(lens, 0, 257); memset(codes, 0, 257 * sizeof(uint32_t));
    TRY(ht_from_weights(weights, n, &t, NULL, NULL, e));
    ht_collect(&t, 0, 0, 0, lens, codes); ht_free(&t); return 0;
}
int rc_huffman_parse(const uint8_t *data, size_t n, uint8_t *lens, uint32_t *codes, size_t *consumed,
                     uint8_t *weights_out, size_t *nweights, rc_error *e) {
    bytep in = { data, n }; htree t; e->code = 0; memset(lens, 0, 257); memset(codes, 0, 257 * sizeof(uint32_t));
    TRY(huf_parse(&in, &t, weights_out, nweights, e));
    ht_collect(&t, 0, 0, 0, lens, codes); ht_free(&t); *consumed = n - in.n; return 0;
}
int rc_huffman_decode_stream(const uint8_t *lens, const uint32_t *codes, const uint8_t *stream, size_t n,
                             uint8_t *out, size_t out_cap, size_t *out_len, rc_error *e) {
    htree t; e->code = 0; ht_from_lens_codes(lens, codes, &t);
    rc_bwd_bits bs; int rc = bwd_init(&bs, stream, n, e); size_t k = 0;
    if (!rc) {
        while (bs.readable != 0) { uint8_t s; rc = ht_decode(&t, &bs, &s, e); if (rc) break; if (k < out_cap) out[k] = s; k++; }
        bwd_drop(&bs);
    }
    ht_free(&t); *out_len = k; return rc;
}

/* ===== sequences.rs / decoders/sequence.rs ============================= */
enum { M_PREDEFINED = 0, M_RLE = 1, M_FSE = 2, M_REPEAT = 3 };       /* sequences.rs:240-256 */
typedef struct { int mode; uint8_t rle; fse_table table; int has_table; } symmode;
static void symmode_drop(symmode *m) { if (m->has_table) { free(m->table.t); m->has_table = 0; } }
static symmode symmode_clone(const symmode *m) { symmode c = *m; if (m->has_table) c.table = fse_clone(&m->table); return c; }

static const int16_t LITERALS_LENGTH_DISTRI[36] = {  /* sequences.rs:29-32 */
    4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1 };
static const int16_t OFFSET_DISTRI[29] = {           /* sequences.rs:33-35 */
    1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1 };
static const int16_t MATCH_LENGTH_DISTRI[53] = {     /* sequences.rs:36-39 */
    1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1 };
/* sequence.rs:95-191 (RFC 8878 tables 3.1.1.3.2.1.1) as (code, baseline, nb_bits) searched linearly */
typedef struct { uint16_t code; size_t baseline, nb_bits; } codeval;
static const codeval ML_CODE_TO_VALUE[53] = {
    {0,3,0},{1,4,0},{2,5,0},{3,6,0},{4,7,0},{5,8,0},{6,9,0},{7,10,0},{8,11,0},{9,12,0},{10,13,0},{11,14,0},{12,15,0},
    {13,16,0},{14,17,0},{15,18,0},{16,19,0},{17,20,0},{18,21,0},{19,22,0},{20,23,0},{21,24,0},{22,25,0},{23,26,0},
    {24,27,0},{25,28,0},{26,29,0},{27,30,0},{28,31,0},{29,32,0},{30,33,0},{31,34,0},{32,35,1},{33,37,1},{34,39,1},
    {35,41,1},{36,43,2},{37,47,2},{38,51,3},{39,59,3},{40,67,4},{41,83,4},{42,99,5},{43,131,7},{44,259,8},{45,515,9},
    {46,1027,10},{47,2051,11},{48,4099,12},{49,8195,13},{50,16387,14},{51,32771,15},{52,65539,16} };
static const codeval LL_CODE_TO_VALUE[36] = {
    {0,0,0},{1,1,0},{2,2,0},{3,3,0},{4,4,0},{5,5,0},{6,6,0},{7,7,0},{8,8,0},{9,9,0},{10,10,0},{11,11,0},{12,12,0},
    {13,13,0},{14,14,0},{15,15,0},{16,16,1},{17,18,1},{18,20,1},{19,22,1},{20,24,2},{21,28,2},{22,32,3},{23,40,3},
    {24,48,4},{25,64,6},{26,128,7},{27,256,8},{28,512,9},{29,1024,10},{30,2048,11},{31,4096,12},{32,8192,13},
    {33,16384,14},{34,32768,15},{35,65536,16} };
#define MAX_OFFSET_CODE 31
#define MAX_LL_CODE 35
#define MAX_ML_CODE 52

typedef struct { size_t ll, off, ml; } seq3;
typedef struct {
    size_t number_of_sequences; symmode ll_mode, of_mode, ml_mode; const uint8_t *bitstream; size_t bitstream_len;
} sequences;  /* sequences.rs:41-48 */

/* parse_num_sequences sequences.rs:77-87 (quirks: +0x7F ; rfc: +0x7F00) */
static int parse_num_sequences(bytep *in, int quirks, size_t *o, rc_error *e) {
    uint8_t b0, b1, b2; TRY(bp_u8(in, &b0, e));
    if (b0 < 128) { *o = b0; return 0; }
    if (b0 < 255) { TRY(bp_u8(in, &b1, e)); *o = ((size_t)(b0 - 128) << 8) + b1; return 0; }
    TRY(bp_u8(in, &b1, e)); TRY(bp_u8(in, &b2, e));
    *o = (size_t)b1 + ((size_t)b2 << 8) + (quirks ? 0x7F : 0x7F00); return 0;
}
/* parse_symbol_compression sequences.rs:91-143 */
static int parse_symbol_compression(bytep *in, symmode out[3], rc_error *e) {
    const uint8_t *hb; TRY(bp_slice(in, 1, &hb, e));
    uint8_t h = *hb;
    if ((h & 3) != 0) return fail(e, RC_SeqReservedSet, 0, 0);
    int tmp[3]; tmp[2] = (h >> 2) & 3; tmp[1] = (h >> 4) & 3; tmp[0] = (h >> 6) & 3;
    for (int i = 0; i < 3; i++) { out[i].mode = M_REPEAT; out[i].has_table = 0; out[i].rle = 0; }
    for (int i = 0; i < 3; i++) {
        switch (tmp[i]) {
        case 3: out[i].mode = M_REPEAT; break;
        case 0: out[i].mode = M_PREDEFINED; break;
        case 1: { uint8_t b; int rc = bp_u8(in, &b, e); if (rc) { for (int j = 0; j < i; j++) symmode_drop(&out[j]); return rc; } out[i].mode = M_RLE; out[i].rle = b; break; }
        case 2: {
            const uint8_t *nd; size_t nlen = in->n; int rc = bp_slice(in, nlen, &nd, e);
            rc_fwd_bits fp;
            if (!rc) rc = fwd_init(&fp, nd, nlen, e);
            if (!rc) rc = fse_table_parse(&fp, &out[i].table, e);
            if (rc) { for (int j = 0; j < i; j++) symmode_drop(&out[j]); return rc; }
            out[i].mode = M_FSE; out[i].has_table = 1;
            size_t br = fwd_bytes_read(&fp); in->p = nd + br; in->n = nlen - br; break;
        } }
    }
    return 0;
}
/* Sequences::parse sequences.rs:52-75 */
static int sequences_parse(bytep *in, int quirks, sequences *s, rc_error *e) {
    memset(s, 0, sizeof *s);
    s->ll_mode.mode = s->of_mode.mode = s->ml_mode.mode = M_REPEAT;
    TRY(parse_num_sequences(in, quirks, &s->number_of_sequences, e));
    if (s->number_of_sequences == 0) return 0;
    symmode m[3]; TRY(parse_symbol_compression(in, m, e));
    s->ll_mode = m[0]; s->of_mode = m[1]; s->ml_mode = m[2];
    int rc = bp_slice(in, in->n, &s->bitstream, e);
    if (rc) { symmode_drop(&s->ll_mode); symmode_drop(&s->of_mode); symmode_drop(&s->ml_mode); return rc; }
    s->bitstream_len = (size_t)(in->p - s->bitstream); return 0;
}
static void sequences_drop(sequences *s) { symmode_drop(&s->ll_mode); symmode_drop(&s->of_mode); symmode_drop(&s->ml_mode); }

/* ===== decoding_context.rs ============================================= */
typedef struct {
    htree huffman; int has_huffman; vec8 decoded; size_t offsets[3]; uint64_t window_size;
    symmode ll_repeat, cmov_repeat, ml_repeat; int has_ll, has_cmov, has_ml;
} dctx;   /* decoding_context.rs:17-26 */
static int dctx_new(dctx *c, uint64_t window, rc_error *e) {            /* :29-47 */
    if (window > MAX_WIN_SIZE) return fail(e, RC_WindowSizeTooBig, MAX_WIN_SIZE, window);
    memset(c, 0, sizeof *c); c->offsets[0] = 1; c->offsets[1] = 4; c->offsets[2] = 8; c->window_size = window; return 0;
}
static void dctx_drop(dctx *c, int keep_decoded) {
    if (c->has_huffman) ht_free(&c->huffman);
    if (c->has_ll) symmode_drop(&c->ll_repeat);
    if (c->has_cmov) symmode_drop(&c->cmov_repeat);
    if (c->has_ml) symmode_drop(&c->ml_repeat);
    if (!keep_decoded) free(c->decoded.p);
}
/* decode_offset decoding_context.rs:50-75 */
static int decode_offset(dctx *c, size_t offset, size_t ll, size_t *o, rc_error *e) {
    size_t *f = c->offsets;
    if (offset == 0) return fail(e, RC_NullOffsetError, 0, 0);
    else if (offset == 3 && ll == 0) { f[2] = f[1]; f[1] = f[0]; if (f[0] == 0) return fail(e, RC_Panic, 40, 0); f[0] -= 1; }
    else if (offset == 3 || (offset == 2 && ll == 0)) { size_t t = f[2]; f[2] = f[1]; f[1] = f[0]; f[0] = t; }
    else if (offset == 2 || (offset == 1 && ll == 0)) { size_t t = f[0]; f[0] = f[1]; f[1] = t; }
    else if (offset == 1) { }
    else { f[2] = f[1]; f[1] = f[0]; f[0] = offset - 3; }
    *o = f[0]; return 0;
}
/* execute_sequences decoding_context.rs:78-106 */
static int execute_sequences(dctx *c, const seq3 *seqs, size_t nseq, const uint8_t *lits, size_t nlits, rc_error *e) {
    for (size_t i = 0; i < nseq; i++) {
        size_t ll = seqs[i].ll, ml = seqs[i].ml, off;
        TRY(decode_offset(c, seqs[i].off, ll, &off, e));
        if (ll > nlits || off > c->decoded.len + ll) return fail(e, RC_ImpossibleValue, 0, 0);
        v8_extend(&c->decoded, lits, ll); lits += ll; nlits -= ll;
        if (ml && off == 0) return fail(e, RC_Panic, 41, 0);           /* decoded[len - 0] index panic */
        for (size_t k = 0; k < ml; k++) v8_push(&c->decoded, c->decoded.p[c->decoded.len - off]);
    }
    for (size_t k = 0; k < nlits; k++) v8_push(&c->decoded, lits[k]);
    return 0;
}
int rc_execute_sequences(uint64_t window, const uint64_t *seqs, size_t nseq, const uint8_t *lits, size_t nlits,
                         uint8_t **out, size_t *out_len, rc_error *e) {
    dctx c; e->code = 0; *out = NULL; *out_len = 0;
    TRY(dctx_new(&c, window, e));
    seq3 *s = (seq3 *)malloc((nseq ? nseq : 1) * sizeof(seq3));
    for (size_t i = 0; i < nseq; i++) { s[i].ll = (size_t)seqs[3 * i]; s[i].off = (size_t)seqs[3 * i + 1]; s[i].ml = (size_t)seqs[3 * i + 2]; }
    int rc = execute_sequences(&c, s, nseq, lits, nlits, e);
    free(s);
    *out = c.decoded.p; *out_len = c.decoded.len; dctx_drop(&c, 1); return rc;
}

/* get_decoder sequences.rs:147-187.  `depth` models the `&None` recursion guard. */
static int get_decoder(int code_type, const symmode *mode, const symmode *prev, int has_prev, bitdec *dec, symmode *used, rc_error *e) {
    memset(dec, 0, sizeof *dec);
    switch (mode->mode) {
    case M_RLE: dec->kind = 1; dec->byte = mode->rle; *used = symmode_clone(mode); return 0;
    case M_FSE: dec->kind = 0; dec->table = fse_clone(&mode->table); *used = symmode_clone(mode); return 0;
    case M_REPEAT:
        if (!has_prev || prev->mode == M_REPEAT) return fail(e, RC_NoPreviousDecoder, 0, 0);
        return get_decoder(code_type, prev, NULL, 0, dec, used, e);
    default: {
        fse_table t; int rc;
        if (code_type == 0) rc = fse_from_distribution(6, LITERALS_LENGTH_DISTRI, 36, &t, e);
        else if (code_type == 1) rc = fse_from_distribution(5, OFFSET_DISTRI, 29, &t, e);
        else rc = fse_from_distribution(6, MATCH_LENGTH_DISTRI, 53, &t, e);
        if (rc) return rc;
        dec->kind = 0; dec->table = t; memset(used, 0, sizeof *used); used->mode = M_PREDEFINED; return 0;
    } }
}
static void bitdec_drop(bitdec *d) { if (d->kind == 0 && d->table.t) { free(d->table.t); d->table.t = NULL; } }

static int get_value(uint16_t code, const codeval *tab, size_t n, rc_bwd_bits *bs, size_t *o, rc_error *e) { /* sequence.rs:30-39 */
    const codeval *f = NULL;
    for (size_t i = 0; i < n; i++) if (tab[i].code == code) { f = &tab[i]; break; }   /* linear find */
    if (!f) return fail(e, RC_Panic, 50, 0);
    uint64_t v; TRY(bwd_take(bs, f->nb_bits, &v, e)); *o = (size_t)v + f->baseline; return 0;
}
/* update_symbol_value sequence.rs:41-55 */
static int update_symbol_value(bitdec *ll, bitdec *of, bitdec *ml, rc_bwd_bits *bs, seq3 *o, rc_error *e) {
    uint16_t oc, lc, mc;
    TRY(bd_symbol(of, &oc, e)); TRY(bd_symbol(ll, &lc, e)); TRY(bd_symbol(ml, &mc, e));
    if (lc > MAX_LL_CODE || mc > MAX_ML_CODE || oc > MAX_OFFSET_CODE) return fail(e, RC_SequenceCodeMaxValueExceeded, 0, 0);
    uint64_t v; TRY(bwd_take(bs, oc, &v, e));
    o->off = ((size_t)1 << oc) + (size_t)v;
    TRY(get_value(mc, ML_CODE_TO_VALUE, 53, bs, &o->ml, e));
    TRY(get_value(lc, LL_CODE_TO_VALUE, 36, bs, &o->ll, e));
    return 0;
}
/* Sequences::decode sequences.rs:191-237 */
static int sequences_decode(sequences *s, dctx *c, seq3 **out, size_t *nout, rc_error *e) {
    bitdec ll, of, ml; symmode nll, nof, nml; int rc; *out = NULL; *nout = 0;
    rc = get_decoder(0, &s->ll_mode, &c->ll_repeat, c->has_ll, &ll, &nll, e); if (rc) return rc;
    rc = get_decoder(1, &s->of_mode, &c->cmov_repeat, c->has_cmov, &of, &nof, e); if (rc) { bitdec_drop(&ll); symmode_drop(&nll); return rc; }
    rc = get_decoder(2, &s->ml_mode, &c->ml_repeat, c->has_ml, &ml, &nml, e);
    if (rc) { bitdec_drop(&ll); bitdec_drop(&of); symmode_drop(&nll); symmode_drop(&nof); return rc; }
    rc_bwd_bits bs; seq3 *res = NULL; size_t n = 0;
    rc = bwd_init(&bs, s->bitstream, s->bitstream_len, e);
    if (!rc) {
        rc = bd_initialize(&ll, &bs, e);                                   /* sequence.rs:59-65 LL, OF, ML */
        if (!rc) rc = bd_initialize(&of, &bs, e);
        if (!rc) rc = bd_initialize(&ml, &bs, e);
        res = (seq3 *)malloc((s->number_of_sequences + 1) * sizeof(seq3));
        if (!rc && s->number_of_sequences == 0) rc = fail(e, RC_Panic, 51, 0);  /* usize underflow :217 */
        for (size_t i = 0; !rc && i + 1 < s->number_of_sequences; i++) {
            rc = update_symbol_value(&ll, &of, &ml, &bs, &res[n], e); if (rc) break; n++;
            rc = bd_update_bits(&ll, &bs, e);                               /* sequence.rs:80-88 LL, ML, OF */
            if (!rc) rc = bd_update_bits(&ml, &bs, e);
            if (!rc) rc = bd_update_bits(&of, &bs, e);
        }
        if (!rc) { rc = update_symbol_value(&ll, &of, &ml, &bs, &res[n], e); if (!rc) n++; }
        bwd_drop(&bs);
    }
    bitdec_drop(&ll); bitdec_drop(&of); bitdec_drop(&ml);
    if (rc) { free(res); symmode_drop(&nll); symmode_drop(&nof); symmode_drop(&nml); return rc; }
    if (c->has_cmov) symmode_drop(&c->cmov_repeat);
    if (c->has_ll) symmode_drop(&c->ll_repeat);
    if (c->has_ml) symmode_drop(&c->ml_repeat);
    c->cmov_repeat = nof; c->ll_repeat = nll; c->ml_repeat = nml; c->has_cmov = c->has_ll = c->has_ml = 1;  /* :232-234 */
    *out = res; *nout = n; return 0;
}

/* ===== literals.rs ===================================================== */
enum { LT_RAW = 0, LT_RLE = 1, LT_COMPRESSED = 2, LT_TREELESS = 3 };
typedef struct {
    int kind;                               /* 0 raw, 1 rle, 2 compressed(+treeless) */
    const uint8_t *data; size_t data_len; uint8_t byte; uint32_t repeat;
    htree tree; int has_tree; size_t regenerated_size; uint16_t jump_table[4];
} literals;   /* literals.rs:22-36 */
/* parse_header literals.rs:135-206 */
static int lit_parse_header(bytep *in, int *type, size_t *regen, size_t *csize, int *n_streams, rc_error *e) {
    uint8_t header; TRY(bp_u8(in, &header, e));
    int lt = header & 3, sf = (header >> 2) & 3; uint8_t b1, b2; const uint8_t *s;
    *csize = 0; *n_streams = 1;
    if (lt == 0 || lt == 1) {
        if (sf == 0 || sf == 2) *regen = header >> 3;
        else if (sf == 1) { TRY(bp_u8(in, &b1, e)); *regen = (size_t)(header >> 4) + ((size_t)b1 << 4); }
        else { TRY(bp_u8(in, &b1, e)); TRY(bp_u8(in, &b2, e)); *regen = (size_t)(header >> 4) + ((size_t)b1 << 4) + ((size_t)b2 << 12); }
    } else {
        size_t extra = sf == 0 || sf == 1 ? 2 : sf == 2 ? 3 : 4;
        TRY(bp_slice(in, extra, &s, e));
        uint64_t v = 0; for (size_t i = 0; i < extra; i++) v |= (uint64_t)s[i] << (8 * i);
        unsigned rb = sf <= 1 ? 6 : sf == 2 ? 10 : 14, cb = sf <= 1 ? 10 : sf == 2 ? 14 : 18;
        *regen = (size_t)(header >> 4) + ((size_t)(v & (((uint64_t)1 << rb) - 1)) << 4);
        *csize = (size_t)((v >> rb) & (((uint64_t)1 << cb) - 1));
        *n_streams = sf == 0 ? 1 : 4;
    }
    *type = lt; return 0;
}
/* LiteralsSection::parse literals.rs:88-133 */
static int literals_parse(bytep *in, int quirks, literals *l, rc_error *e) {
    int lt, ns; size_t regen, csize; memset(l, 0, sizeof *l);
    TRY(lit_parse_header(in, &lt, &regen, &csize, &ns, e));
    if (lt == LT_RAW && !quirks && regen == 0) { l->kind = 0; l->data = in->p; l->data_len = 0; return 0; }   /* RFC: legal (Q2) */
    if (lt == LT_RAW) { l->kind = 0; TRY(bp_slice(in, regen, &l->data, e)); l->data_len = regen; return 0; }
    if (lt == LT_RLE) { l->kind = 1; TRY(bp_u8(in, &l->byte, e)); l->repeat = (uint32_t)regen; return 0; }
    const uint8_t *w; TRY(bp_slice(in, csize, &w, e));
    bytep ni = { w, csize };
    l->kind = 2; l->regenerated_size = regen;
    if (lt != LT_TREELESS) { TRY(huf_parse(&ni, &l->tree, NULL, NULL, e)); l->has_tree = 1; }
    size_t total = ni.n; int rc = 0;
    if (ns == 4) {
        uint16_t s1, s2, s3;
        rc = bp_le_u16(&ni, &s1, e); if (!rc) rc = bp_le_u16(&ni, &s2, e); if (!rc) rc = bp_le_u16(&ni, &s3, e);
        if (!rc && (size_t)s1 + s2 + s3 > total - 6) rc = fail(e, RC_CorruptedStreamsSizeTooBig, 0, 0);
        if (!rc) { size_t s4 = total - 6 - s1 - s2 - s3; l->jump_table[0] = s1; l->jump_table[1] = s2; l->jump_table[2] = s3; l->jump_table[3] = (uint16_t)s4; }
    } else { l->jump_table[0] = (uint16_t)ni.n; }
    if (!rc) { rc = bp_slice(&ni, ni.n, &l->data, e); l->data_len = (size_t)(ni.p - l->data); }
    if (rc && l->has_tree) { ht_free(&l->tree); l->has_tree = 0; }
    return rc;
}
static void literals_drop(literals *l) { if (l->has_tree) { ht_free(&l->tree); l->has_tree = 0; } }
/* LiteralsSection::decode literals.rs:49-86 */
static int literals_decode(literals *l, dctx *c, vec8 *res, rc_error *e) {
    if (l->kind == 0) { v8_extend(res, l->data, l->data_len); return 0; }
    if (l->kind == 1) { v8_reserve(res, l->repeat); memset(res->p + res->len, l->byte, l->repeat); res->len += l->repeat; return 0; }
    if (l->has_tree) { if (c->has_huffman) ht_free(&c->huffman); c->huffman = l->tree; c->has_huffman = 1; l->has_tree = 0; }
    if (!c->has_huffman) return fail(e, RC_HuffmanDecoderMissing, 0, 0);
    const uint8_t *data = l->data; size_t left = l->data_len;
    for (int k = 0; k < 4; k++) {
        size_t ss = l->jump_table[k];
        if (ss == 0) break;
        if (ss > left) return fail(e, RC_Panic, 60, 0);
        rc_bwd_bits bs; TRY(bwd_init(&bs, data, ss, e));
        data += ss; left -= ss;
        while (bs.readable != 0) { uint8_t s; int rc = ht_decode(&c->huffman, &bs, &s, e); if (rc) { bwd_drop(&bs); return rc; } v8_push(res, s); }
        bwd_drop(&bs);
    }
    return 0;
}

/* ===== block.rs ======================================================== */
typedef struct { int kind; const uint8_t *raw; size_t raw_len; uint8_t byte; uint32_t repeat; literals lit; sequences seq; } block;
/* Block::parse block.rs:43-72 */
static int block_parse(bytep *in, int quirks, block *b, int *last, rc_error *e) {
    const uint8_t *h; memset(b, 0, sizeof *b);
    TRY(bp_slice(in, 3, &h, e));
    uint32_t v = (uint32_t)h[0] | (uint32_t)h[1] << 8 | (uint32_t)h[2] << 16;
    *last = v & 1; int type = (v >> 1) & 3; size_t size = v >> 3;
    switch (type) {
    case 0:
        b->kind = 0;
        if (!quirks && size == 0) { b->raw = in->p; b->raw_len = 0; return 0; }   /* RFC: empty raw block is legal (Q2) */
        TRY(bp_slice(in, size, &b->raw, e)); b->raw_len = size; return 0;
    case 1: b->kind = 1; TRY(bp_u8(in, &b->byte, e)); b->repeat = (uint32_t)size; return 0;
    case 2: {
        const uint8_t *p; TRY(bp_slice(in, size, &p, e));
        bytep ni = { p, size }; b->kind = 2;
        TRY(literals_parse(&ni, quirks, &b->lit, e));
        int rc = sequences_parse(&ni, quirks, &b->seq, e);
        if (rc) literals_drop(&b->lit);
        return rc; }
    default: return fail(e, RC_ReservedBlockType, 0, 0);
    }
}
static void block_drop(block *b) { if (b->kind == 2) { literals_drop(&b->lit); sequences_drop(&b->seq); } }
/* Block::decode block.rs:74-99 */
static int block_decode(block *b, dctx *c, rc_error *e) {
    if (b->kind == 0) { v8_extend(&c->decoded, b->raw, b->raw_len); return 0; }
    if (b->kind == 1) { v8_reserve(&c->decoded, b->repeat); memset(c->decoded.p + c->decoded.len, b->byte, b->repeat); c->decoded.len += b->repeat; return 0; }
    vec8 lits = { 0, 0, 0 }; int rc = literals_decode(&b->lit, c, &lits, e);
    seq3 *s = NULL; size_t ns = 0;
    if (!rc) rc = sequences_decode(&b->seq, c, &s, &ns, e);
    if (!rc) rc = execute_sequences(c, s, ns, lits.p, lits.len, e);
    free(s); free(lits.p); return rc;
}

/* ===== XXH64 (published algorithm; twox-hash 1.6.3 is not in the tree) === */
#define P1 0x9E3779B185EBCA87ull
#define P2 0xC2B2AE3D27D4EB4Full
#define P3 0x165667B19E3779F9ull
#define P4 0x85EBCA77C2B2AE63ull
#define P5 0x27D4EB2F165667C5ull
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t xround(uint64_t acc, uint64_t in) { return rotl64(acc + in * P2, 31) * P1; }
static uint64_t xmerge(uint64_t h, uint64_t v) { return (h ^ xround(0, v)) * P1 + P4; }
uint64_t rc_xxh64(const uint8_t *p, size_t n, uint64_t seed) {
    const uint8_t *end = p + n; uint64_t h;
    if (n >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do { v1 = xround(v1, rd64(p)); v2 = xround(v2, rd64(p + 8)); v3 = xround(v3, rd64(p + 16)); v4 = xround(v4, rd64(p + 24)); p += 32; } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else h = seed + P5;
    h += (uint64_t)n;
    while (p + 8 <= end) { h ^= xround(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p) * P5; h = rotl64(h, 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32; return h;
}

/* ===== frame.rs ======================================================== */
#define MAGIC_ZSTD 0xFD2FB528u
#define MAGIC_SKIP 0x184D2A50u
uint64_t rc_window_descriptor(uint8_t b) {                 /* frame.rs:179-187 */
    uint64_t mantissa = b & 7, exponent = b >> 3;
    uint64_t base = (uint64_t)1 << (exponent + 10);
    return base + (base / 8) * mantissa;
}
/* Header::parse frame.rs:111-177 */
static int header_parse(bytep *in, rc_header *h, rc_error *e) {
    const uint8_t *d; memset(h, 0, sizeof *h);
    TRY(bp_slice(in, 1, &d, e));
    uint8_t b = *d;
    unsigned dict_id_flag = b & 3, checksum = (b >> 2) & 1, reserved = (b >> 3) & 1, single = (b >> 5) & 1, csf = b >> 6;
    if (reserved) return fail(e, RC_FrameReservedSet, 0, 0);
    int fcs = (csf == 0 && !single) ? 0 : (csf == 0 ? 1 : 1 << csf);
    uint64_t window = 0; int has_window = 0;
    if (!single) { uint8_t wd; TRY(bp_u8(in, &wd, e)); window = rc_window_descriptor(wd); has_window = 1; }
    if (dict_id_flag) {
        size_t n = (size_t)1 << (dict_id_flag - 1); TRY(bp_slice(in, n, &d, e));
        uint64_t v = 0; for (size_t i = 0; i < n; i++) v |= (uint64_t)d[i] << (8 * i);
        h->has_dict_id = 1; h->dictionnary_id = v;
    }
    if (fcs) {
        TRY(bp_slice(in, (size_t)fcs, &d, e));
        uint64_t v = 0; for (int i = 0; i < fcs; i++) v |= (uint64_t)d[i] << (8 * i);
        if (fcs == 2) v += 256;
        h->has_content_size = 1; h->content_size = v;
    }
    h->content_checksum_flag = (uint8_t)checksum;
    h->window_size = has_window ? window : h->content_size;
    return 0;
}
int rc_header_parse(const uint8_t *data, size_t n, rc_header *h, size_t *consumed, rc_error *e) {
    bytep in = { data, n }; e->code = 0; int rc = header_parse(&in, h, e); *consumed = n - in.n; return rc;
}

typedef struct { rc_header header; block *blocks; size_t nblocks; int has_checksum; uint32_t checksum; } zframe;
static void zframe_drop(zframe *z) { for (size_t i = 0; i < z->nblocks; i++) block_drop(&z->blocks[i]); free(z->blocks); z->blocks = NULL; z->nblocks = 0; }
/* ZStandard::parse frame.rs:198-230 (eager: every block incl. its tables) */
static int zframe_parse(bytep *in, int quirks, zframe *z, rc_error *e) {
    memset(z, 0, sizeof *z);
    TRY(header_parse(in, &z->header, e));
    if (z->header.window_size > MAX_WIN_SIZE) return fail(e, RC_WindowSizeTooBig, MAX_WIN_SIZE, z->header.window_size);
    size_t cap = 0;
    for (;;) {
        if (z->nblocks == cap) { cap = cap ? cap * 2 : 8; z->blocks = (block *)realloc(z->blocks, cap * sizeof(block)); }
        int last; int rc = block_parse(in, quirks, &z->blocks[z->nblocks], &last, e);
        if (rc) { zframe_drop(z); return rc; }
        z->nblocks++;
        if (last) break;
    }
    if (z->header.content_checksum_flag) {
        rc_error pe; int rc = bp_le_u32(in, &z->checksum, &pe);
        if (rc) { zframe_drop(z); return fail(e, RC_MissingChecksum, pe.a, pe.b); }
        z->has_checksum = 1;
    }
    return 0;
}
/* ZStandard::decode frame.rs:232-260.  The reference's checksum comparison is a no-op (it hashes a
 * copy) and the function always returns Ok(decoded); we compute the real XXH64 only to
 * report it. */
static int zframe_decode(zframe *z, int quirks, vec8 *out, uint32_t *xxh_low, rc_error *e) {
    dctx c; TRY(dctx_new(&c, z->header.window_size, e));
    int rc = 0;
    for (size_t i = 0; i < z->nblocks && !rc; i++) {
        block *b = &z->blocks[i];
        if (!quirks && b->kind == 2 && b->seq.number_of_sequences == 0) {   /* RFC: literals only (Q1) */
            vec8 lits = { 0, 0, 0 }; rc = literals_decode(&b->lit, &c, &lits, e);
            if (!rc) v8_extend(&c.decoded, lits.p, lits.len);
            free(lits.p);
        } else rc = block_decode(b, &c, e);
    }
    if (!rc) { *xxh_low = (uint32_t)rc_xxh64(c.decoded.p, c.decoded.len, 0); *out = c.decoded; dctx_drop(&c, 1); }
    else dctx_drop(&c, 0);
    return rc;
}

/* Frame::parse frame.rs:61-77 ; FrameIterator frame.rs:87-100 */
typedef struct { int kind; uint32_t magic; const uint8_t *skip; size_t skip_len; zframe z; } frame;
static int frame_parse(bytep *in, int quirks, frame *f, rc_error *e) {
    uint32_t magic; memset(f, 0, sizeof *f);
    TRY(bp_le_u32(in, &magic, e));
    f->magic = magic;
    if (magic == MAGIC_ZSTD) { f->kind = 0; return zframe_parse(in, quirks, &f->z, e); }
    if ((magic ^ MAGIC_SKIP) <= 0x0F) {
        uint32_t len; TRY(bp_le_u32(in, &len, e));
        f->kind = 1;
        if (!quirks && len == 0) { f->skip = in->p; f->skip_len = 0; return 0; }   /* RFC: empty skippable legal (Q2) */
        TRY(bp_slice(in, len, &f->skip, e)); f->skip_len = len; return 0;
    }
    return fail(e, RC_UnrecognizedMagic, magic, 0);
}

int rc_decode_frames(const uint8_t *src, size_t n, int quirks, uint8_t **out, size_t *out_len,
                     rc_frame_info **frames, size_t *nframes, rc_error *e) {
    bytep in = { src, n }; vec8 res = { 0, 0, 0 }; rc_frame_info *fi = NULL; size_t nf = 0, cap = 0; int rc = 0;
    e->code = 0; e->a = e->b = 0;
    while (in.n != 0) {
        frame f; const uint8_t *start = in.p;
        rc = frame_parse(&in, quirks, &f, e); if (rc) break;
        if (nf == cap) { cap = cap ? cap * 2 : 16; fi = (rc_frame_info *)realloc(fi, cap * sizeof *fi); }
        rc_frame_info *x = &fi[nf]; memset(x, 0, sizeof *x);
        x->kind = (uint32_t)f.kind; x->magic = f.magic; x->src_off = (uint64_t)(start - src); x->src_len = (uint64_t)(in.p - start);
        x->out_off = res.len;
        if (f.kind == 1) { v8_extend(&res, f.skip, f.skip_len); x->out_len = f.skip_len; }
        else {
            vec8 d = { 0, 0, 0 }; uint32_t xl = 0;
            x->n_blocks = (uint32_t)f.z.nblocks; x->has_checksum = (uint8_t)f.z.has_checksum; x->stored_checksum = f.z.checksum; x->header = f.z.header;
            rc = zframe_decode(&f.z, quirks, &d, &xl, e);
            zframe_drop(&f.z);
            if (rc) break;
            v8_extend(&res, d.p, d.len); x->out_len = d.len; x->computed_xxh64_low32 = xl; free(d.p);
        }
        nf++;
    }
    *out = res.p; *out_len = res.len; if (frames) *frames = fi; else free(fi); *nframes = nf; return rc;
}

int rc_main_decode(const uint8_t *src, size_t n, int print_skippable, uint8_t **out, size_t *out_len, rc_error *e) {
    bytep in = { src, n }; vec8 res = { 0, 0, 0 }; int rc = 0; e->code = 0; e->a = e->b = 0;
    while (in.n != 0) {                                        /* main.rs:43-53 */
        frame f; rc = frame_parse(&in, 1, &f, e); if (rc) break;
        if (f.kind == 1) { if (print_skippable) v8_extend(&res, f.skip, f.skip_len); continue; }
        vec8 d = { 0, 0, 0 }; uint32_t xl;
        rc = zframe_decode(&f.z, 1, &d, &xl, e); zframe_drop(&f.z);
        if (rc) break;
        v8_extend(&res, d.p, d.len); free(d.p);
    }
    if (rc) { free(res.p); res.p = NULL; res.len = 0; }        /* main.rs:51 no partial output */
    *out = res.p; *out_len = res.len; return rc;
}

/* Frame-parallel driver for the timed CPU baseline only.  Pass 1 walks frame boundaries (cheap:
 * magic, header, 3-byte block headers); workers then run frame_parse + decode on whole frames. */
typedef struct { const uint8_t *p; size_t n; int kind; vec8 out; int rc; rc_error e; } mt_item;
typedef struct { mt_item *items; size_t n; size_t next; pthread_mutex_t mu; int print_skippable; } mt_pool;
static int frame_extent(bytep *in, rc_error *e) {
    uint32_t magic; TRY(bp_le_u32(in, &magic, e));
    if (magic == MAGIC_ZSTD) {
        rc_header h; TRY(header_parse(in, &h, e));
        for (;;) {
            const uint8_t *hh; TRY(bp_slice(in, 3, &hh, e));
            uint32_t v = (uint32_t)hh[0] | (uint32_t)hh[1] << 8 | (uint32_t)hh[2] << 16;
            size_t sz = ((v >> 1) & 3) == 1 ? 1 : (v >> 3);
            if (((v >> 1) & 3) == 3) return fail(e, RC_ReservedBlockType, 0, 0);
            if (in->n < sz) return fail(e, RC_NotEnoughBytes, sz, in->n);
            in->p += sz; in->n -= sz;
            if (v & 1) break;
        }
        if (h.content_checksum_flag) { uint32_t c; TRY(bp_le_u32(in, &c, e)); }
        return 0;
    }
    if ((magic ^ MAGIC_SKIP) <= 0x0F) { uint32_t len; TRY(bp_le_u32(in, &len, e)); const uint8_t *d; TRY(bp_slice(in, len, &d, e)); return 0; }
    return fail(e, RC_UnrecognizedMagic, magic, 0);
}
static void *mt_worker(void *arg) {
    mt_pool *p = (mt_pool *)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu); size_t i = p->next++; pthread_mutex_unlock(&p->mu);
        if (i >= p->n) return NULL;
        mt_item *it = &p->items[i]; bytep in = { it->p, it->n }; frame f;
        it->rc = frame_parse(&in, 1, &f, &it->e); if (it->rc) continue;
        if (f.kind == 1) { if (p->print_skippable) v8_extend(&it->out, f.skip, f.skip_len); continue; }
        uint32_t xl; it->rc = zframe_decode(&f.z, 1, &it->out, &xl, &it->e); zframe_drop(&f.z);
    }
}
int rc_main_decode_mt(const uint8_t *src, size_t n, int print_skippable, int threads,
                      uint8_t **out, size_t *out_len, rc_error *e) {
    bytep in = { src, n }; mt_pool pool; memset(&pool, 0, sizeof pool); size_t cap = 0; int rc = 0;
    e->code = 0; e->a = e->b = 0; *out = NULL; *out_len = 0;
    while (in.n != 0) {
        const uint8_t *s = in.p; rc = frame_extent(&in, e);
        if (rc) { /* let the serial path produce the exact reference error */ free(pool.items); return rc_main_decode(src, n, print_skippable, out, out_len, e); }
        if (pool.n == cap) { cap = cap ? cap * 2 : 64; pool.items = (mt_item *)realloc(pool.items, cap * sizeof(mt_item)); }
        memset(&pool.items[pool.n], 0, sizeof(mt_item)); pool.items[pool.n].p = s; pool.items[pool.n].n = (size_t)(in.p - s); pool.n++;
    }
    pool.print_skippable = print_skippable; pthread_mutex_init(&pool.mu, NULL);
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc((size_t)threads * sizeof(pthread_t));
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, mt_worker, &pool);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th); pthread_mutex_destroy(&pool.mu);
    size_t total = 0;
    for (size_t i = 0; i < pool.n; i++) { if (pool.items[i].rc && !rc) { rc = pool.items[i].rc; *e = pool.items[i].e; } total += pool.items[i].out.len; }
    if (!rc) {
        uint8_t *o = (uint8_t *)malloc(total ? total : 1); size_t k = 0;
        for (size_t i = 0; i < pool.n; i++) { memcpy(o + k, pool.items[i].out.p, pool.items[i].out.len); k += pool.items[i].out.len; }
        *out = o; *out_len = total;
    }
    for (size_t i = 0; i < pool.n; i++) free(pool.items[i].out.p);
    free(pool.items); return rc;
}

void rc_free(void *p) { free(p); }
const char *rc_strerror(int code) {
    switch (code) {
    case RC_OK: return "Ok";
    case RC_NotEnoughBytes: return "parsing::Error::NotEnoughBytes";
    case RC_NotEnoughBits: return "parsing::Error::NotEnoughBits";
    case RC_MaximumReadableBitsExceeded: return "parsing::Error::MaximumReadableBitsExceeded";
    case RC_EmptyInputData: return "parsing::Error::EmptyInputData";
    case RC_NullByte: return "parsing::Error::NullByte";
    case RC_EmptySliceError: return "parsing::Error::EmptySliceError";
    case RC_LargeAccuracyLog: return "decoders::Error::LargeAccuracyLog";
    case RC_CorruptedTable: return "decoders::Error::CorruptedTable";
    case RC_SequenceCodeMaxValueExceeded: return "decoders::Error::SequenceCodeMaxValueExceeded";
    case RC_HuffmanDecoderMissing: return "literals::Error::HuffmanDecoderMissing";
    case RC_CorruptedStreamsSizeTooBig: return "literals::Error::CorruptedStreamsSizeTooBig";
    case RC_SeqReservedSet: return "sequences::Error::ReservedSet";
    case RC_NoPreviousDecoder: return "sequences::Error::NoPreviousDecoder";
    case RC_WindowSizeTooBig: return "Error::WindowSizeTooBig";
    case RC_NullOffsetError: return "decoding_context::Error::NullOffsetError";
    case RC_ImpossibleValue: return "decoding_context::Error::ImpossibleValue";
    case RC_ReservedBlockType: return "block::Error::ReservedBlockType";
    case RC_UnrecognizedMagic: return "frame::Error::UnrecognizedMagic";
    case RC_FrameReservedSet: return "frame::Error::ReservedSet";
    case RC_MissingChecksum: return "frame::Error::MissingChecksum";
    case RC_Panic: return "panic";
    default: return "unknown";
    }
}
