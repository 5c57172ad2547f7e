/*
 * stenos_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked, loaded or called by stenos_b200/).
 *
 * Plain-C, scalar restatement of the level-1 hot path of Thermadiag/stenos v0.2, written from the
 * format rules -- not a transcription of the SIMD code.  It is the CPU
 * checker the CUDA path is compared against.  Parity of THIS file is pinned by tests/test_oracle.py
 * against (a) the golden vectors under tests/golden/ that were produced by the untouched reference
 * and (b) the compiled reference itself (oracle/_ref/libstenos_ref.so) when present.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 *
 * Scope: superblock codes 1 (BLOCK) and 6 (COPY) -- level 1, any element size T >= 2 whose LZ
 * element width is 4 or 8 bytes (T%4==0) or for which LZ is not attempted (T%4 != 0); level 0;
 * the byte shuffle / delta filters.  The < 128 byte superblock (code 2, Zstd) is delegated to
 * libzstd.so.1 through dlopen when available (stenos/internal/stenos.cpp:435-437).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>

#define SO_ERR_UNDEFINED ((size_t)-1)
#define SO_ERR_SRC_OVERFLOW ((size_t)-2)
#define SO_ERR_ALLOC ((size_t)-3)
#define SO_ERR_INVALID_INPUT ((size_t)-4)
#define SO_ERR_DST_OVERFLOW ((size_t)-6)
#define SO_ERR_INVALID_BYTESOFTYPE ((size_t)-7)
#define SO_ERR_ZSTD_INTERNAL ((size_t)-8)
#define SO_ERR_INVALID_PARAMETER ((size_t)-9)
#define SO_LAST_ERROR ((size_t)-100)
#define SO_IS_ERR(r) ((size_t)(r) >= SO_LAST_ERROR)

#define SO_SUPERBLOCK 131072u
#define SO_MAX_BLOCK_BYTES ((1u << 24) - 1u)

enum { KIND_SAME = 0, KIND_RAW = 1, KIND_NORMAL = 2, KIND_NORMAL_RLE = 3 };
enum { MARK_COPY = 252, MARK_LZ = 253, MARK_PARTIAL = 254 };
enum { CODE_BLOCK = 1, CODE_ZSTD = 2, CODE_COPY = 6 };

/* ------------------------------------------------------------------------------------------ */
/* Filters                                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* stenos/internal/shuffle-generic.h:33-74, shuffle.cpp:82-90: dst[k*n+j] = src[j*T+k]; tail copied. */
void so_shuffle(size_t T, size_t bytes, const uint8_t* src, uint8_t* dst)
{
	if (T <= 1) {
		memcpy(dst, src, bytes);
		return;
	}
	size_t n = bytes / T;
	for (size_t k = 0; k < T; ++k)
		for (size_t j = 0; j < n; ++j)
			dst[k * n + j] = src[j * T + k];
	memcpy(dst + n * T, src + n * T, bytes - n * T);
}

/* stenos/internal/shuffle-generic.h:83-125, shuffle.cpp:94-102 */
void so_unshuffle(size_t T, size_t bytes, const uint8_t* src, uint8_t* dst)
{
	if (T <= 1) {
		memcpy(dst, src, bytes);
		return;
	}
	size_t n = bytes / T;
	for (size_t j = 0; j < n; ++j)
		for (size_t k = 0; k < T; ++k)
			dst[j * T + k] = src[k * n + j];
	memcpy(dst + n * T, src + n * T, bytes - n * T);
}

/* stenos/internal/delta.cpp:30-71: four independent quarter streams when bytes > 2048; the bytes
 * from 4*(bytes/4) on continue the chain of the last stream. */
void so_delta(const uint8_t* src, uint8_t* dst, size_t bytes)
{
	if (!bytes)
		return;
	for (size_t i = bytes; i-- > 1;)
		dst[i] = (uint8_t)(src[i] - src[i - 1]);
	dst[0] = src[0];
	if (bytes > 2048) {
		size_t q = bytes / 4;
		for (int k = 1; k < 4; ++k)
			dst[k * q] = src[k * q];
	}
}

/* stenos/internal/delta.cpp:230-267 */
void so_delta_inv(const uint8_t* src, uint8_t* dst, size_t bytes)
{
	if (!bytes)
		return;
	size_t q = bytes > 2048 ? bytes / 4 : 0;
	uint8_t acc = 0;
	for (size_t i = 0; i < bytes; ++i) {
		if (i == 0 || (q && i < 4 * q && i % q == 0))
			acc = src[i];
		else
			acc = (uint8_t)(acc + src[i]);
		dst[i] = acc;
	}
}

/* ------------------------------------------------------------------------------------------ */
/* Plane analysis (find_pack_bits_params, block_compress.h:385-535)                            */
/* ------------------------------------------------------------------------------------------ */

typedef struct
{
	int kind;           /* KIND_* */
	unsigned size;      /* encoded size of the plane */
	uint8_t first;      /* first byte of the plane */
	uint8_t hdr[16];    /* per-row header nibble */
	uint8_t mins[16];   /* per-row minimum (of the values or of the deltas) */
	uint8_t bits[16];   /* per-row bit width */
	uint8_t type[16];   /* 0: plain, 1: delta */
	uint8_t rsize[16];  /* per-row encoded size (before the mins-RLE adjustment) */
	uint16_t rmask[16]; /* RLE mask on values */
	uint16_t dmask[16]; /* RLE mask on deltas */
	uint16_t mmask;     /* RLE mask on the vector of mins */
	uint8_t delta[16][16];
} so_plane;

static int s8(uint8_t v) { return (int)(int8_t)v; }

/* bit_scan_reverse_8_2 (block_compress.h:334-352): width 7 is reported as 8 */
static unsigned so_nbits(unsigned v)
{
	v &= 255u;
	if (v >= 64)
		return 8;
	unsigned n = 0;
	while (v) {
		++n;
		v >>= 1;
	}
	return n;
}

/* compute_rle_row_single / compute_rle_row (block_compress.h:248-275): bit k set when value k
 * repeats its predecessor (prev for k == 0).  Returns the number of NON repeated values. */
static unsigned so_rle_mask(const uint8_t* v, uint8_t prev, uint16_t* mask)
{
	unsigned m = 0, cnt = 0;
	for (int k = 0; k < 16; ++k) {
		if (v[k] == prev)
			m |= 1u << k;
		else
			++cnt;
		prev = v[k];
	}
	*mask = (uint16_t)m;
	return cnt;
}

static void so_analyse_plane(const uint8_t* P, int rle, so_plane* a)
{
	a->first = P[0];
	int same = 1;
	for (int i = 1; i < 256 && same; ++i)
		same = (P[i] == P[0]);
	if (same) { /* :396-418 */
		a->kind = KIND_SAME;
		a->size = 1;
		return;
	}
	unsigned total = 8, nomin = 0;
	for (int r = 0; r < 16; ++r) {
		const uint8_t* row = P + 16 * r;
		uint8_t prev_last = r ? P[16 * r - 1] : 0; /* :399 -- 0 before the first row */
		int mn = 127, mx = -128, mnd = 127, mxd = -128;
		for (int k = 0; k < 16; ++k) {
			uint8_t d = (uint8_t)(row[k] - (k ? row[k - 1] : prev_last));
			a->delta[r][k] = d;
			if (s8(row[k]) < mn) mn = s8(row[k]);
			if (s8(row[k]) > mx) mx = s8(row[k]);
			if (s8(d) < mnd) mnd = s8(d);
			if (s8(d) > mxd) mxd = s8(d);
		}
		unsigned b0 = so_nbits((unsigned)(mx - mn)), b1 = so_nbits((unsigned)(mxd - mnd));
		if (b0 == 6) /* :422 header 6 is reserved for delta-RLE */
			b0 = 8;
		unsigned b = b0 < b1 ? b0 : b1;
		int plain = (b0 == b); /* ties go to the plain type, :426-427 */
		a->bits[r] = (uint8_t)b;
		a->type[r] = plain ? 0 : 1;
		a->mins[r] = (uint8_t)(plain ? mn : mnd);
		unsigned sz = 2 * b + (b == 8 ? 0 : 1); /* :433-435 */
		unsigned h = plain ? (b0 == 8 ? 15 : b0) : (8 + (b1 == 8 ? 7 : b1)); /* :499-502 */
		a->rmask[r] = a->dmask[r] = 0;
		if (rle) { /* :439-474 */
			unsigned rs = so_rle_mask(row, prev_last, &a->rmask[r]) + 2;
			unsigned ds = so_rle_mask(a->delta[r], 0, &a->dmask[r]) + 2;
			int use_rle = rs < sz;
			if (rs < sz) sz = rs;
			int use_drle = ds < sz;
			if (ds < sz) sz = ds;
			if (use_drle)
				h = 6;
			else if (use_rle)
				h = 7;
		}
		a->hdr[r] = (uint8_t)h;
		a->rsize[r] = (uint8_t)sz;
		total += sz;
		if (h == 6 || h == 7 || h == 15)
			++nomin;
	}
	a->kind = KIND_NORMAL;
	a->mmask = 0;
	if (rle) { /* :478-490 -- RLE over the whole vector of 16 mins, predecessor 0 */
		unsigned cnt = so_rle_mask(a->mins, 0, &a->mmask);
		if (cnt + 2 < 16 - nomin) {
			a->kind = KIND_NORMAL_RLE;
			total -= (16 - nomin) - (cnt + 2);
		}
	}
	a->size = total;
}

/* write_16 / write_16_bmi2 (block_compress.h:540-602): two groups of eight values, each packed
 * LSB first into `bits` bytes. */
static uint8_t* so_pack16(const uint8_t* v, unsigned bits, uint8_t* out)
{
	for (int g = 0; g < 16; g += 8) {
		uint64_t w = 0;
		for (int k = 0; k < 8; ++k)
			w |= (uint64_t)v[g + k] << (k * bits);
		for (unsigned b = 0; b < bits; ++b)
			*out++ = (uint8_t)(w >> (8 * b));
	}
	return out;
}

/* write_rle_single (block_compress.h:258-265): 2-byte mask, then the non repeated values */
static uint8_t* so_emit_rle(const uint8_t* v, uint16_t mask, uint8_t* out)
{
	*out++ = (uint8_t)mask;
	*out++ = (uint8_t)(mask >> 8);
	for (int k = 0; k < 16; ++k)
		if (!((mask >> k) & 1))
			*out++ = v[k];
	return out;
}

/* encode16x16_generic (block_compress.h:739-806) and encode_lines (:686-737, lines < 16) */
static uint8_t* so_encode_plane(const so_plane* a, const uint8_t* P, unsigned lines, uint8_t* out)
{
	if (a->kind == KIND_SAME) {
		*out++ = a->first;
		return out;
	}
	if (a->kind == KIND_RAW) {
		memcpy(out, P, 256);
		return out + 256;
	}
	for (unsigned i = 0; i + 1 < lines; i += 2)
		*out++ = (uint8_t)(a->hdr[i] | (a->hdr[i + 1] << 4));
	if (lines & 1)
		*out++ = a->hdr[lines - 1];
	if (a->kind == KIND_NORMAL_RLE) {
		out = so_emit_rle(a->mins, a->mmask, out);
	}
	else {
		for (unsigned r = 0; r < lines; ++r)
			if (a->hdr[r] != 6 && a->hdr[r] != 7 && a->hdr[r] != 15)
				*out++ = a->mins[r];
	}
	for (unsigned r = 0; r < lines; ++r) {
		unsigned h = a->hdr[r];
		const uint8_t* row = P + 16 * r;
		if (h == 15) {
			memcpy(out, row, 16);
			out += 16;
		}
		else if (h == 7)
			out = so_emit_rle(row, a->rmask[r], out);
		else if (h == 6)
			out = so_emit_rle(a->delta[r], a->dmask[r], out);
		else if (a->bits[r]) {
			uint8_t v[16];
			const uint8_t* s = a->type[r] ? a->delta[r] : row;
			for (int k = 0; k < 16; ++k)
				v[k] = (uint8_t)(s[k] - a->mins[r]);
			out = so_pack16(v, a->bits[r], out);
		}
	}
	return out;
}

/* ------------------------------------------------------------------------------------------ */
/* LZ-like matcher (lz_compress.h:47-232)                                                      */
/* ------------------------------------------------------------------------------------------ */

static uint64_t so_load(const uint8_t* p, unsigned B)
{
	uint64_t v = 0;
	memcpy(&v, p, B);
	return v;
}
static unsigned so_hash(uint64_t v, unsigned B)
{
	if (B == 8 || B == 6)
		return (unsigned)((v * 14313749767032793493ULL) >> 56); /* :52-56 */
	return (unsigned)(((uint32_t)v * 2654435761u) & 255u);         /* :47-51 */
}

/* lz_compress<B> (:191-232). Table semantics: empty at block start.
 * Returns the stream length, or 0 on failure (stream longer than max_size / early exit). */
static size_t so_lz_compress(const uint8_t* in, unsigned B, size_t count, size_t max_size, uint8_t* out)
{
	int table[256];
	for (int i = 0; i < 256; ++i)
		table[i] = -1;
	size_t produced = 0;
	unsigned failed = 0, max_failed = 3;
	int once = 0;
	for (size_t i = 0; i < count; i += 8) {
		uint8_t* anchor = out + produced++;
		*anchor = 0;
		if (failed == max_failed) {
			failed = 0;
			if (--max_failed == 0)
				max_failed = 1;
			memcpy(out + produced, in + i * B, 8 * B);
			produced += 8 * B;
		}
		else {
			for (unsigned k = 0; k < 8; ++k) {
				size_t p = i + k;
				uint64_t v = so_load(in + p * B, B);
				unsigned h = so_hash(v, B);
				int q = table[h];
				if (q >= 0 && (size_t)q < p && so_load(in + (size_t)q * B, B) == v) {
					unsigned off = (unsigned)(p - (size_t)q);
					*anchor |= (uint8_t)(1u << k);
					if (off < 128)
						out[produced++] = (uint8_t)off; /* write_diff :140-151 */
					else {
						out[produced++] = (uint8_t)((off & 127) | 128);
						out[produced++] = (uint8_t)(off >> 7);
					}
				}
				else {
					memcpy(out + produced, in + p * B, B);
					produced += B;
				}
				table[h] = (int)p;
			}
			failed += (*anchor == 0);
		}
		if (produced > max_size)
			return 0;
		if (!once && i > count / 4) { /* :224-229, 0.4 == 2/5 exactly in integers */
			if (5 * produced > 2 * max_size)
				return 0;
			once = 1;
		}
	}
	return produced;
}

/* lz_decompress<B> (:234-277). Returns bytes consumed, 0 on error. */
static size_t so_lz_decompress(const uint8_t* in, size_t in_size, unsigned B, size_t count, uint8_t* dst)
{
	const uint8_t* p = in;
	const uint8_t* end = in + in_size;
	uint8_t* d = dst;
	for (size_t i = 0; i < count; i += 8) {
		if (p + 2 > end)
			return 0;
		unsigned anchor = *p++;
		if (!anchor) {
			if (p + 8 * B > end)
				return 0;
			memcpy(d, p, 8 * B);
			d += 8 * B;
			p += 8 * B;
			continue;
		}
		for (unsigned k = 0; k < 8; ++k) {
			if ((anchor >> k) & 1) {
				unsigned off = *p & 127u;
				if (*p++ > 127u) {
					if (p == end)
						return 0;
					off |= (unsigned)(*p++) << 7;
				}
				if ((size_t)(d - dst) < (size_t)off * B)
					return 0; /* the reference only asserts this in debug builds */
				memmove(d, d - (size_t)off * B, B);
				d += B;
			}
			else {
				if (p + B > end)
					return 0;
				memcpy(d, p, B);
				d += B;
				p += B;
			}
		}
	}
	return (size_t)(p - in);
}

static unsigned so_lz_width(size_t T)
{
	/* lz_compress_generic dispatch (:279-299) */
	if (T > 512) return 0;
	if (T % 8 == 0) return 8;
	if (T <= 2 || T % 4 == 0) return 4;
	if (T % 6 == 0) return 6;
	if (T % 3 == 0) return 3;
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Block stream (block_compress, block_compress.h:1099-1302)                                   */
/* ------------------------------------------------------------------------------------------ */

static void so_planes_of(const uint8_t* src, size_t T, uint8_t* planes)
{
	for (size_t j = 0; j < 256; ++j)
		for (size_t p = 0; p < T; ++p)
			planes[p * 256 + j] = src[j * T + p];
}

/* block_compress_partial (:947-1020). dst points after the 254 marker. */
static size_t so_encode_partial(const uint8_t* src, size_t T, size_t bytes, uint8_t* dst, uint8_t* dst_end, uint8_t* scratch)
{
	size_t line = 16 * T, lines = bytes / line, hs = (T + 1) / 2;
	uint8_t* out = dst;
	if (lines) {
		uint8_t* buf = scratch;            /* T*256 */
		uint8_t* planes = scratch + T * 256; /* T*256 */
		memcpy(buf, src, bytes);
		memset(buf + bytes, buf[bytes - 1], 256 * T - bytes); /* :967-968 pad with the last BYTE */
		so_planes_of(buf, T, planes);
		uint8_t* head = out;
		out += hs;
		for (size_t p = 0; p < T; ++p) {
			so_plane a;
			so_analyse_plane(planes + p * 256, 0, &a); /* RLE off, no RAW promotion (:982) */
			if (a.kind == KIND_SAME) {
				if (out >= dst_end)
					return SO_ERR_DST_OVERFLOW;
				*out++ = a.first;
			}
			else {
				unsigned size = 8;
				for (size_t j = 0; j < lines; ++j)
					size += a.rsize[j];
				if (out + size + 8 > dst_end) /* :994 */
					return SO_ERR_DST_OVERFLOW;
				out = so_encode_plane(&a, planes + p * 256, (unsigned)lines, out);
			}
			if ((p & 1) == 0)
				head[p >> 1] = 0;
			head[p >> 1] |= (uint8_t)(a.kind << (4 * (p & 1)));
		}
	}
	size_t rem = bytes - lines * line;
	if (rem) {
		if (out + rem > dst_end)
			return SO_ERR_DST_OVERFLOW;
		memcpy(out, src + lines * line, rem);
		out += rem;
	}
	return (size_t)(out - dst);
}

/* block_compress with block_level 2 (RLE + LZ), no time limit, no pre-shuffled input.
 * The dst room arithmetic is kept exactly: it changes decisions. */
size_t so_block_compress(const uint8_t* src, size_t T, size_t bytes, uint8_t* dst, size_t dst_size)
{
	if (!bytes)
		return 0;
	size_t bs = T * 256, hs = (T + 1) / 2;
	size_t nblocks = (bytes == bs) ? 1 : bytes / bs;
	uint8_t* out = dst;
	uint8_t* dst_end = dst + dst_size;
	uint8_t* scratch = (uint8_t*)malloc(2 * bs + 16);
	so_plane* an = (so_plane*)malloc(sizeof(so_plane) * T);
	if (!scratch || !an) {
		free(scratch);
		free(an);
		return SO_ERR_ALLOC;
	}
	uint8_t* planes = scratch;
	size_t ret = 0;
	unsigned B = so_lz_width(T);
	for (size_t b = 0; b < nblocks; ++b, src += bs) {
		uint8_t* head = out;
		out += hs;
		so_planes_of(src, T, planes);
		size_t full = 0;
		for (size_t p = 0; p < T; ++p) {
			so_analyse_plane(planes + p * 256, 1, &an[p]);
			if (an[p].size > 256) { /* :1200-1204 */
				an[p].kind = KIND_RAW;
				an[p].size = 256;
			}
			full += an[p].size;
		}
		if (T % 4 == 0 && full * 3 > bs) { /* :1210 */
			if (dst_end > out + (full + T * 8u + 2u)) { /* :1214 */
				size_t r = B ? so_lz_compress(src, B, bs / B, full, head + 1) : 0;
				if (r) {
					*head = MARK_LZ;
					out = head + 1 + r;
					continue;
				}
			}
		}
		if (out + full > dst_end) { /* :1225 */
			ret = SO_ERR_DST_OVERFLOW;
			goto done;
		}
		for (size_t p = 0; p < T; ++p) {
			if (an[p].kind == KIND_RAW) {
				memcpy(out, planes + p * 256, 256);
				out += 256;
			}
			else {
				if (out + an[p].size + 16 > dst_end) { /* :1241 */
					ret = SO_ERR_DST_OVERFLOW;
					goto done;
				}
				out = so_encode_plane(&an[p], planes + p * 256, 16, out);
			}
			if ((p & 1) == 0) {
				if (head + (p >> 1) >= dst_end) { /* :1248 */
					ret = SO_ERR_DST_OVERFLOW;
					goto done;
				}
				head[p >> 1] = 0;
			}
			head[p >> 1] |= (uint8_t)(an[p].kind << (4 * (p & 1)));
		}
	}
	{
		size_t rem = bytes - nblocks * bs;
		if (rem) { /* :1277-1298 */
			if (out + 2 > dst_end) {
				ret = SO_ERR_DST_OVERFLOW;
				goto done;
			}
			*out++ = MARK_PARTIAL;
			size_t r = so_encode_partial(src, T, rem, out, dst_end, scratch);
			if (SO_IS_ERR(r)) {
				ret = r;
				goto done;
			}
			out += r;
		}
	}
	ret = (size_t)(out - dst);
done:
	free(scratch);
	free(an);
	return ret;
}

/* ------------------------------------------------------------------------------------------ */
/* Block stream decoder (block_decompress_sse / block_decompress, block_compress.h:1797-2175)  */
/* ------------------------------------------------------------------------------------------ */

static const uint8_t* so_read_rle(const uint8_t* p, const uint8_t* end, uint8_t prev, uint8_t* out)
{
	/* decode_rle / decode_rle_flat (:1585-1613, :1939-1968) */
	if (end - p < 2)
		return NULL;
	unsigned mask = p[0] | (p[1] << 8);
	p += 2;
	unsigned need = 0;
	for (int k = 0; k < 16; ++k)
		need += !((mask >> k) & 1);
	if (need > (size_t)(end - p))
		return NULL;
	for (int k = 0; k < 16; ++k) {
		if (!((mask >> k) & 1))
			prev = *p++;
		out[k] = prev;
	}
	return p;
}

static const uint8_t* so_unpack16(const uint8_t* p, const uint8_t* end, unsigned bits, uint8_t* v)
{
	if ((size_t)(end - p) < 2 * bits)
		return NULL;
	for (int g = 0; g < 16; g += 8) {
		uint64_t w = 0;
		for (unsigned b = 0; b < bits; ++b)
			w |= (uint64_t)(*p++) << (8 * b);
		for (int k = 0; k < 8; ++k)
			v[g + k] = (uint8_t)((w >> (k * bits)) & ((1u << bits) - 1u));
	}
	return p;
}

/* decode_block / decode_block_rle (+_flat variants) and decode_line (:1615-1745, :1970-2084).
 * P receives `lines` rows of 16 bytes. */
static const uint8_t* so_decode_plane(const uint8_t* p, const uint8_t* end, int kind, unsigned lines, uint8_t* P)
{
	uint8_t hdr[16], mins[16];
	unsigned hl = lines / 2 + (lines & 1);
	memset(mins, 0, sizeof(mins));
	if ((size_t)(end - p) < hl)
		return NULL;
	for (unsigned i = 0; i < lines; ++i)
		hdr[i] = (uint8_t)((p[i >> 1] >> (4 * (i & 1))) & 15);
	p += hl;
	if (kind == KIND_NORMAL_RLE) {
		p = so_read_rle(p, end, 0, mins);
		if (!p)
			return NULL;
	}
	else {
		for (unsigned i = 0; i < lines; ++i)
			if (hdr[i] != 6 && hdr[i] != 7 && hdr[i] != 15) {
				if (p >= end)
					return NULL;
				mins[i] = *p++;
			}
	}
	for (unsigned r = 0; r < lines; ++r) {
		uint8_t* row = P + 16 * r;
		uint8_t prev = r ? P[16 * r - 1] : 0;
		unsigned h = hdr[r];
		if (h == 15) {
			if (end - p < 16)
				return NULL;
			memcpy(row, p, 16);
			p += 16;
		}
		else if (h == 7) {
			p = so_read_rle(p, end, prev, row);
			if (!p)
				return NULL;
		}
		else if (h == 6) {
			uint8_t d[16];
			p = so_read_rle(p, end, 0, d);
			if (!p)
				return NULL;
			for (int k = 0; k < 16; ++k)
				row[k] = prev = (uint8_t)(prev + d[k]);
		}
		else {
			unsigned bits = h & 7; /* h in 0..5 or 8..14 */
			uint8_t v[16];
			memset(v, 0, 16);
			if (bits) {
				p = so_unpack16(p, end, bits, v);
				if (!p)
					return NULL;
			}
			if (h < 8)
				for (int k = 0; k < 16; ++k)
					row[k] = (uint8_t)(v[k] + mins[r]);
			else
				for (int k = 0; k < 16; ++k)
					row[k] = prev = (uint8_t)(prev + v[k] + mins[r]);
		}
	}
	return p;
}

/* block_decompress_partial (:1749-1795). Returns consumed bytes or an error code. */
static size_t so_decode_partial(const uint8_t* src, size_t size, size_t T, size_t bytes, uint8_t* dst)
{
	const uint8_t* p = src;
	const uint8_t* end = src + size;
	size_t line = 16 * T, lines = bytes / line, hs = (T + 1) / 2;
	if (lines) {
		const uint8_t* head = p;
		p += hs;
		if (p >= end)
			return SO_ERR_SRC_OVERFLOW;
		uint8_t P[256];
		for (size_t i = 0; i < T; ++i) {
			int kind = (head[i >> 1] >> (4 * (i & 1))) & 15;
			if (kind == KIND_SAME) {
				if (p >= end)
					return SO_ERR_SRC_OVERFLOW;
				memset(P, *p++, 256);
			}
			else if (kind == KIND_NORMAL) {
				p = so_decode_plane(p, end, kind, (unsigned)lines, P);
				if (!p)
					return SO_ERR_SRC_OVERFLOW;
			}
			else
				return SO_ERR_INVALID_INPUT;
			for (size_t j = 0; j < lines * 16; ++j)
				dst[j * T + i] = P[j];
		}
	}
	size_t rem = bytes - lines * line;
	if (rem) {
		if (p + rem > end)
			return SO_ERR_SRC_OVERFLOW;
		memcpy(dst + lines * line, p, rem);
		p += rem;
	}
	return (size_t)(p - src);
}

size_t so_block_decompress(const uint8_t* src, size_t size, size_t T, size_t bytes, uint8_t* dst)
{
	if (!bytes || !size)
		return 0;
	const uint8_t* p = src;
	const uint8_t* end = src + size;
	size_t bs = T * 256, hs = (T + 1) / 2;
	size_t nblocks = (bytes == bs) ? 1 : bytes / bs;
	if (size < hs + T && nblocks)
		return SO_ERR_SRC_OVERFLOW;
	uint8_t P[256];
	for (size_t b = 0; b < nblocks; ++b, dst += bs) {
		const uint8_t* head = p;
		p += hs;
		if (p >= end)
			return SO_ERR_SRC_OVERFLOW;
		if (*head == MARK_COPY) {
			p = head + 1;
			if ((size_t)(end - p) < bs)
				return SO_ERR_SRC_OVERFLOW;
			memcpy(dst, p, bs);
			p += bs;
			continue;
		}
		if (*head == MARK_LZ) {
			unsigned B = so_lz_width(T);
			p = head + 1;
			size_t r = B ? so_lz_decompress(p, (size_t)(end - p), B, bs / B, dst) : 0;
			if (!r)
				return SO_ERR_INVALID_INPUT;
			p += r;
			continue;
		}
		for (size_t i = 0; i < T; ++i) {
			int kind = (head[i >> 1] >> (4 * (i & 1))) & 15;
			if (kind == KIND_RAW) {
				if (end - p < 256)
					return SO_ERR_SRC_OVERFLOW;
				memcpy(P, p, 256);
				p += 256;
			}
			else if (kind == KIND_SAME) {
				if (p >= end)
					return SO_ERR_SRC_OVERFLOW;
				memset(P, *p++, 256);
			}
			else if (kind == KIND_NORMAL || kind == KIND_NORMAL_RLE) {
				p = so_decode_plane(p, end, kind, 16, P);
				if (!p)
					return SO_ERR_SRC_OVERFLOW;
			}
			else
				return SO_ERR_INVALID_INPUT;
			for (size_t j = 0; j < 256; ++j)
				dst[j * T + i] = P[j];
		}
	}
	size_t rem = bytes - nblocks * bs;
	if (rem) {
		if (p == end)
			return SO_ERR_SRC_OVERFLOW;
		if (*p++ != MARK_PARTIAL)
			return SO_ERR_INVALID_INPUT;
		size_t r = so_decode_partial(p, (size_t)(end - p), T, rem, dst);
		if (SO_IS_ERR(r))
			return r;
		p += r;
	}
	return (size_t)(p - src);
}

/* ------------------------------------------------------------------------------------------ */
/* Superblocks and frames (stenos.cpp)                                                         */
/* ------------------------------------------------------------------------------------------ */

typedef size_t (*zstd_compress_fn)(void*, size_t, const void*, size_t, int);
typedef size_t (*zstd_decompress_fn)(void*, size_t, const void*, size_t);
typedef unsigned (*zstd_iserror_fn)(size_t);
static zstd_compress_fn p_zstd_compress;
static zstd_decompress_fn p_zstd_decompress;
static zstd_iserror_fn p_zstd_iserror;

static int so_load_zstd(void)
{
	static int tried = 0;
	if (!tried) {
		tried = 1;
		void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
		if (h) {
			p_zstd_compress = (zstd_compress_fn)dlsym(h, "ZSTD_compress");
			p_zstd_decompress = (zstd_decompress_fn)dlsym(h, "ZSTD_decompress");
			p_zstd_iserror = (zstd_iserror_fn)dlsym(h, "ZSTD_isError");
		}
	}
	return p_zstd_compress && p_zstd_decompress && p_zstd_iserror;
}

static void put24(uint8_t* d, size_t v)
{
	d[0] = (uint8_t)v;
	d[1] = (uint8_t)(v >> 8);
	d[2] = (uint8_t)(v >> 16);
}
static size_t get24(const uint8_t* s) { return (size_t)s[0] | ((size_t)s[1] << 8) | ((size_t)s[2] << 16); }

/* compress_memcpy (stenos.cpp:363-374) */
static size_t so_copy_superblock(const uint8_t* src, size_t bytes, uint8_t* dst, size_t dst_size)
{
	if (dst_size < bytes + 4)
		return SO_ERR_DST_OVERFLOW;
	dst[0] = CODE_COPY;
	put24(dst + 1, bytes);
	memcpy(dst + 4, src, bytes);
	return bytes + 4;
}

/* compress_generic_superblock restricted to level 0 / level 1 with T > 1 (stenos.cpp:403-450,606-615,658-678) */
size_t so_compress_superblock(const uint8_t* src, size_t T, size_t bytes, uint8_t* dst, size_t dst_size, int level)
{
	if (dst_size < 4)
		return SO_ERR_DST_OVERFLOW;
	if (bytes == 0 || level == 0)
		return so_copy_superblock(src, bytes, dst, dst_size);
	if (bytes < 128) { /* :435-437 -> ZSTD at zstd level 1 (zstd_from_reduced_level(0)) */
		if (!so_load_zstd())
			return SO_ERR_ZSTD_INTERNAL;
		size_t r = p_zstd_compress(dst + 4, dst_size - 4, src, bytes, 1);
		if (p_zstd_iserror(r) || r > bytes)
			return so_copy_superblock(src, bytes, dst, dst_size);
		if (4 + r > dst_size)
			return SO_ERR_DST_OVERFLOW;
		dst[0] = CODE_ZSTD;
		put24(dst + 1, r);
		return r + 4;
	}
	size_t r = so_block_compress(src, T, bytes, dst + 4, dst_size - 4);
	if (SO_IS_ERR(r) || r > bytes)
		return so_copy_superblock(src, bytes, dst, dst_size);
	if (4 + r > dst_size)
		return SO_ERR_DST_OVERFLOW;
	dst[0] = CODE_BLOCK;
	put24(dst + 1, r);
	return r + 4;
}

/* super_block_size (stenos.cpp:71-76) */
size_t so_default_superblock(size_t T)
{
	size_t bs = T * 256;
	return bs > SO_SUPERBLOCK ? bs : (SO_SUPERBLOCK / bs) * bs;
}

/* stenos::compress_bound (stenos.h:37-42) */
size_t so_bound(size_t bytes)
{
	size_t n = bytes / 65792 + (bytes % 65792 ? 1 : 0);
	return 12 + (n ? n : 1) * 4 + bytes;
}

/* stenos_compress_generic, single threaded (stenos.cpp:844-907), prepare (:115-185).
 * block_shift == (size_t)-1: default superblock size; otherwise custom (T*256)<<block_shift. */
size_t so_compress(const uint8_t* src, size_t T, size_t bytes, uint8_t* dst, size_t dst_size, int level, size_t block_shift)
{
	if (T == 0 || T >= SO_MAX_BLOCK_BYTES / 256)
		return SO_ERR_INVALID_BYTESOFTYPE;
	if (T < 2 && level != 0)
		return SO_ERR_INVALID_PARAMETER; /* T == 1 at level >= 1 is Zstd only: out of scope */
	if (level > 1)
		return SO_ERR_INVALID_PARAMETER;
	size_t sb, shift = 0;
	if (block_shift != (size_t)-1) {
		if (block_shift >= 16)
			return SO_ERR_INVALID_PARAMETER;
		sb = (T * 256) << block_shift;
		shift = 255;
	}
	else {
		sb = so_default_superblock(T); /* level <= 1 -> shift 0 */
	}
	if (sb < T * 256 || sb >= SO_MAX_BLOCK_BYTES)
		return SO_ERR_INVALID_PARAMETER;
	uint8_t* out = dst;
	uint8_t* end = dst + dst_size;
	if (out + 8 > end)
		return SO_ERR_DST_OVERFLOW;
	*out++ = (uint8_t)shift;
	for (int i = 0; i < 7; ++i)
		*out++ = (uint8_t)((uint64_t)bytes >> (8 * i));
	if (shift == 255) {
		if (out + 4 > end)
			return SO_ERR_DST_OVERFLOW;
		for (int i = 0; i < 4; ++i)
			*out++ = (uint8_t)(sb >> (8 * i));
	}
	if (!bytes)
		return (size_t)(out - dst);
	size_t count = bytes / sb + (bytes % sb ? 1 : 0);
	for (size_t i = 0; i < count; ++i) {
		size_t in = (i == count - 1) ? bytes - i * sb : sb;
		size_t r = so_compress_superblock(src + i * sb, T, in, out, (size_t)(end - out), level);
		if (SO_IS_ERR(r))
			return r;
		out += r;
	}
	return (size_t)(out - dst);
}

/* decompress_generic_superblock (stenos.cpp:681-753), codes 1, 2, 6 */
static size_t so_decompress_superblock(unsigned code, const uint8_t* src, size_t T, size_t csize, uint8_t* dst, size_t dsize)
{
	switch (code) {
		case CODE_BLOCK: {
			size_t r = so_block_decompress(src, csize, T, dsize, dst);
			if (SO_IS_ERR(r))
				return SO_ERR_INVALID_INPUT;
		} break;
		case CODE_ZSTD: {
			if (!so_load_zstd())
				return SO_ERR_ZSTD_INTERNAL;
			size_t r = p_zstd_decompress(dst, dsize, src, csize);
			if (p_zstd_iserror(r))
				return SO_ERR_INVALID_INPUT;
		} break;
		case CODE_COPY:
			if (dsize != csize)
				return SO_ERR_INVALID_INPUT;
			memcpy(dst, src, csize);
			break;
		default:
			return SO_ERR_INVALID_INPUT;
	}
	return dsize;
}

/* stenos_decompress_generic, single threaded (stenos.cpp:1052-1149).
 * Deliberate divergence: a final superblock whose remainder is 0 is
 * decoded as a FULL superblock; the reference computes dsize = 0 for it and fails. */
size_t so_decompress(const uint8_t* src, size_t T, size_t size, uint8_t* dst, size_t dst_size)
{
	if (T == 0 || T >= SO_MAX_BLOCK_BYTES / 256)
		return SO_ERR_INVALID_BYTESOFTYPE;
	const uint8_t* p = src;
	const uint8_t* end = src + size;
	if (p + 8 > end)
		return SO_ERR_SRC_OVERFLOW;
	unsigned shift = *p++;
	if (shift > 4 && shift != 255)
		return SO_ERR_INVALID_INPUT;
	uint64_t total = 0;
	for (int i = 0; i < 7; ++i)
		total |= (uint64_t)(*p++) << (8 * i);
	if (total > dst_size)
		return SO_ERR_DST_OVERFLOW;
	if (!total)
		return 0;
	size_t sb;
	if (shift == 255) {
		if (p + 4 > end)
			return SO_ERR_SRC_OVERFLOW;
		sb = (size_t)p[0] | ((size_t)p[1] << 8) | ((size_t)p[2] << 16) | ((size_t)p[3] << 24);
		p += 4;
		if (!sb)
			return SO_ERR_INVALID_INPUT;
	}
	else
		sb = so_default_superblock(T) << shift;
	size_t count = total / sb + (total % sb ? 1 : 0);
	uint8_t* out = dst;
	for (size_t i = 0; i < count; ++i) {
		if (p + 4 > end)
			return SO_ERR_SRC_OVERFLOW;
		unsigned code = *p++;
		size_t csize = get24(p);
		p += 3;
		size_t dsize = (i == count - 1) ? (size_t)(total - (uint64_t)i * sb) : sb;
		if (csize > (size_t)(end - p))
			return SO_ERR_INVALID_INPUT;
		size_t r = so_decompress_superblock(code, p, T, csize, out, dsize);
		if (r != dsize)
			return r;
		out += dsize;
		p += csize;
	}
	return (size_t)total;
}

/* Byte offset (from the start of the frame) of every superblock header; out[count] = end.
 * Helper for tests of the device frame index (mirrors the serial walk at stenos.cpp:1124-1143). */
size_t so_frame_index(const uint8_t* src, size_t T, size_t size, uint64_t* out, size_t cap)
{
	const uint8_t* p = src;
	const uint8_t* end = src + size;
	if (size < 8)
		return SO_ERR_SRC_OVERFLOW;
	unsigned shift = *p++;
	uint64_t total = 0;
	for (int i = 0; i < 7; ++i)
		total |= (uint64_t)(*p++) << (8 * i);
	size_t sb;
	if (shift == 255) {
		if (size < 12)
			return SO_ERR_SRC_OVERFLOW;
		sb = (size_t)p[0] | ((size_t)p[1] << 8) | ((size_t)p[2] << 16) | ((size_t)p[3] << 24);
		p += 4;
	}
	else
		sb = so_default_superblock(T) << shift;
	size_t count = total ? total / sb + (total % sb ? 1 : 0) : 0;
	if (count + 1 > cap)
		return SO_ERR_DST_OVERFLOW;
	for (size_t i = 0; i < count; ++i) {
		if (p + 4 > end)
			return SO_ERR_SRC_OVERFLOW;
		out[i] = (uint64_t)(p - src);
		p += 4 + get24(p + 1);
	}
	out[count] = (uint64_t)(p - src);
	return count;
}
