// sb_decode.cuh -- one-warp-per-256-element-block decoder (device side).
//
// Exact inverse of sb_encode.cuh; replaces block_decompress_sse / block_decompress
// (block_compress.h:2088-2175 / 1797-1879), decode_block_flat[_rle] and decolde_line_flat
// (:1970-2084), read_16_bits (:1451-1486), decode_rle_flat (:1939-1968), lz_decompress
// (lz_compress.h:234-277), block_decompress_partial (block_compress.h:1749-1795) and the final
// per-block unshuffle (:2155) for element sizes T in {2,4,8}.
//
// Same lane mapping as the encoder: lane l produces elements 8l..8l+7 (row l>>1, half l&1).
// Row-to-row and half-to-half dependencies (delta rows, RLE rows that start with a repeat) are
// resolved with a 5-step warp scan over affine maps last = a*prev + c, a in {0,1} -- not a
// 16-step serial chain.
#pragma once
#include "sb_common.cuh"

namespace sb
{
	// bounds-aware 8-byte read: bytes at or after `lim` read as 0
	__device__ __forceinline__ void rd64_safe(const uint8_t* p, const uint8_t* lim, uint32_t& lo, uint32_t& hi)
	{
		if (p + 12 <= lim) {
			rd64_unaligned(p, lo, hi);
		}
		else {
			lo = hi = 0;
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				if (p + j < lim)
					lo |= (uint32_t)p[j] << (8 * j);
				if (p + 4 + j < lim)
					hi |= (uint32_t)p[4 + j] << (8 * j);
			}
		}
	}

	// spreads four `bits`-wide values from the low bits of a into the four bytes of the result
	__device__ __forceinline__ uint32_t unpack4(uint32_t a, uint32_t bits)
	{
		const uint32_t m = (1u << bits) - 1u;
		return (a & m) | ((a << (8u - bits)) & (m << 8)) | ((a << (16u - 2u * bits)) & (m << 16)) | ((a << (24u - 3u * bits)) & (m << 24));
	}

	// Expands 8 values of an RLE row half: bit k of m8 set = "repeats its predecessor".
	// vals: the non repeated bytes in order.  Leading repeats (count returned in `lead`) are left
	// as 0 and must be filled by the caller with the carry-in.
	__device__ __forceinline__ void rle_expand8(uint32_t m8, uint32_t vlo, uint32_t vhi, uint32_t& olo, uint32_t& ohi, uint32_t& lead)
	{
		const uint32_t keep = ~m8 & 0xFFu;
		lead = keep ? (uint32_t)(__ffs((int)keep) - 1) : 8u;
		uint32_t x = keep; // inclusive prefix count of kept bytes, one nibble per position
		x = (x | (x << 12)) & 0x000F000Fu;
		x = (x | (x << 6)) & 0x03030303u;
		x = (x | (x << 3)) & 0x11111111u;
		x += x << 4;
		x += x << 8;
		x += x << 16;
		// index of the source byte = count - 1; leading positions (count 0) are lifted to 0 first
		const uint32_t leadn = (lead >= 8u ? 0xFFFFFFFFu : ((1u << (4u * lead)) - 1u)) & 0x11111111u;
		const uint32_t idx = x + leadn - 0x11111111u;
		olo = __byte_perm(vlo, vhi, idx & 0xFFFFu);
		ohi = __byte_perm(vlo, vhi, idx >> 16);
		// clear the leading positions
		const uint32_t mlo = lead >= 4u ? 0xFFFFFFFFu : ((1u << (8u * lead)) - 1u);
		const uint32_t mhi = lead <= 4u ? 0u : (lead >= 8u ? 0xFFFFFFFFu : ((1u << (8u * (lead - 4u))) - 1u));
		olo &= ~mlo;
		ohi &= ~mhi;
	}

	// fills the first `lead` (0..8) bytes of (lo,hi) with byte c
	__device__ __forceinline__ void fill_lead8(uint32_t& lo, uint32_t& hi, uint32_t lead, uint32_t c)
	{
		const uint32_t mlo = lead >= 4u ? 0xFFFFFFFFu : ((1u << (8u * lead)) - 1u);
		const uint32_t mhi = lead <= 4u ? 0u : (lead >= 8u ? 0xFFFFFFFFu : ((1u << (8u * (lead - 4u))) - 1u));
		const uint32_t s = splat(c);
		lo = (lo & ~mlo) | (s & mlo);
		hi = (hi & ~mhi) | (s & mhi);
	}

	// Decodes one NORMAL / NORMAL_RLE plane starting at q.  On return (lo,hi) hold this lane's 8
	// bytes of the plane (rows >= lines are unspecified) and the consumed byte count is returned
	// (0xFFFFFFFF on truncated input).  `end` = end of the superblock payload, `lim` = end of the
	// readable buffer (>= end).
	__device__ __forceinline__ uint32_t decode_plane(const uint8_t* q, const uint8_t* end, const uint8_t* lim, uint32_t kind, uint32_t lines, int lane, uint32_t& lo, uint32_t& hi)
	{
		const int half = lane & 1;
		const uint32_t row = (uint32_t)lane >> 1;
		const bool live = row < lines;
		const uint32_t hb = (lines + 1u) >> 1;
		if (q + hb + (kind == KIND_NORMAL_RLE ? 2u : 0u) > end)
			return 0xFFFFFFFFu;
		const uint32_t h = live ? ((uint32_t)(q[row >> 1] >> (4u * (row & 1u))) & 15u) : 15u;
		const bool needmin = live && !(h == 6u || h == 7u || h == 15u);
		const bool is_rle = live && (h == 6u || h == 7u);

		// ---- mins
		uint32_t minv = 0, mins_len;
		if (kind == KIND_NORMAL_RLE) {
			// decode_rle_flat on the 16 mins with predecessor 0 (:2071-2084)
			const uint32_t mm = rd16(q + hb);
			const uint32_t nonrep = ~mm & 0xFFFFu;
			mins_len = 2u + __popc(nonrep);
			if (q + hb + mins_len > end)
				return 0xFFFFFFFFu;
			const uint32_t c = __popc(nonrep & ((2u << row) - 1u));
			minv = c ? q[hb + 2u + c - 1u] : 0u;
		}
		else {
			const uint32_t nm = __ballot_sync(FULL, needmin) & 0x55555555u;
			mins_len = __popc(nm);
			if (q + hb + mins_len > end)
				return 0xFFFFFFFFu;
			if (needmin)
				minv = q[hb + __popc(nm & lanemask_lt(lane & ~1))];
		}
		const uint8_t* rows = q + hb + mins_len;

		// ---- row payload sizes and offsets
		const uint32_t bits = h & 7u;
		uint32_t payload = !live ? 0u : (h == 15u ? 16u : (is_rle ? 0u : 2u * bits));
		uint32_t incl = half == 0 ? payload : 0u;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(FULL, incl, d);
			if (lane >= d)
				incl += t;
		}
		uint32_t rowoff = incl - payload;
		uint32_t consumed = __shfl_sync(FULL, incl, 31);
		uint32_t mask16 = 0;
		uint32_t rle_rows = __ballot_sync(FULL, is_rle) & 0x55555555u;
		if (rle_rows) {
			// RLE rows have data dependent sizes (2 + number of non repeats): resolve them in row order
			uint32_t acc = 0, extra = 0;
			while (rle_rows) {
				const int L = __ffs((int)rle_rows) - 1;
				rle_rows &= rle_rows - 1u;
				const uint32_t at = __shfl_sync(FULL, rowoff, L) + acc;
				const uint8_t* mp = rows + at;
				if (mp + 2 > end)
					return 0xFFFFFFFFu;
				const uint32_t m = rd16(mp);
				const uint32_t psz = 2u + __popc(~m & 0xFFFFu);
				if ((lane >> 1) == (L >> 1))
					mask16 = m;
				if (lane > L + 1)
					extra += psz;
				acc += psz;
			}
			rowoff += extra;
			consumed += acc;
		}
		if (rows + consumed > end)
			return 0xFFFFFFFFu;
		const uint8_t* rp = rows + rowoff;

		// ---- this lane's 8 bytes, before carries
		uint32_t a = 0, c = 0, lead = 0; // affine map of this lane: last = a*carry + c
		bool is_sum = false;             // values = carry + prefix sums of (lo,hi)
		if (!live) {
			lo = hi = 0;
		}
		else if (h == 15u) {
			rd64_safe(rp + 8 * half, lim, lo, hi);
			c = hi >> 24;
		}
		else if (is_rle) {
			const uint32_t m8 = (mask16 >> (8 * half)) & 0xFFu;
			const uint32_t skipn = half ? __popc(~mask16 & 0xFFu) : 0u;
			uint32_t vlo, vhi;
			rd64_safe(rp + 2 + skipn, lim, vlo, vhi);
			rle_expand8(m8, vlo, vhi, lo, hi, lead);
		}
		else {
			uint32_t vlo = 0, vhi = 0;
			if (bits) {
				uint32_t plo, phi;
				rd64_safe(rp + bits * half, lim, plo, phi);
				vlo = unpack4(plo, bits);
				vhi = unpack4(__funnelshift_r(plo, phi, 4u * bits), bits);
			}
			const uint32_t m4 = splat(minv);
			lo = __vadd4(vlo, m4);
			hi = __vadd4(vhi, m4);
			if (h < 8u)
				c = hi >> 24;
			else
				is_sum = true;
		}
		// delta-RLE rows: the expanded bytes are deltas; leading repeats of the high half repeat the
		// low half's last delta, those of the low half are 0 (the row restarts from 0, :1982)
		const bool any_drle = __any_sync(FULL, live && h == 6u);
		if (any_drle) {
			const uint32_t prevd = __shfl_up_sync(FULL, hi >> 24, 1);
			if (live && h == 6u) {
				// low half: leading bytes stay 0.  high half: low half's last delta (0 if it was all repeats too)
				if (half)
					fill_lead8(lo, hi, lead, prevd);
				is_sum = true;
				lead = 0;
			}
		}
		if (is_sum) {
			lo = prefix4(lo);
			hi = __vadd4(prefix4(hi), splat(lo >> 24));
			a = 1;
			c = hi >> 24;
		}
		else if (is_rle) { // h == 7
			a = (lead == 8u);
			c = a ? 0u : (hi >> 24);
		}

		// ---- carries across lanes: inclusive scan of the affine maps, then shift by one lane
		if (__any_sync(FULL, a != 0u || lead != 0u)) {
			uint32_t x = (c & 0xFFu) | (a << 8);
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t y = __shfl_up_sync(FULL, x, d);
				if (lane >= d) {
					// x := x o y  (apply y first)
					const uint32_t cy = (x & 0x100u) ? (y & 0xFFu) : 0u;
					x = ((x + cy) & 0xFFu) | (x & y & 0x100u);
				}
			}
			uint32_t carry = __shfl_up_sync(FULL, x, 1) & 0xFFu;
			if (lane == 0)
				carry = 0;
			if (is_sum) {
				const uint32_t s = splat(carry);
				lo = __vadd4(lo, s);
				hi = __vadd4(hi, s);
			}
			else if (lead)
				fill_lead8(lo, hi, lead, carry);
		}
		return hb + mins_len + consumed;
	}

	// LZ block decoder (lz_compress.h:234-277).  The item stream is parsed by lane 0 (inherently
	// sequential: item sizes are data dependent); every element's source is then resolved by
	// pointer jumping over the back-reference chains, and values are fetched in parallel.
	// scratch: per-warp shared memory, 256 x u16.  Returns consumed bytes, 0xFFFFFFFF on error.
	template<int T>
	__device__ __noinline__ uint32_t lz_decode_block(const uint8_t* p, const uint8_t* end, uint16_t* src_of, int lane, uint32_t (&w)[2 * T])
	{
		constexpr uint32_t B = T;
		// src_of[e]: stream offset of element e's raw bytes, or 0x8000 | position of the element it copies
		uint32_t consumed = 0xFFFFFFFFu;
		if (lane == 0) {
			const uint8_t* s = p;
			bool ok = true;
			for (uint32_t g = 0; g < 32 && ok; ++g) {
				if (s + 2 > end) {
					ok = false;
					break;
				}
				const uint32_t anchor = *s++;
				if (!anchor) {
					if (s + 8 * B > end) {
						ok = false;
						break;
					}
					for (uint32_t k = 0; k < 8; ++k)
						src_of[8 * g + k] = (uint16_t)(s - p + k * B);
					s += 8 * B;
					continue;
				}
				for (uint32_t k = 0; k < 8; ++k) {
					const uint32_t e = 8 * g + k;
					if ((anchor >> k) & 1u) {
						uint32_t off = *s & 127u;
						if (*s++ > 127u) {
							if (s == end) {
								ok = false;
								break;
							}
							off |= (uint32_t)(*s++) << 7;
						}
						if (off == 0 || off > e) {
							ok = false;
							break;
						}
						src_of[e] = (uint16_t)(0x8000u | (e - off));
					}
					else {
						if (s + B > end) {
							ok = false;
							break;
						}
						src_of[e] = (uint16_t)(s - p);
						s += B;
					}
				}
			}
			if (ok)
				consumed = (uint32_t)(s - p);
		}
		consumed = __shfl_sync(FULL, consumed, 0);
		if (consumed == 0xFFFFFFFFu)
			return consumed;
		__syncwarp();
		// pointer jumping: chains only point backwards, 8 rounds cover any chain of length < 256
		for (int r = 0; r < 8; ++r) {
			uint32_t nxt[8];
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				const uint32_t v = src_of[8 * lane + k];
				nxt[k] = (v & 0x8000u) ? src_of[v & 0x7FFFu] : v;
			}
			__syncwarp();
#pragma unroll
			for (int k = 0; k < 8; ++k)
				src_of[8 * lane + k] = (uint16_t)nxt[k];
			__syncwarp();
		}
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const uint8_t* s = p + src_of[8 * lane + k];
#pragma unroll
			for (uint32_t i = 0; i < B / 4; ++i) {
				uint32_t v = 0;
#pragma unroll
				for (int j = 0; j < 4; ++j)
					v |= (uint32_t)s[4 * i + j] << (8 * j);
				w[(k * (B / 4) + i) % (2 * T)] = v;
			}
		}
		__syncwarp();
		return consumed;
	}

	// One full block: p -> 256 elements at `out` (16-byte aligned).  Returns consumed bytes or
	// 0xFFFFFFFF (truncated / corrupt input; nothing out of bounds is ever read or written).
	template<int T>
	__device__ __forceinline__ uint32_t decode_block(const uint8_t* p, const uint8_t* end, const uint8_t* lim, uint8_t* out, uint16_t* lz_scratch, int lane)
	{
		constexpr uint32_t HS = (T + 1) / 2;
		if (p + HS >= end) // block_compress.h:2114-2116
			return 0xFFFFFFFFu;
		uint32_t w[2 * T];
		const uint32_t marker = p[0];
		if (marker == (uint32_t)MARK_COPY) { // :2118-2123 (time limited streams)
			if (p + 1 + T * 256 > end)
				return 0xFFFFFFFFu;
			const uint8_t* s = p + 1 + (size_t)lane * 8 * T;
#pragma unroll
			for (int i = 0; i < T; ++i)
				rd64_safe(s + 8 * i, lim, w[2 * i], w[2 * i + 1]);
			store_lane_words<T>(out, lane, w);
			return 1u + T * 256u;
		}
		if (marker == (uint32_t)MARK_LZ) { // :2124-2130
			if ((T % 4) != 0)
				return 0xFFFFFFFFu;
			const uint32_t r = lz_decode_block<T>(p + 1, end, lz_scratch, lane, w);
			if (r == 0xFFFFFFFFu)
				return r;
			store_lane_words<T>(out, lane, w);
			return r + 1u;
		}
		uint32_t kinds = 0;
#pragma unroll
		for (uint32_t i = 0; i < HS; ++i)
			kinds |= (uint32_t)p[i] << (8 * i);
		uint32_t lo[T], hi[T];
		const uint8_t* q = p + HS;
#pragma unroll
		for (int pl = 0; pl < T; ++pl) {
			const uint32_t kind = (kinds >> (4 * pl)) & 15u;
			if (kind == KIND_SAME) {
				if (q >= end)
					return 0xFFFFFFFFu;
				lo[pl] = hi[pl] = splat(*q);
				q += 1;
			}
			else if (kind == KIND_RAW) {
				if (q + 256 > end)
					return 0xFFFFFFFFu;
				rd64_safe(q + 8 * lane, lim, lo[pl], hi[pl]);
				q += 256;
			}
			else if (kind == KIND_NORMAL || kind == KIND_NORMAL_RLE) {
				const uint32_t r = decode_plane(q, end, lim, kind, 16, lane, lo[pl], hi[pl]);
				if (r == 0xFFFFFFFFu)
					return r;
				q += r;
			}
			else
				return 0xFFFFFFFFu;
		}
		planes_to_words<T>(lo, hi, w);
		store_lane_words<T>(out, lane, w);
		return (uint32_t)(q - p);
	}

	// Partial tail block (after the 254 marker): `bytes` < T*256 decoded bytes at `out`.
	template<int T>
	__device__ __noinline__ uint32_t decode_partial_block(const uint8_t* p, const uint8_t* end, const uint8_t* lim, uint8_t* out, uint32_t bytes, int lane)
	{
		constexpr uint32_t HS = (T + 1) / 2;
		const uint32_t line = 16u * T, lines = bytes / line;
		const uint8_t* q = p;
		if (lines) {
			q += HS;
			if (q >= end) // :1765
				return 0xFFFFFFFFu;
			uint32_t kinds = 0;
#pragma unroll
			for (uint32_t i = 0; i < HS; ++i)
				kinds |= (uint32_t)p[i] << (8 * i);
			uint32_t lo[T], hi[T];
#pragma unroll
			for (int pl = 0; pl < T; ++pl) {
				const uint32_t kind = (kinds >> (4 * pl)) & 15u;
				if (kind == KIND_SAME) {
					if (q >= end)
						return 0xFFFFFFFFu;
					lo[pl] = hi[pl] = splat(*q);
					q += 1;
				}
				else if (kind == KIND_NORMAL) {
					const uint32_t r = decode_plane(q, end, lim, kind, lines, lane, lo[pl], hi[pl]);
					if (r == 0xFFFFFFFFu)
						return r;
					q += r;
				}
				else
					return 0xFFFFFFFFu; // :1779
			}
			uint32_t w[2 * T];
			planes_to_words<T>(lo, hi, w);
			// only the first lines*16 elements are real
			const uint32_t valid = lines * line;
#pragma unroll
			for (int i = 0; i < 2 * T; ++i) {
				const uint32_t at = (uint32_t)lane * 8u * T + 4u * i;
				if (at + 4 <= valid)
					*reinterpret_cast<uint32_t*>(out + at) = w[i];
			}
		}
		const uint32_t rem = bytes - lines * line;
		if (rem) {
			if (q + rem > end)
				return 0xFFFFFFFFu;
			for (uint32_t i = lane; i < rem; i += 32)
				out[lines * line + i] = q[i];
			q += rem;
		}
		return (uint32_t)(q - p);
	}
}
