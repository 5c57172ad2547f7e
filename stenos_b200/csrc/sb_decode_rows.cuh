// sb_decode_rows.cuh -- the level-1 block decoder of the fast path: one HALF-WARP per superblock,
// one lane per 16-element ROW, two superblocks per warp.
//
// Same bytes as sb_decode.cuh (which stays for LZ blocks, COPY blocks, the partial tail block, COPY
// superblocks and any block that lies within one worst-case block of the end of the compressed
// buffer), i.e. the reference's block_decompress_sse (block_compress.h:2088-2175) with
// decolde_line_flat (:1970-2033), read_16_bits (:1451-1486), decode_rle_flat (:1939-1968),
// prefix_sum_16 (:1897-1904) and the final unshuffle (:2155), laid out so that everything the
// reference does per 16-byte row is lane-local:
//
//   lane l: superblock stream (l >> 4) of the warp, row r = l & 15 of that stream's current block.
//   The block stream is not seekable, so a stream walks its blocks in order; but a
//   row's header nibble, min, payload offset (a 4-step scan over 16 lanes; RLE rows, whose length is
//   data dependent, are resolved in row order), unpacking (two IMAD + three LOP3 per four values),
//   in-row prefix sum (16-bit lanes: IMAD by 0x00010001) and the final byte transpose need no other
//   lane.  Row-to-row carries (delta rows, RLE rows starting with a repeat) are one 4-step scan
//   over affine maps last = a * prev + c, a in {0, 1}.  Output: 16 * T contiguous bytes per lane,
//   T STG.E.128.
//
// Memory safety without per-access checks: a block is decoded here only when a worst-case block
// (HS + T * 312 bytes) plus the over-read of the unaligned 16-byte loads fits before the end of the
// compressed buffer; what it consumed is checked against the superblock end afterwards.
#pragma once
#include "sb_common.cuh"
#include "sb_decode.cuh"
#include "sb_encode_rows.cuh"

namespace sb
{
#ifndef DECODE_AHEAD_BYTES
#define DECODE_AHEAD_BYTES 768u
#endif
#ifndef DECODE_AHEAD_LINES
#define DECODE_AHEAD_LINES 2
#endif
	__device__ __forceinline__ uint32_t bitsel(uint32_t k, uint32_t x, uint32_t y) { return (x & k) | (y & ~k); }

	// four `bits`-wide values in the low bits of a -> one per byte.  P2 = 1 << 2 * (8 - bits), P1 = 1 << (8 - bits),
	// m4 = ((1 << bits) - 1) * 0x01010101.  (read_16_bits, block_compress.h:1451-1486)
	__device__ __forceinline__ uint32_t unpack4f(uint32_t a, uint32_t P2, uint32_t P1, uint32_t m4)
	{
		const uint32_t t = bitsel(0x0000FFFFu, a, a * P2);
		return bitsel(0x00FF00FFu, t, t * P1) & m4;
	}

	// bytes [b, b + 8) of the 16 bytes v[0..3], b in 0..8
	__device__ __forceinline__ void window8(const uint32_t (&v)[4], uint32_t b, uint32_t& lo, uint32_t& hi)
	{
		const uint32_t wi = b >> 2;
		const uint32_t x0 = wi == 0u ? v[0] : (wi == 1u ? v[1] : v[2]);
		const uint32_t x1 = wi == 0u ? v[1] : (wi == 1u ? v[2] : v[3]);
		const uint32_t x2 = wi == 0u ? v[2] : v[3];
		const uint32_t sh = (b & 3u) * 8u;
		lo = __funnelshift_r(x0, x1, sh);
		hi = __funnelshift_r(x1, x2, sh);
	}

	// 16 bytes at an arbitrary address (5 aligned words; the caller guarantees 20 readable bytes)
	__device__ __forceinline__ void load16_unaligned(const uint8_t* p, uint32_t (&v)[4])
	{
		const uintptr_t a = reinterpret_cast<uintptr_t>(p);
		const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
		const uint32_t sh = (uint32_t)(a & 3u) * 8u;
		const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
		v[0] = __funnelshift_r(w0, w1, sh);
		v[1] = __funnelshift_r(w1, w2, sh);
		v[2] = __funnelshift_r(w2, w3, sh);
		v[3] = __funnelshift_r(w3, w4, sh);
	}

	// inclusive prefix sums (mod 256) of the 16 bytes x[0..3], each increased by m, on top of carry c:
	// out byte k = c + sum_{i <= k} (x_i + m).  16-bit lanes: no byte-wise adds, one IMAD per word.
	__device__ __forceinline__ void prefix16(const uint32_t (&x)[4], uint32_t m, uint32_t c, uint32_t (&out)[4])
	{
		const uint32_t M11 = m * 0x00010001u, M22 = m * 0x00020002u;
		uint32_t c2 = c * 0x00010001u;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const uint32_t e = x[j] & 0x00FF00FFu;         // [b0, b2]
			const uint32_t o = __byte_perm(x[j], 0u, 0x4341); // [b1, b3]
			const uint32_t s = e + o + M22;
			const uint32_t PS = s * 0x00010001u + c2; // [P1, P3]
			const uint32_t PE = PS - o - M11;         // [P0, P2]
			out[j] = __byte_perm(PE, PS, 0x6240);
			c2 = __byte_perm(PS, 0u, 0x3232);
		}
	}

	// fills the first `lead` (0..16) bytes of o[0..3] with byte c
	__device__ __forceinline__ void fill_lead16(uint32_t (&o)[4], uint32_t lead, uint32_t c)
	{
		fill_lead8(o[0], o[1], min(lead, 8u), c);
		fill_lead8(o[2], o[3], lead > 8u ? lead - 8u : 0u, c);
	}

	// One NORMAL / NORMAL_RLE plane (block_compress.h:2046-2084) at q, for the lanes with `live` set (all 16 lanes of
	// a half-warp or none); out = the 16 bytes of the lane's row.  Returns the bytes consumed (uniform over the
	// half-warp; 0 for halves that are not live).  Must be called by the whole warp.
	__device__ __forceinline__ uint32_t decode_plane_row(const uint8_t* q, uint32_t kind, bool live, int r, int hsh, uint32_t (&out)[4])
	{
		const uint32_t below = (1u << r) - 1u;
		const uint32_t h = live ? ((uint32_t)(q[r >> 1] >> (4 * (r & 1))) & 15u) : 15u;
		const bool nomin = (0x80C0u >> h) & 1u; // headers 6, 7, 15 carry no min
		const bool is_rle = live && (h | 1u) == 7u;
		const uint32_t bits = h & 7u;

		// ---- mins: one byte per row that uses one, or RLE coded (all 16) when the plane kind says so (:2071-2084)
		uint32_t minv = 0, mins_len;
		const uint32_t nmb = (__ballot_sync(FULL, live && !nomin) >> hsh) & 0xFFFFu;
		if (kind == (uint32_t)KIND_NORMAL_RLE && live) {
			const uint32_t nonrep = ~rd16(q + 8) & 0xFFFFu;
			mins_len = 2u + __popc(nonrep);
			const uint32_t c = __popc(nonrep & ((2u << r) - 1u));
			minv = c ? q[8u + 2u + c - 1u] : 0u;
		}
		else {
			mins_len = __popc(nmb);
			if (live && !nomin)
				minv = q[8u + __popc(nmb & below)];
		}
		const uint8_t* rows = q + 8u + mins_len;

		// ---- row payload offsets: a scan over the 16 rows; RLE rows (2 + number of non repeats) resolved in row order
		const uint32_t pay = !live ? 0u : (h == 15u ? 16u : (is_rle ? 0u : 2u * bits));
		uint32_t incl = pay;
#pragma unroll
		for (int d = 1; d < 16; d <<= 1) {
			const uint32_t t = __shfl_up_sync(FULL, incl, d, 16);
			if (r >= d)
				incl += t;
		}
		uint32_t rowoff = incl - pay;
		uint32_t consumed = __shfl_sync(FULL, incl, 15, 16);
		uint32_t mask16 = 0;
		uint32_t rle_rows = (__ballot_sync(FULL, is_rle) >> hsh) & 0xFFFFu;
		if (__any_sync(FULL, rle_rows != 0u)) {
			uint32_t acc = 0, extra = 0;
			while (__any_sync(FULL, rle_rows != 0u)) {
				const bool on = rle_rows != 0u;
				const int L = on ? (__ffs((int)rle_rows) - 1) : 0;
				rle_rows &= rle_rows - 1u;
				const uint32_t at = __shfl_sync(FULL, rowoff, L, 16) + acc;
				if (on) {
					const uint32_t m = rd16(rows + at);
					const uint32_t psz = 2u + __popc(~m & 0xFFFFu);
					if (r == L)
						mask16 = m;
					if (r > L)
						extra += psz;
					acc += psz;
				}
			}
			rowoff += extra;
			consumed += acc;
		}

		// ---- the row
		uint32_t v[4] = { 0u, 0u, 0u, 0u };
		if (live && pay + (is_rle ? 1u : 0u) != 0u)
			load16_unaligned(rows + rowoff + (is_rle ? 2u : 0u), v);
		uint32_t a = 0, c = 0, lead = 0;
		uint32_t x[4] = { 0u, 0u, 0u, 0u }; // rows that end in a prefix sum: the 16 summands (without the min)
		bool is_sum = false;
		uint32_t summin = 0;
		if (!live) {
			out[0] = out[1] = out[2] = out[3] = 0u;
		}
		else if (h == 15u) { // raw row
			out[0] = v[0];
			out[1] = v[1];
			out[2] = v[2];
			out[3] = v[3];
			c = v[3] >> 24;
		}
		else if (is_rle) {
			// [mask:2][non repeated bytes] (decode_rle_flat, :1939-1968); header 6: the bytes are deltas and the row
			// restarts from delta 0 (:1982)
			uint32_t lead_lo, lead_hi, s0, s1;
			rle_expand8(mask16 & 0xFFu, v[0], v[1], x[0], x[1], lead_lo);
			window8(v, __popc(~mask16 & 0xFFu), s0, s1);
			rle_expand8(mask16 >> 8, s0, s1, x[2], x[3], lead_hi);
			if (h == 6u || lead_lo < 8u)
				fill_lead8(x[2], x[3], lead_hi, x[1] >> 24);
			if (h == 6u)
				is_sum = true;
			else {
				lead = lead_lo < 8u ? lead_lo : 8u + lead_hi;
				out[0] = x[0];
				out[1] = x[1];
				out[2] = x[2];
				out[3] = x[3];
				a = lead == 16u;
				c = a ? 0u : (x[3] >> 24);
			}
		}
		else {
			// bit packed: two groups of 8 values, `bits` bytes each (:1451-1486); header >= 8: deltas
			if (bits) {
				const uint32_t s = 8u - bits;
				const uint32_t P1 = 1u << s, P2 = P1 * P1, m4 = (0xFFu >> s) * 0x01010101u;
				const bool up = bits >= 4u;
				const uint32_t y0 = up ? v[1] : v[0], y1 = up ? v[2] : v[1], y2 = up ? v[3] : v[2];
				const uint32_t sh = (bits & 3u) * 8u;
				const uint32_t g1l = __funnelshift_r(y0, y1, sh), g1h = __funnelshift_r(y1, y2, sh);
				x[0] = unpack4f(v[0], P2, P1, m4);
				x[1] = unpack4f(__funnelshift_r(v[0], v[1], 4u * bits), P2, P1, m4);
				x[2] = unpack4f(g1l, P2, P1, m4);
				x[3] = unpack4f(__funnelshift_r(g1l, g1h, 4u * bits), P2, P1, m4);
			}
			if (h >= 8u) {
				is_sum = true;
				summin = minv;
			}
			else {
				// value = packed + min per byte; packed < 64, so adding the low 7 bits of the min cannot carry
				const uint32_t m4v = splat(minv);
				const uint32_t m7 = m4v & 0x7F7F7F7Fu, mh = m4v & 0x80808080u;
#pragma unroll
				for (int j = 0; j < 4; ++j)
					out[j] = (x[j] + m7) ^ mh;
				c = out[3] >> 24;
			}
		}
		if (is_sum) {
			// the row's total: its affine map is last = prev + total
			uint32_t tot = 16u * summin;
#pragma unroll
			for (int j = 0; j < 4; ++j)
				tot = sad4_acc(x[j], 0u, tot);
			a = 1;
			c = tot & 0xFFu;
		}

		// ---- carries across the rows of the block: inclusive scan of the affine maps, shifted by one row
		uint32_t carry = 0;
		if (__any_sync(FULL, a != 0u || lead != 0u)) {
			uint32_t z = (c & 0xFFu) | (a << 8);
#pragma unroll
			for (int d = 1; d < 16; d <<= 1) {
				const uint32_t y = __shfl_up_sync(FULL, z, d, 16);
				if (r >= d) {
					const uint32_t cy = (z & 0x100u) ? (y & 0xFFu) : 0u;
					z = ((z + cy) & 0xFFu) | (z & y & 0x100u);
				}
			}
			carry = __shfl_up_sync(FULL, z, 1, 16) & 0xFFu;
			if (r == 0)
				carry = 0; // the byte before the block is 0
		}
		if (is_sum)
			prefix16(x, summin, carry, out);
		else if (lead)
			fill_lead16(out, lead, carry);
		return live ? 8u + mins_len + consumed : 0u;
	}

	// the lane's row of every plane -> 16 elements at block + r * 16 * T (inverse of load_row_planes)
	template<int T>
	__device__ __forceinline__ void store_row_planes(uint8_t* __restrict__ block, int r, const uint32_t (&pw)[T][4])
	{
		uint32_t e[4 * T];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if constexpr (T == 2) {
				e[2 * j] = __byte_perm(pw[0][j], pw[1][j], 0x5140);
				e[2 * j + 1] = __byte_perm(pw[0][j], pw[1][j], 0x7362);
			}
			else if constexpr (T == 4) {
				transpose4(pw[0][j], pw[1][j], pw[2][j], pw[3][j], e[4 * j], e[4 * j + 1], e[4 * j + 2], e[4 * j + 3]);
			}
			else {
				transpose4(pw[0][j], pw[1][j], pw[2][j], pw[3][j], e[8 * j], e[8 * j + 2], e[8 * j + 4], e[8 * j + 6]);
				transpose4(pw[4][j], pw[5][j], pw[6][j], pw[7][j], e[8 * j + 1], e[8 * j + 3], e[8 * j + 5], e[8 * j + 7]);
			}
		}
		uint4* dst = reinterpret_cast<uint4*>(block + (size_t)r * 16 * T);
#if !defined(DECODE_STORE_128)
		if ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
			// whole 32-byte sectors per instruction (STG.E.256): the lane's 16 * T bytes are contiguous
#pragma unroll
			for (int i = 0; i < T; i += 2)
				st_global_256(dst + i, make_uint4(e[4 * i], e[4 * i + 1], e[4 * i + 2], e[4 * i + 3]), make_uint4(e[4 * i + 4], e[4 * i + 5], e[4 * i + 6], e[4 * i + 7]));
			return;
		}
#endif
#pragma unroll
		for (int i = 0; i < T; ++i) {
			const uint4 v = make_uint4(e[4 * i], e[4 * i + 1], e[4 * i + 2], e[4 * i + 3]);
#if defined(DECODE_STORE_CS) && !defined(STENOS_EMU)
			__stcs(dst + i, v); // the output is written once and never read here: keep it out of the way of the stream's lines
#elif defined(DECODE_STORE_CG) && !defined(STENOS_EMU)
			__stcg(dst + i, v);
#else
			dst[i] = v;
#endif
		}
	}

	// worst case of a plane coded block: kinds + T planes of 8 header bytes, 16 mins and 16 rows of 18 bytes
	template<int T>
	struct WorstBlock
	{
		static constexpr uint32_t READ = (T + 1) / 2 + T * 312u + 24u;
	};

	// The affine maps last = a * prev + c (a in {0, 1}) of TWO planes, one per 16-bit lane (c in bits 0..7, a in bit 8):
	// x := x o y (apply y first), lane wise, without any carry between the lanes
	__device__ __forceinline__ uint32_t compose_maps2(uint32_t x, uint32_t y)
	{
		const uint32_t m = (x >> 8) & 0x00010001u;
		const uint32_t sel = (m << 8) - m; // 0x00FF in the lanes whose a is 1
		return (((x & 0x00FF00FFu) + (y & sel)) & 0x00FF00FFu) | (x & y & 0x01000100u);
	}

	// One full plane-coded block per half-warp: p -> 256 elements at out.  live: the lane's half has a block here.
	// Returns the bytes consumed (uniform over the half-warp), 0xFFFFFFFF for an invalid plane kind.
	template<int T>
	__device__ __forceinline__ uint32_t decode_block_rows(const uint8_t* p, bool live, int r, int hsh, uint8_t* out)
	{
		constexpr uint32_t HS = (T + 1) / 2;
		uint32_t kinds = 0;
		if (live) {
#pragma unroll
			for (uint32_t i = 0; i < HS; ++i)
				kinds |= (uint32_t)p[i] << (8 * i);
		}
		uint32_t pw[T][4];
		const uint8_t* q = p + HS;
		bool bad = false;
#pragma unroll
		for (int pl = 0; pl < T; ++pl) {
			const uint32_t kind = (kinds >> (4 * pl)) & 15u;
			const bool normal = live && (kind == (uint32_t)KIND_NORMAL || kind == (uint32_t)KIND_NORMAL_RLE);
			uint32_t c = 0;
			if (__any_sync(FULL, normal))
				c = decode_plane_row(q, kind, normal, r, hsh, pw[pl]);
			if (live && !normal) {
				if (kind == (uint32_t)KIND_SAME) {
					pw[pl][0] = pw[pl][1] = pw[pl][2] = pw[pl][3] = splat(*q);
					c = 1;
				}
				else if (kind == (uint32_t)KIND_RAW) {
					load16_unaligned(q + 16 * r, pw[pl]);
					c = 256;
				}
				else {
					bad = true;
					pw[pl][0] = pw[pl][1] = pw[pl][2] = pw[pl][3] = 0u;
				}
			}
			q += c;
		}
		if (live && !bad)
			store_row_planes<T>(out, r, pw);
		return bad ? 0xFFFFFFFFu : (uint32_t)(q - p);
	}

	// Two superblocks per warp: the lanes of half h decode the superblock [code][csize:3][payload] at offset `at`
	// (of that half) into dsize bytes at `out`.  valid: the half has a superblock.  Returns the half's device error bits.
	template<int T>
	__device__ __forceinline__ uint32_t decode_superblock_pair(const uint8_t* src, uint64_t src_size, uint64_t at, uint32_t dsize, uint8_t* out, bool valid,
								   bool allow_zstd_tail, uint16_t* lz_scratch, int lane, void* scratch_word = nullptr)
	{
		constexpr uint32_t BLOCK = T * 256u;
		constexpr uint32_t HS = (T + 1) / 2;
		const int half = lane >> 4, r = lane & 15, hsh = 16 * half;
		const uint8_t* lim = src + src_size;
		uint32_t err = 0;
		bool run = false;
		const uint8_t* q = src;
		const uint8_t* end = src;
		uint32_t nleft = 0, rem = 0;

		// ---- superblock headers (stenos.cpp:1126-1134, :681-753); everything but a block stream goes through the
		// warp-wide path of sb_kernels.cuh (COPY superblocks, the Zstd tail, errors)
#pragma unroll 1
		for (int hh = 0; hh < 2; ++hh) {
			if (!__shfl_sync(FULL, (int)valid, 16 * hh))
				continue;
			const uint64_t at_h = __shfl_sync(FULL, (unsigned long long)at, 16 * hh);
			const uint32_t dsize_h = __shfl_sync(FULL, dsize, 16 * hh);
			bool stream = false;
			if (at_h + 4 <= src_size) {
				const uint8_t* p = src + at_h;
				const uint32_t code = p[0], csize = rd24(p + 1);
				stream = code == (uint32_t)CODE_BLOCK && at_h + 4 + csize <= src_size && csize != 0u;
				if (stream && half == hh) {
					run = true;
					q = p + 4;
					end = q + csize;
					nleft = dsize / BLOCK;
					rem = dsize - nleft * BLOCK;
				}
			}
			if (!stream) {
				uint8_t* out_h = reinterpret_cast<uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)out, 16 * hh));
				const bool allow_h = __shfl_sync(FULL, (int)allow_zstd_tail, 16 * hh);
				const uint32_t e = decode_superblock_warp<T>(src, src_size, at_h, dsize_h, out_h, lz_scratch, lane, allow_h);
				if (half == hh)
					err = e;
			}
		}

		// ---- the blocks of both streams, in step
		uint8_t* o = out;
		while (__any_sync(FULL, run && nleft != 0u)) {
			const bool act = run && nleft != 0u;
			// block_compress.h:2114-2116: the kinds must be followed by data; and a worst case block must be readable
			bool fast = act && q + HS < end && q + WorstBlock<T>::READ <= lim;
			const uint32_t marker = fast ? (uint32_t)q[0] : 0u;
			const bool special = act && (!fast || marker >= (uint32_t)MARK_COPY);
			fast = fast && !special;
			uint32_t consumed = 0;
			if (__any_sync(FULL, special)) {
				// LZ blocks, COPY blocks, blocks near the end of the buffer: the checked warp-wide decoder, one stream at a time
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					if (!__shfl_sync(FULL, (int)special, 16 * hh))
						continue;
					const uint8_t* q_h = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)q, 16 * hh));
					const uint8_t* end_h = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)end, 16 * hh));
					uint8_t* o_h = reinterpret_cast<uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)o, 16 * hh));
					const uint32_t c = decode_block<T>(q_h, end_h, lim, o_h, lz_scratch, lane);
					if (half == hh)
						consumed = c;
				}
			}
			if (__any_sync(FULL, fast)) {
				// the stream is read front to back: keep the lines a few blocks ahead on their way to L1
				if (fast && r < DECODE_AHEAD_LINES) {
					const uint8_t* ahead = reinterpret_cast<const uint8_t*>((reinterpret_cast<uintptr_t>(q) + DECODE_AHEAD_BYTES + 128u * r) & ~(uintptr_t)127);
					if (ahead + 128 <= lim) {
#ifdef DECODE_TOUCH_SYNC
						touch_l1(ahead);
#else
						if (scratch_word)
							prefetch_l1_async(ahead, scratch_word);
						else
							touch_l1(ahead);
#endif
					}
				}
				const uint32_t c = decode_block_rows<T>(q, fast, r, hsh, o);
				if (fast)
					consumed = c;
			}
			if (act) {
				if (consumed == 0xFFFFFFFFu || consumed > (uint32_t)(end - q)) {
					err = DEV_ERR_INVALID_INPUT;
					run = false;
				}
				else {
					q += consumed;
					o += BLOCK;
					--nleft;
				}
			}
		}

		// ---- partial tail blocks (block_compress.h:2158-2172)
#pragma unroll 1
		for (int hh = 0; hh < 2; ++hh) {
			if (!__shfl_sync(FULL, (int)(run && rem != 0u), 16 * hh))
				continue;
			const uint8_t* q_h = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)q, 16 * hh));
			const uint8_t* end_h = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)end, 16 * hh));
			uint8_t* o_h = reinterpret_cast<uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)o, 16 * hh));
			const uint32_t rem_h = __shfl_sync(FULL, rem, 16 * hh);
			uint32_t e = 0;
			if (q_h >= end_h || *q_h != (uint8_t)MARK_PARTIAL)
				e = DEV_ERR_INVALID_INPUT;
			else if (decode_partial_block<T>(q_h + 1, end_h, lim, o_h, rem_h, lane) == 0xFFFFFFFFu)
				e = DEV_ERR_INVALID_INPUT;
			if (half == hh && e)
				err = e;
		}
		return err;
	}

#ifndef DECODE2_WARPS_PER_CTA
#define DECODE2_WARPS_PER_CTA 4
#endif
#ifndef DECODE2_MIN_CTAS_T2
#define DECODE2_MIN_CTAS_T2 7 // 72 registers without spills: 28 warps per SM (1.28 -> 0.95 ms per GiB of int16)
#endif
#ifndef DECODE2_MIN_CTAS_T8
#define DECODE2_MIN_CTAS_T8 7 // spills, but 56 streams per SM hold 1 GiB in one round (0.66 -> 0.61 ms); T=4 gains nothing
#endif
#ifndef DECODE2_MIN_CTAS
#define DECODE2_MIN_CTAS 6
#endif
	constexpr int DECODE2_WARPS = DECODE2_WARPS_PER_CTA;

	// frame decoder: persistent warps; a warp takes the next two superblocks (one per half-warp) by ticket, so
	// streams of very different cost (constant data next to noisy data) do not leave SMs idle behind a slow CTA
	template<int T>
	__global__ void __launch_bounds__(DECODE2_WARPS * 32, (T == 2 ? DECODE2_MIN_CTAS_T2 : T == 8 ? DECODE2_MIN_CTAS_T8 : DECODE2_MIN_CTAS)) decode_pairs_kernel(DecodeParams P)
	{
		STENOS_DYN_SMEM(uint8_t, smem);
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		uint16_t* lz_scratch = reinterpret_cast<uint16_t*>(smem) + 256 * warp;
		void* scratch_word = smem + DECODE2_WARPS * 512 + 4 * threadIdx.x; // target of the asynchronous L1 prefetches, never read
		for (;;) {
			uint32_t w = 0;
			if (lane == 0)
				w = (uint32_t)atomicAdd(P.ticket, 1ull);
			w = __shfl_sync(FULL, w, 0);
			if (2ull * w >= (unsigned long long)P.n_sb)
				break;
			const uint32_t i = 2u * w + (uint32_t)(lane >> 4);
			const bool valid = i < P.n_sb;
			const uint32_t s = P.first_sb + (valid ? i : 0u);
			const uint64_t doff = (uint64_t)s * P.sb_bytes;
			const uint32_t dsize = (uint32_t)min((uint64_t)P.sb_bytes, P.total - doff); // remainder 0 = full superblock (appendix C1)
			const bool last = (doff + dsize == P.total);
			const uint32_t bad = decode_superblock_pair<T>(P.src, P.src_size, P.sb_offsets[s], dsize, P.dst + (doff - P.dst_origin), valid, P.skip_zstd_tail && last,
								       lz_scratch, lane, scratch_word);
			if (bad && valid && (lane & 15) == 0)
				atomicOr(&P.result[1], (unsigned long long)bad);
		}
	}

	// stenos::cvector random access: one half-warp per requested bucket (cvector.hpp:2879 -> :1862-1883).
	// A bucket is three dependent random accesses away (its id, its offset, its bytes), so persistent warps run a
	// four stage software pipeline over their bucket pairs -- id of pair n + 3, offset of pair n + 2, the lines of
	// pair n + 1 on their way to L1, pair n decoded -- and no stage waits for the latency of the one before.
#ifndef GATHER_MIN_CTAS
#define GATHER_MIN_CTAS 5 // 96 registers, no spills: measured best of 3..7
#endif
	template<int T>
	__global__ void __launch_bounds__(DECODE2_WARPS * 32, GATHER_MIN_CTAS) gather_pairs_kernel(GatherParams P)
	{
		STENOS_DYN_SMEM(uint8_t, smem);
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, r = lane & 15;
		uint16_t* lz_scratch = reinterpret_cast<uint16_t*>(smem) + 256 * warp;
		void* scratch_word = smem + DECODE2_WARPS * 512 + 4 * threadIdx.x;
		const uint32_t n_warps = gridDim.x * DECODE2_WARPS;
		const uint32_t w0 = blockIdx.x * DECODE2_WARPS + warp;
		const uint32_t n_pairs = (P.n + 1u) / 2u;
		const uint8_t* lim = P.src + P.src_size;
		// pipeline registers of this half: request index (>= P.n: nothing), bucket id, offset of its header
		uint32_t i1 = 0xFFFFFFFFu, i2 = 0xFFFFFFFFu, i3 = 0xFFFFFFFFu, id1 = 0, id2 = 0, id3 = 0;
		unsigned long long off1 = 0, off2 = 0;
		auto fetch_id = [&](uint32_t pair, uint32_t& i, uint32_t& id) {
			i = pair < n_pairs ? 2u * pair + (uint32_t)half : 0xFFFFFFFFu;
			if (i >= P.n)
				i = 0xFFFFFFFFu;
			id = i != 0xFFFFFFFFu ? P.ids[i] : 0u;
		};
		auto fetch_off = [&](uint32_t i, uint32_t id) -> unsigned long long { return (i != 0xFFFFFFFFu && id < P.n_buckets) ? P.sb_offsets[id] : 0ull; };
		auto touch = [&](uint32_t i, uint32_t id, unsigned long long off) {
			if (i != 0xFFFFFFFFu && id < P.n_buckets && r < 2) {
				// a bucket of one block: its header and the first lines of its stream (the rest is asked for on demand)
				const uint8_t* a = reinterpret_cast<const uint8_t*>((reinterpret_cast<uintptr_t>(P.src + off) & ~(uintptr_t)127) + 128u * (uint32_t)r);
				if (a >= P.src && a + 128 <= lim)
					prefetch_l1_async(a, scratch_word);
			}
		};
		fetch_id(w0, i1, id1);
		fetch_id(w0 + n_warps, i2, id2);
		fetch_id(w0 + 2u * n_warps, i3, id3);
		off1 = fetch_off(i1, id1);
		off2 = fetch_off(i2, id2);
		touch(i1, id1, off1);
		for (uint32_t pair = w0; pair < n_pairs; pair += n_warps) {
			const uint32_t i = i1, id = id1;
			const unsigned long long off = off1;
			// advance the pipeline: the loads issued here are consumed one iteration later
			i1 = i2;
			id1 = id2;
			off1 = off2;
			i2 = i3;
			id2 = id3;
			touch(i1, id1, off1);
			off2 = fetch_off(i2, id2);
			fetch_id(pair + 3u * n_warps, i3, id3);

			bool valid = i != 0xFFFFFFFFu;
			uint32_t bad = 0;
			if (valid && id >= P.n_buckets) {
				bad = DEV_ERR_INVALID_INPUT;
				valid = false;
			}
			const uint64_t doff = (uint64_t)id * P.bucket_bytes;
			const uint32_t dsize = valid ? (uint32_t)min((uint64_t)P.bucket_bytes, P.total - doff) : 0u;
			uint8_t* out = P.dst + (uint64_t)(valid ? i : 0u) * P.bucket_bytes;
			// Buckets of ONE plane-coded block (cvector's default): straight to the row decoder, without the superblock
			// machinery of decode_superblock_pair (header loop, block loop, special-block and tail paths: a third of the
			// instructions of a one-block bucket).  Anything else -- COPY / LZ / marker blocks, a bucket near the end of
			// the buffer, larger or partial buckets, a corrupt header -- takes the general path for the pair.
			bool fast = false;
			const uint8_t* q = P.src;
			uint32_t csize = 0;
			if (valid && dsize == (uint32_t)T * 256u && off + 4 <= P.src_size) {
				const uint8_t* p = P.src + off;
				csize = rd24(p + 1);
				q = p + 4;
				fast = p[0] == (uint8_t)CODE_BLOCK && off + 4 + csize <= P.src_size && csize > (uint32_t)(T + 1) / 2 && q + WorstBlock<T>::READ <= lim &&
				       q[0] < (uint8_t)MARK_COPY;
			}
			if (__all_sync(FULL, fast || !valid)) {
				const uint32_t c = decode_block_rows<T>(q, fast, r, 16 * half, out);
				if (fast && (c == 0xFFFFFFFFu || c > csize))
					bad |= DEV_ERR_INVALID_INPUT;
			}
			else {
				const uint32_t e = decode_superblock_pair<T>(P.src, P.src_size, valid ? off : 0ull, dsize, out, valid, false, lz_scratch, lane, scratch_word);
				bad |= e;
			}
			if (bad && (lane & 15) == 0)
				atomicOr(&P.result[1], (unsigned long long)bad);
		}
	}
}
