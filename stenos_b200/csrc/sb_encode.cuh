// sb_encode.cuh -- one-warp-per-256-element-block encoder (device side).
//
// Replaces, for element sizes T in {2,4,8}, the reference's per-block pipeline
//   shuffle (shuffle.cpp:82-90) -> transpose_16x16 + find_pack_bits_params
//   (block_compress.h:864-906, 385-535) -> lz_compress (lz_compress.h:191-232, gate at
//   block_compress.h:1210-1223) -> encode16x16_generic (block_compress.h:739-806)
// and the tail encoder block_compress_partial / encode_lines (block_compress.h:947-1020, 686-737).
//
// Mapping: lane l of the warp owns elements 8l..8l+7 of the block, i.e. row r = l>>1 (16 bytes of
// every byte plane), half h = l&1.  No 16x16 transpose and no per-block shuffle exist here: byte
// planes are extracted in registers with PRMT, row statistics use 16x2 SIMD min/max plus one
// xor-shuffle, sizes/offsets come from ballots, REDUX and a 5-step scan.
#pragma once
#include "sb_common.cuh"

namespace sb
{
	// signed min / max of the 8 bytes (lo,hi) as s16x2 pairs (two partial results per register)
	__device__ __forceinline__ void minmax8_s16x2(uint32_t lo, uint32_t hi, uint32_t& mn, uint32_t& mx)
	{
		uint32_t a0 = prmt_sx(lo, 0, 0x9180); // sign-extended bytes 0,1
		uint32_t a1 = prmt_sx(lo, 0, 0xB3A2); // bytes 2,3
		uint32_t a2 = prmt_sx(hi, 0, 0x9180);
		uint32_t a3 = prmt_sx(hi, 0, 0xB3A2);
		mn = __vmins2(__vimin3_s16x2(a0, a1, a2), a3);
		mx = __vmaxs2(__vimax3_s16x2(a0, a1, a2), a3);
	}

	// packs four values (< 2^bits each, one per byte of x) into the low 4*bits bits
	__device__ __forceinline__ uint32_t pack4(uint32_t x, uint32_t bits)
	{
		uint32_t y = (x & 0x00FF00FFu) | ((x & 0xFF00FF00u) >> (8u - bits));
		return (y & 0xFFFFu) | ((y >> 16) << (2u * bits));
	}

	// stores the low `len` (0..8) bytes of (lo,hi) at the byte address o (shared memory)
	__device__ __forceinline__ void store_bytes8(uint8_t* o, uint32_t lo, uint32_t hi, uint32_t len)
	{
#pragma unroll
		for (uint32_t j = 0; j < 4; ++j)
			if (j < len)
				o[j] = (uint8_t)(lo >> (8 * j));
#pragma unroll
		for (uint32_t j = 0; j < 4; ++j)
			if (j + 4 < len)
				o[j + 4] = (uint8_t)(hi >> (8 * j));
	}

	// 8 mask bits -> 8 nibbles holding the inclusive prefix count of the set bits
	__device__ __forceinline__ uint32_t nibble_prefix_count(uint32_t m8)
	{
		uint32_t x = m8 & 0xFFu;
		x = (x | (x << 12)) & 0x000F000Fu;
		x = (x | (x << 6)) & 0x03030303u;
		x = (x | (x << 3)) & 0x11111111u;
		x += x << 4;
		x += x << 8;
		x += x << 16;
		return x;
	}

	// stores, in order and contiguously at o, the bytes of (lo,hi) whose bit in `keep` is set
	__device__ __forceinline__ void store_kept_bytes(uint8_t* o, uint32_t lo, uint32_t hi, uint32_t keep)
	{
		const uint32_t rank = nibble_prefix_count(keep); // inclusive
#pragma unroll
		for (uint32_t k = 0; k < 8; ++k)
			if ((keep >> k) & 1u)
				o[((rank >> (4 * k)) & 0xFu) - 1u] = (uint8_t)((k < 4 ? lo : hi) >> (8 * (k & 3)));
	}

	// Result of analysing + emitting one byte plane
	struct PlaneOut
	{
		uint32_t kind;
		uint32_t size;       // bytes written at o
		uint32_t check_size; // the size the reference uses in its room checks (8 + sum of row sizes)
	};

	// Analyses one byte plane of the block and writes its encoding at `o` (shared memory).
	// FULL_BLOCK = true : level-1 full block: row RLE / delta-RLE / mins-RLE and RAW promotion on, 16 rows
	// FULL_BLOCK = false: partial tail block: RLE off, no RAW promotion, only `lines` (< 16) rows emitted
	// Follows find_pack_bits_params (block_compress.h:385-535) for the decisions and
	// encode16x16_generic / encode_lines (:739-806 / :686-737) for the byte layout.
	template<bool FULL_BLOCK>
	__device__ __forceinline__ PlaneOut encode_plane(uint32_t lo, uint32_t hi, uint8_t* o, int lane, uint32_t lines)
	{
		constexpr bool RLE = FULL_BLOCK;
		const int half = lane & 1;
		const uint32_t row = (uint32_t)lane >> 1;
		PlaneOut out;

		// ---- all-same test (:396-418)
		const uint32_t first = __shfl_sync(FULL, lo, 0) & 0xFFu;
		const uint32_t pat = splat(first);
		if (__all_sync(FULL, (lo == pat) & (hi == pat))) {
			if (lane == 0)
				o[0] = (uint8_t)first;
			out.kind = KIND_SAME;
			out.size = out.check_size = 1;
			return out;
		}

		// ---- deltas: byte - previous byte; previous of the block's first byte is 0 (:399)
		uint32_t before = __shfl_up_sync(FULL, hi, 1);
		if (lane == 0)
			before = 0;
		const uint32_t dlo = __vsub4(lo, prev_bytes(before, lo));
		const uint32_t dhi = __vsub4(hi, prev_bytes(lo, hi));

		// ---- per-row signed min/max of values and of deltas (:403-413)
		uint32_t mnb, mxb, mnd, mxd;
		minmax8_s16x2(lo, hi, mnb, mxb);
		minmax8_s16x2(dlo, dhi, mnd, mxd);
		uint32_t MN = __vmins2(__byte_perm(mnb, mnd, 0x5410), __byte_perm(mnb, mnd, 0x7632)); // [min values | min deltas]
		uint32_t MX = __vmaxs2(__byte_perm(mxb, mxd, 0x5410), __byte_perm(mxb, mxd, 0x7632));
		MN = __vmins2(MN, __shfl_xor_sync(FULL, MN, 1));
		MX = __vmaxs2(MX, __shfl_xor_sync(FULL, MX, 1));
		const uint32_t R = __vsub2(MX, MN); // per 16-bit lane (a plain 32-bit subtract would borrow across lanes)
		const uint32_t rb = R & 0xFFFFu, rd = R >> 16;
		// bit widths (:334-352, :420-423): 7 -> 8 for both, and 6 -> 8 for the plain type
		const uint32_t b0 = rb >= 32u ? 8u : (32u - (uint32_t)__clz((int)rb));
		const uint32_t b1 = rd >= 64u ? 8u : (32u - (uint32_t)__clz((int)rd));
		const uint32_t bits = min(b0, b1);
		const bool plain = (b0 == bits);
		const uint32_t minv = (plain ? MN : (MN >> 16)) & 0xFFu;

		uint32_t sz = 2u * bits + (bits != 8u ? 1u : 0u);            // :433-435
		uint32_t h = plain ? (bits == 8u ? 15u : bits) : (8u + bits); // :499-502
		uint32_t payload = (bits == 8u) ? 16u : 2u * bits;            // bytes of the row after header/min

		// ---- RLE on values and on deltas (:439-474)
		uint32_t zr_lo = 0, zr_hi = 0, zd_lo = 0, zd_hi = 0;
		uint32_t cnt_pair = 0; // own non-repeat counts: values | deltas << 8
		if (RLE) {
			zr_lo = zero_bytes(dlo); // value repeats its predecessor <=> delta byte is 0
			zr_hi = zero_bytes(dhi);
			uint32_t dbefore = __shfl_up_sync(FULL, dhi, 1);
			if (!half)
				dbefore = 0; // delta-RLE restarts from 0 at every row (:449)
			zd_lo = zero_bytes(dlo ^ prev_bytes(dbefore, dlo));
			zd_hi = zero_bytes(dhi ^ prev_bytes(dlo, dhi));
			cnt_pair = (8u - __popc(zr_lo) - __popc(zr_hi)) | ((8u - __popc(zd_lo) - __popc(zd_hi)) << 8);
			const uint32_t c = cnt_pair + __shfl_xor_sync(FULL, cnt_pair, 1);
			const uint32_t rs = (c & 0xFFu) + 2u, ds = (c >> 8) + 2u;
			const bool use_rle = rs < sz;
			sz = min(sz, rs);
			const bool use_drle = ds < sz;
			sz = min(sz, ds);
			if (use_drle) {
				h = 6u;
				payload = ds;
			}
			else if (use_rle) {
				h = 7u;
				payload = rs;
			}
		}

		// ---- plane size, mins-RLE decision (:476-490), RAW promotion (:1200-1204)
		const bool live = row < lines; // rows that are emitted (partial blocks emit `lines` rows)
		uint32_t total = 8u + __reduce_add_sync(FULL, (half == 0 && live) ? sz : 0u);
		out.check_size = total;
		const bool needmin = !(h == 6u || h == 7u || h == 15u);
		uint32_t kind = KIND_NORMAL;
		uint32_t mball = 0;
		if (RLE) {
			const uint32_t nomin = __popc(__ballot_sync(FULL, !needmin) & 0x55555555u);
			uint32_t pm = __shfl_up_sync(FULL, minv, 2);
			if (lane < 2)
				pm = 0;
			mball = __ballot_sync(FULL, minv == pm) & 0x55555555u;
			const uint32_t mcnt = 16u - __popc(mball);
			if (mcnt + 2u < 16u - nomin) {
				kind = KIND_NORMAL_RLE;
				total -= (16u - nomin) - (mcnt + 2u);
			}
			if (total > 256u) { // raw plane: the 256 bytes in natural order
				store_bytes8(o + 8 * lane, lo, hi, 8);
				out.kind = KIND_RAW;
				out.size = out.check_size = 256;
				return out;
			}
		}
		out.kind = kind;

		// ---- row headers: two nibbles per byte, row 2i in the low nibble
		{
			const uint32_t hn = __shfl_down_sync(FULL, h, 2);
			const uint32_t hb = (row + 1u < lines) ? (h | (hn << 4)) : h; // odd last line: high nibble 0
			if ((lane & 3) == 0 && live)
				o[lane >> 2] = (uint8_t)hb;
		}
		uint8_t* q = o + ((lines + 1u) >> 1);
		// ---- mins: raw (rows that use one) or RLE coded
		if (RLE && kind == KIND_NORMAL_RLE) {
			uint32_t x = mball; // compress the even bits to a 16-bit mask
			x = (x | (x >> 1)) & 0x33333333u;
			x = (x | (x >> 2)) & 0x0F0F0F0Fu;
			x = (x | (x >> 4)) & 0x00FF00FFu;
			x = (x | (x >> 8)) & 0xFFFFu;
			if (lane == 0) {
				q[0] = (uint8_t)x;
				q[1] = (uint8_t)(x >> 8);
			}
			const uint32_t keep = ~mball & 0x55555555u;
			if ((keep >> lane) & 1u)
				q[2 + __popc(keep & lanemask_lt(lane))] = (uint8_t)minv;
			q += 2 + __popc(keep);
		}
		else {
			const uint32_t nm = __ballot_sync(FULL, needmin && live) & 0x55555555u;
			if (half == 0 && needmin && live)
				q[__popc(nm & lanemask_lt(lane))] = (uint8_t)minv;
			q += __popc(nm);
		}

		// ---- row offsets: exclusive scan of the row payload sizes (even lanes carry the row)
		const uint32_t mine = live ? payload : 0u;
		uint32_t incl = (half == 0) ? mine : 0u;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(FULL, incl, d);
			if (lane >= d)
				incl += t;
		}
		uint8_t* rowp = q + (incl - mine); // odd lanes contributed 0, so both halves agree
		out.size = (uint32_t)(q - o) + __shfl_sync(FULL, incl, 31);

		// ---- row payloads
		// RLE rows: [mask:2][non repeated values] (:258-293); each lane holds 8 of the 16 values.
		uint32_t m8 = 0, other = 0;
		const bool is_rle = RLE && (h == 7u || h == 6u);
		if (RLE && __any_sync(FULL, is_rle)) {
			const uint32_t zl = (h == 6u) ? zd_lo : zr_lo, zh = (h == 6u) ? zd_hi : zr_hi;
			m8 = flags_to_mask4(zl) | (flags_to_mask4(zh) << 4);
			const uint32_t own = (h == 6u) ? (cnt_pair >> 8) : (cnt_pair & 0xFFu);
			other = __shfl_xor_sync(FULL, m8 | (own << 8), 1); // the other half's mask bits and count
		}
		if (!live) {
		}
		else if (h == 15u) { // raw row (:635-637)
			store_bytes8(rowp + 8 * half, lo, hi, 8);
		}
		else if (is_rle) {
			uint8_t* p = rowp + 2;
			if (half)
				p += other >> 8;
			else {
				rowp[0] = (uint8_t)m8;
				rowp[1] = (uint8_t)other;
			}
			store_kept_bytes(p, h == 6u ? dlo : lo, h == 6u ? dhi : hi, ~m8 & 0xFFu);
		}
		else if (bits) {
			// bit packing of (value - min) or (delta - min), two groups of 8 (:540-602)
			const uint32_t m4 = splat(minv);
			const uint32_t vlo = __vsub4(plain ? lo : dlo, m4), vhi = __vsub4(plain ? hi : dhi, m4);
			const uint32_t a = pack4(vlo, bits), b = pack4(vhi, bits); // 4*bits bits each (bits <= 6 here)
			const uint32_t s = 4u * bits;                              // 4..24
			store_bytes8(rowp + bits * half, a | (b << s), b >> (32u - s), bits);
		}
		return out;
	}

	// ------------------------------------------------------------------------------------------
	// LZ-like matcher (lz_compress.h:191-232), warp-parallel reformulation.
	//
	// Sequential semantics: for element p, q = the most recent earlier position (in a group that
	// was not skipped) with the same 8-bit hash; a match iff the VALUES at q and p are equal.
	// Groups of 8 elements map to lanes.  `failed/max_failed` skipping is a 32-step state machine
	// on the "anchor is zero" ballot; because skipping a group changes later candidates, the
	// (matches -> ballot -> skip set) map is iterated to its fixed point, which is reached from
	// the front: after i iterations the first i groups are final.
	// ------------------------------------------------------------------------------------------
	template<int B>
	__device__ __forceinline__ uint32_t lz_hash(uint32_t vlo, uint32_t vhi)
	{
		if (B == 8) {
			unsigned long long v = ((unsigned long long)vhi << 32) | vlo;
			return (uint32_t)((v * 14313749767032793493ULL) >> 56); // lz_compress.h:52-56
		}
		return (vlo * 2654435761u) & 255u; // lz_compress.h:47-51
	}

	__device__ __forceinline__ uint32_t lz_skip_set(uint32_t zero_anchor)
	{
		// lz_compress.h:206-219: the group after `max_failed` zero anchors is copied raw, not hashed
		if (__popc(zero_anchor) < 3)
			return 0;
		uint32_t failed = 0, maxf = 3, skip = 0;
		for (int g = 0; g < 32; ++g) {
			if (failed == maxf) {
				skip |= 1u << g;
				failed = 0;
				maxf = maxf > 1 ? maxf - 1 : 1;
			}
			else
				failed += (zero_anchor >> g) & 1u;
		}
		return skip;
	}

	constexpr uint32_t LZ_SCRATCH_BYTES = 256 * 4 + 256;

	// T in {4,8}: element width B = T, 256 elements, lane = group of 8.
	// scratch: per-warp shared memory, 256 x u32 (hash -> lanes containing it) + 256 x u8 hashes.
	// Returns the LZ stream length (written at `o`), or 0 when the matcher gives up.
	template<int T>
	__device__ __noinline__ uint32_t lz_encode_block(const uint8_t* __restrict__ gsrc, const uint32_t (&w)[2 * T], uint8_t* o, uint32_t max_size, uint32_t* scratch, int lane)
	{
		constexpr int B = T; // 4 or 8
		uint32_t* lanes_with = scratch;                              // [256]
		uint8_t* hashes = reinterpret_cast<uint8_t*>(scratch + 256); // [256], position indexed
		for (int i = lane; i < 256; i += 32)
			lanes_with[i] = 0;
		__syncwarp();
		uint32_t hpack_lo = 0, hpack_hi = 0;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			uint32_t hv = (B == 8) ? lz_hash<B>(w[(2 * k) % (2 * T)], w[(2 * k + 1) % (2 * T)]) : lz_hash<B>(w[k % (2 * T)], 0);
			if (k < 4)
				hpack_lo |= hv << (8 * k);
			else
				hpack_hi |= hv << (8 * (k - 4));
			atomicOr(&lanes_with[hv], 1u << lane);
		}
		*reinterpret_cast<uint2*>(hashes + 8 * lane) = make_uint2(hpack_lo, hpack_hi);
		__syncwarp();

		uint32_t skip = 0, match = 0; // match: 8 bits, element k of this lane found a match
		uint32_t offs[8];             // back offsets (elements)
		for (int iter = 0; iter < 33; ++iter) {
			match = 0;
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				const uint32_t hv = ((k < 4 ? hpack_lo : hpack_hi) >> (8 * (k & 3))) & 0xFFu;
				// candidate inside this group: latest k' < k with the same hash
				const uint32_t same_lo = flags_to_mask4(zero_bytes(hpack_lo ^ splat(hv)));
				const uint32_t same_hi = flags_to_mask4(zero_bytes(hpack_hi ^ splat(hv)));
				const uint32_t same = (same_lo | (same_hi << 4)) & ((1u << k) - 1u);
				int q = -1;
				if (same)
					q = 8 * lane + (31 - __clz((int)same));
				else {
					const uint32_t m = lanes_with[hv] & lanemask_lt(lane) & ~skip;
					if (m) {
						const int l2 = 31 - __clz((int)m);
						const uint2 hh = *reinterpret_cast<const uint2*>(hashes + 8 * l2);
						const uint32_t s2 = flags_to_mask4(zero_bytes(hh.x ^ splat(hv))) | (flags_to_mask4(zero_bytes(hh.y ^ splat(hv))) << 4);
						q = 8 * l2 + (31 - __clz((int)s2));
					}
				}
				offs[k] = 0;
				if (q >= 0) {
					bool eq;
					if (B == 8) {
						const uint2 c = *reinterpret_cast<const uint2*>(gsrc + (size_t)q * 8);
						eq = (c.x == w[(2 * k) % (2 * T)]) & (c.y == w[(2 * k + 1) % (2 * T)]);
					}
					else
						eq = *reinterpret_cast<const uint32_t*>(gsrc + (size_t)q * 4) == w[k % (2 * T)];
					if (eq) {
						match |= 1u << k;
						offs[k] = (uint32_t)(8 * lane + k - q);
					}
				}
			}
			const bool skipped_now = (skip >> lane) & 1u;
			const uint32_t zero_anchor = __ballot_sync(FULL, !skipped_now && match == 0);
			const uint32_t nskip = lz_skip_set(zero_anchor);
			if (nskip == skip)
				break;
			skip = nskip;
		}
		const bool skipped = (skip >> lane) & 1u;
		if (skipped)
			match = 0;

		// sizes: [anchor] + per element B raw bytes or a 1-2 byte back offset (:140-151)
		uint32_t gsz = 1;
#pragma unroll
		for (int k = 0; k < 8; ++k)
			gsz += ((match >> k) & 1u) ? (offs[k] < 128u ? 1u : 2u) : (uint32_t)B;
		uint32_t incl = gsz;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(FULL, incl, d);
			if (lane >= d)
				incl += t;
		}
		const uint32_t total = __shfl_sync(FULL, incl, 31);
		// early exit once, after the first group with i > count/4 (count = 256 -> group 9) (:224-229)
		const uint32_t at9 = __shfl_sync(FULL, incl, 9);
		if (total > max_size || 5u * at9 > 2u * max_size)
			return 0;

		uint8_t* p = o + (incl - gsz);
		*p++ = (uint8_t)match;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			if ((match >> k) & 1u) {
				if (offs[k] < 128u)
					*p++ = (uint8_t)offs[k];
				else {
					*p++ = (uint8_t)((offs[k] & 127u) | 128u);
					*p++ = (uint8_t)(offs[k] >> 7);
				}
			}
			else {
#pragma unroll
				for (int j = 0; j < B; ++j)
					*p++ = (uint8_t)(((B == 8) ? w[(2 * k + (j >> 2)) % (2 * T)] : w[k % (2 * T)]) >> (8 * (j & 3)));
			}
		}
		return total;
	}

	// ------------------------------------------------------------------------------------------
	// One full 256-element block -> slot (shared memory).  Returns its encoded size.
	//
	// room : bytes the reference would have had between the block's first output byte and dst_end
	//        (the room arithmetic changes decisions).  EXACT=false assumes
	//        the checks are inert (the caller proved room is ample).
	// err  : set when the reference would have returned an error for the superblock (-> COPY).
	// ------------------------------------------------------------------------------------------
	template<int T, bool EXACT>
	__device__ __forceinline__ uint32_t encode_block(const uint8_t* __restrict__ gsrc, uint8_t* slot, uint32_t* lz_scratch, int lane, uint32_t room, bool& err)
	{
		constexpr uint32_t HS = (T + 1) / 2;
		uint32_t w[2 * T];
		load_lane_words<T>(gsrc, lane, w);
		uint32_t lo[T], hi[T];
		words_to_planes<T>(w, lo, hi);

		uint32_t pos = HS;
		uint32_t kinds = 0;
		bool e = false;
#pragma unroll
		for (int p = 0; p < T; ++p) {
			const PlaneOut r = encode_plane<true>(lo[p], hi[p], slot + pos, lane, 16);
			kinds |= r.kind << (4 * p);
			if (EXACT) {
				// block_compress.h:1241 (non raw planes) and :1248 (kind nibble byte)
				if (r.kind != KIND_RAW && pos + r.size + 16u > room)
					e = true;
				if ((p & 1) == 0 && (uint32_t)(p >> 1) >= room)
					e = true;
			}
			pos += r.size;
		}
		const uint32_t full = pos - HS;

		// ---- LZ attempt (block_compress.h:1210-1223).  Matching reads the block from global
		// memory and all sizes are decided before a byte is written, so on success the stream
		// simply overwrites the plane encoding in the slot.
		if ((T % 4) == 0 && full * 3u > (uint32_t)T * 256u) {
			if (!EXACT || room > HS + full + (uint32_t)T * 8u + 2u) {
				__syncwarp();
				const uint32_t r = lz_encode_block<T>(gsrc, w, slot + 1, full, lz_scratch, lane);
				if (r) {
					if (lane == 0)
						slot[0] = (uint8_t)MARK_LZ;
					return r + 1u;
				}
			}
		}
		if (EXACT && (e || HS + full > room)) // :1225
			err = true;
		if (lane < (int)HS)
			slot[lane] = (uint8_t)(kinds >> (8 * lane));
		return pos;
	}

	// ------------------------------------------------------------------------------------------
	// Tail of fewer than 256 elements (block_compress.h:1277-1298 + block_compress_partial :947-1020):
	// [254] [kinds] planes (only the full 16-element lines, RLE off) [remaining bytes raw].
	// `bytes` < T*256.  Returns the encoded size including the marker.
	// ------------------------------------------------------------------------------------------
	template<int T, bool EXACT>
	__device__ __noinline__ uint32_t encode_partial_block(const uint8_t* __restrict__ gsrc, uint32_t bytes, uint8_t* slot, int lane, uint32_t room, bool& err)
	{
		constexpr uint32_t HS = (T + 1) / 2;
		const uint32_t line = 16u * T, lines = bytes / line;
		bool e = EXACT && room < 2u; // :1284
		if (lane == 0)
			slot[0] = (uint8_t)MARK_PARTIAL;
		uint32_t pos = 1;
		if (lines) {
			// the block padded to 256 elements with the LAST BYTE of the tail (:967-968)
			const uint32_t padb = gsrc[bytes - 1];
			uint32_t w[2 * T];
#pragma unroll
			for (int i = 0; i < 2 * T; ++i) {
				uint32_t v = 0;
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const uint32_t idx = (uint32_t)lane * 8u * T + 4u * i + j;
					v |= (idx < bytes ? (uint32_t)gsrc[idx] : padb) << (8 * j);
				}
				w[i] = v;
			}
			uint32_t lo[T], hi[T];
			words_to_planes<T>(w, lo, hi);
			uint32_t kinds = 0;
			pos += HS;
#pragma unroll
			for (int p = 0; p < T; ++p) {
				const PlaneOut r = encode_plane<false>(lo[p], hi[p], slot + pos, lane, lines);
				kinds |= r.kind << (4 * p);
				if (EXACT) {
					// :984 and :994 (positions relative to the byte after the marker, room likewise)
					if (r.kind == KIND_SAME ? (pos >= room) : (pos + r.check_size + 8u > room))
						e = true;
				}
				pos += r.size;
				__syncwarp();
			}
			if (lane < (int)HS)
				slot[1 + lane] = (uint8_t)(kinds >> (8 * lane));
		}
		const uint32_t rem = bytes - lines * line;
		if (EXACT && rem && pos + rem > room) // :1013
			e = true;
		for (uint32_t i = lane; i < rem; i += 32)
			slot[pos + i] = gsrc[lines * line + i];
		if (e)
			err = true;
		return pos + rem;
	}
}
