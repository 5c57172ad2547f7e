// sb_encode_rows.cuh -- the level-1 block encoder of the fast path: one warp per PAIR of
// 256-element blocks, one lane per 16-element ROW.
//
// Same decisions and bytes as sb_encode.cuh (which stays for the dst-room-exact path, the partial
// tail block and the LZ matcher), i.e. the reference's
//   shuffle (shuffle.cpp:82-90) -> transpose_16x16 + find_pack_bits_params
//   (block_compress.h:864-906, 385-535) -> encode16x16_generic (block_compress.h:739-806)
// but laid out so that everything the reference computes per 16-byte row is lane-local:
//
//   lane l: block (l >> 4) of the pair, row r = l & 15 = elements 16r..16r+15 = 16*T contiguous
//   bytes = T LDG.E.128.  Byte plane p of the row = 4 registers built with PRMT.  Row statistics
//   (signed min/max of the values and of the deltas, RLE / delta-RLE counts, bit width, size, header
//   nibble) need no shuffle at all; across the 16 lanes of a block only the previous row's last byte,
//   the previous row's min, the all-same vote, the ballots of the min bookkeeping and one 4-step scan
//   of the row sizes are exchanged.  Two blocks per warp double the work per issued instruction
//   against the lane-per-half-row mapping and give every lane four independent words of ILP.
#pragma once
#include "sb_common.cuh"
#include "sb_encode.cuh"

namespace sb
{
	// sum of absolute byte differences of a and b, plus c (one VABSDIFF4.U8.ACC)
	__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t c)
	{
#ifdef STENOS_EMU
		uint32_t s = c;
		for (int i = 0; i < 4; ++i) {
			const int x = (int)((a >> (8 * i)) & 0xFFu), y = (int)((b >> (8 * i)) & 0xFFu);
			s += (uint32_t)(x > y ? x - y : y - x);
		}
		return s;
#else
		uint32_t r;
		asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
		return r;
#endif
	}

	// 0x80 in every NON-zero byte of x
	__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }

	// signed min and max of the 16 bytes a[0..3] as s16x2 pairs whose HIGH bytes carry the result:
	// a 16-bit lane compares like its (signed) high byte first, so the odd bytes need no unpacking and
	// the even bytes only a shift into the high position.
	__device__ __forceinline__ void minmax16_hi(const uint32_t (&a)[4], uint32_t& mn, uint32_t& mx)
	{
		const uint32_t y0 = a[0] << 8, y1 = a[1] << 8, y2 = a[2] << 8, y3 = a[3] << 8;
		mn = __vimin3_s16x2(a[0], a[1], a[2]);
		mx = __vimax3_s16x2(a[0], a[1], a[2]);
		mn = __vimin3_s16x2(mn, a[3], y0);
		mx = __vimax3_s16x2(mx, a[3], y0);
		mn = __vimin3_s16x2(mn, y1, y2);
		mx = __vimax3_s16x2(mx, y1, y2);
		mn = __vmins2(mn, y3);
		mx = __vmaxs2(mx, y3);
	}

	// byte deltas of a row: d[j] = a[j] - (previous bytes); pw's byte 3 is the byte before the row
	__device__ __forceinline__ void row_deltas(const uint32_t (&a)[4], uint32_t pw, uint32_t (&d)[4])
	{
		d[0] = __vsub4(a[0], __byte_perm(pw, a[0], 0x6543));
		d[1] = __vsub4(a[1], __byte_perm(a[0], a[1], 0x6543));
		d[2] = __vsub4(a[2], __byte_perm(a[1], a[2], 0x6543));
		d[3] = __vsub4(a[3], __byte_perm(a[2], a[3], 0x6543));
	}
	// x[j] = d[j] ^ (the delta before it, 0 before the first): zero bytes <=> the delta repeats (:449)
	__device__ __forceinline__ void delta_repeats(const uint32_t (&d)[4], uint32_t (&x)[4])
	{
		x[0] = d[0] ^ (d[0] << 8);
		x[1] = d[1] ^ __byte_perm(d[0], d[1], 0x6543);
		x[2] = d[2] ^ __byte_perm(d[1], d[2], 0x6543);
		x[3] = d[3] ^ __byte_perm(d[2], d[3], 0x6543);
	}

	// Per-row analysis (find_pack_bits_params, block_compress.h:399-474): header nibble, min byte and
	// payload bytes of the row.  The row size the reference sums is pay + (the row stores a min ? 1 : 0).
	__device__ __forceinline__ void analyse_row(const uint32_t (&a)[4], uint32_t pw, uint32_t& h, uint32_t& minv, uint32_t& pay)
	{
		uint32_t d[4];
		row_deltas(a, pw, d);
		uint32_t mnv, mxv, mnd, mxd;
		minmax16_hi(a, mnv, mxv);
		minmax16_hi(d, mnd, mxd);
		// fold the two 16-bit lanes: [values | deltas] as sign-extended s16x2
		const uint32_t MN = __vmins2(prmt_sx(mnv, mnd, 0xD591), prmt_sx(mnv, mnd, 0xF7B3));
		const uint32_t MX = __vmaxs2(prmt_sx(mxv, mxd, 0xD591), prmt_sx(mxv, mxd, 0xF7B3));
		const uint32_t R = __vsub2(MX, MN);
		const uint32_t rb = R & 0xFFFFu, rd = R >> 16;
		// bit widths (:334-352, :420-423): 7 -> 8 for both, and 6 -> 8 for the plain type
		const uint32_t b0 = rb >= 32u ? 8u : (32u - (uint32_t)__clz((int)rb));
		const uint32_t b1 = rd >= 64u ? 8u : (32u - (uint32_t)__clz((int)rd));
		const uint32_t bits = min(b0, b1);
		const bool plain = (b0 == bits);
		minv = (plain ? MN : (MN >> 16)) & 0xFFu;
		uint32_t sz = 2u * bits + (bits != 8u ? 1u : 0u);   // :433-435
		h = plain ? (bits == 8u ? 15u : bits) : (8u + bits); // :499-502
		pay = (bits == 8u) ? 16u : 2u * bits;
		// RLE on the values and on the deltas (:439-474): non repeated bytes + 2 mask bytes
		uint32_t x[4];
		delta_repeats(d, x);
		uint32_t nr = 0, nd = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			nr = sad4_acc(nonzero_bytes(d[j]), 0u, nr);
			nd = sad4_acc(nonzero_bytes(x[j]), 0u, nd);
		}
		const uint32_t rs = (nr >> 7) + 2u, ds = (nd >> 7) + 2u;
		const bool use_rle = rs < sz;
		sz = min(sz, rs);
		const bool use_drle = ds < sz;
		if (use_drle) {
			h = 6u;
			pay = ds;
		}
		else if (use_rle) {
			h = 7u;
			pay = rs;
		}
	}

	// 16 mask bits (bit k = byte k of x[] is zero)
	__device__ __forceinline__ uint32_t zero_mask16(const uint32_t (&x)[4])
	{
		uint32_t m = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j)
			m |= flags_to_mask4(nonzero_bytes(x[j]) ^ 0x80808080u) << (4 * j);
		return m;
	}

	// stores the first n (0..16) bytes of v[0..3] at the byte address o
	__device__ __forceinline__ void store_bytes16(uint8_t* o, const uint32_t (&v)[4], uint32_t n)
	{
#pragma unroll
		for (uint32_t j = 0; j < 16; ++j)
			if (j < n)
				o[j] = (uint8_t)(v[j >> 2] >> (8 * (j & 3)));
	}

	// stores, in order and contiguously at o, the bytes of v[0..3] whose bit in `keep` (16 bits) is set
	__device__ __forceinline__ void store_kept_bytes16(uint8_t* o, const uint32_t (&v)[4], uint32_t keep)
	{
		store_kept_bytes(o, v[0], v[1], keep & 0xFFu);
		store_kept_bytes(o + __popc(keep & 0xFFu), v[2], v[3], (keep >> 8) & 0xFFu);
	}

	// packs four values (< 2^bits each, one per byte of x) into the low 4*bits bits; mul = 1 << bits
	__device__ __forceinline__ uint32_t pack4_mul(uint32_t x, uint32_t mul)
	{
		const uint32_t c = (x & 0x00FF00FFu) + ((x >> 8) & 0x00FF00FFu) * mul; // two lanes of v0 | v1 << bits
		return (c & 0xFFFFu) + (c >> 16) * (mul * mul);
	}

	template<int T>
	__device__ __forceinline__ void load_row_planes(const uint8_t* __restrict__ block, int r, uint32_t (&pw)[T][4])
	{
		const uint4* src = reinterpret_cast<const uint4*>(block + (size_t)r * 16 * T);
		uint32_t e[4 * T];
#if !defined(ENCODE_LOAD_128)
		if ((reinterpret_cast<uintptr_t>(src) & 31u) == 0) {
			// the row's 16 * T bytes are contiguous: whole 32-byte sectors per instruction (LDG.E.256)
#pragma unroll
			for (int i = 0; i < T; i += 2) {
				uint4 a, b;
				ld_global_256(src + i, a, b);
				e[4 * i + 0] = a.x;
				e[4 * i + 1] = a.y;
				e[4 * i + 2] = a.z;
				e[4 * i + 3] = a.w;
				e[4 * i + 4] = b.x;
				e[4 * i + 5] = b.y;
				e[4 * i + 6] = b.z;
				e[4 * i + 7] = b.w;
			}
		}
		else
#endif
		{
#pragma unroll
			for (int i = 0; i < T; ++i) {
				const uint4 v = src[i];
				e[4 * i + 0] = v.x;
				e[4 * i + 1] = v.y;
				e[4 * i + 2] = v.z;
				e[4 * i + 3] = v.w;
			}
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if constexpr (T == 2) {
				pw[0][j] = __byte_perm(e[2 * j], e[2 * j + 1], 0x6420);
				pw[1][j] = __byte_perm(e[2 * j], e[2 * j + 1], 0x7531);
			}
			else if constexpr (T == 4) {
				transpose4(e[4 * j], e[4 * j + 1], e[4 * j + 2], e[4 * j + 3], pw[0][j], pw[1][j], pw[2][j], pw[3][j]);
			}
			else {
				transpose4(e[8 * j], e[8 * j + 2], e[8 * j + 4], e[8 * j + 6], pw[0][j], pw[1][j], pw[2][j], pw[3][j]);
				transpose4(e[8 * j + 1], e[8 * j + 3], e[8 * j + 5], e[8 * j + 7], pw[4][j], pw[5][j], pw[6][j], pw[7][j]);
			}
		}
	}

	// ------------------------------------------------------------------------------------------
	// Two full 256-element blocks -> shared memory.  blk0: block of the lower half-warp; the upper
	// half-warp works on blk0 + T*256 when `second` is set, else it shadows the lower half without
	// writing.  Everything is sized first; then place(size) -- called once, convergently, with the
	// lane's block size (uniform over each half-warp) -- returns where the lane's block goes (or
	// nullptr when its bytes are not wanted).  LZ streams, whose size is only known once they exist,
	// are built at tmp0 + half * tmp_stride and moved.  Ample dst room is assumed (the caller proved
	// the reference's room checks inert).
	// Returns the encoded size of the lane's block (uniform over each half-warp).
	// ------------------------------------------------------------------------------------------
	template<int T, class Place>
	__device__ __forceinline__ uint32_t encode_block_pair(const uint8_t* __restrict__ blk0, bool second, uint8_t* tmp0, uint32_t tmp_stride, uint32_t* lz_scratch, int lane,
							      Place&& place, const uint8_t* __restrict__ blk1 = nullptr)
	{
		constexpr uint32_t HS = (T + 1) / 2;
		const int hb = lane >> 4, r = lane & 15;
		const bool upper = hb && second;
		const bool writer = !hb || second;
		if (blk1 == nullptr)
			blk1 = blk0 + T * 256; // the second block follows the first one (cvector buckets picked by id: anywhere)
		const uint8_t* blk = upper ? blk1 : blk0;
		const uint32_t below = (1u << r) - 1u; // rows before mine

		uint32_t pw[T][4];
		load_row_planes<T>(blk, r, pw);

		// ---- analysis: per plane [h | minv << 8 | pay << 16], all-same planes flagged in `same`
		uint32_t info[T], prevw[T];
		uint32_t same = 0, both_same = 0;
#pragma unroll
		for (int p = 0; p < T; ++p) {
			const uint32_t s = __byte_perm(pw[p][0], pw[p][0], 0x0000);
			const uint32_t t = ((pw[p][0] ^ s) | (pw[p][1] ^ s)) | ((pw[p][2] ^ s) | (pw[p][3] ^ s));
			const uint32_t first = __shfl_sync(FULL, s, 0, 16);
			const uint32_t sm = __ballot_sync(FULL, t == 0u && s == first); // :396-418
			info[p] = 0;
			prevw[p] = 0;
			if (((sm >> (16 * hb)) & 0xFFFFu) == 0xFFFFu)
				same |= 1u << p;
			if (sm == FULL) {
				both_same |= 1u << p;
				continue;
			}
			uint32_t pv = __shfl_up_sync(FULL, pw[p][3], 1, 16);
			if (r == 0)
				pv = 0; // the byte before the block is 0 (:399)
			prevw[p] = pv;
			uint32_t h, minv, pay;
			analyse_row(pw[p], pv, h, minv, pay);
			info[p] = h | (minv << 8) | (pay << 16);
		}

		// ---- plane sizes and kinds (:476-490, :1200-1204); row offsets
		uint32_t kinds = 0;      // 4 bits per plane
		uint32_t psize[T];       // encoded size of the plane
		uint32_t rowoff[T];      // offset of my row's payload from the start of the plane
		uint32_t minoff[T];      // offset of my min byte from the start of the plane (0xFFFF: none)
		uint32_t mmask[T];       // mins-RLE mask of the plane (kind 3)
#pragma unroll
		for (int p = 0; p < T; ++p) {
			psize[p] = 1;
			rowoff[p] = 0;
			minoff[p] = 0xFFFFu;
			mmask[p] = 0;
			if ((both_same >> p) & 1u)
				continue; // kind 0
			const uint32_t h = info[p] & 0xFu, minv = (info[p] >> 8) & 0xFFu, pay = info[p] >> 16;
			const bool needmin = !(h == 6u || h == 7u || h == 15u);
			uint32_t pm = __shfl_up_sync(FULL, minv, 1, 16);
			if (r == 0)
				pm = 0;
			const uint32_t nmb = (__ballot_sync(FULL, needmin) >> (16 * hb)) & 0xFFFFu;
			const uint32_t mmb = (__ballot_sync(FULL, minv == pm) >> (16 * hb)) & 0xFFFFu;
			uint32_t incl = pay;
#pragma unroll
			for (int dlt = 1; dlt < 16; dlt <<= 1) {
				const uint32_t t = __shfl_up_sync(FULL, incl, dlt, 16);
				if (r >= dlt)
					incl += t;
			}
			const uint32_t paysum = __shfl_sync(FULL, incl, 15, 16);
			const uint32_t nmins = __popc(nmb);
			uint32_t total = 8u + paysum + nmins;
			uint32_t kind = KIND_NORMAL, minbytes = nmins;
			const uint32_t mcnt = 16u - __popc(mmb);
			if (mcnt + 2u < nmins) { // mins-RLE (:480-490): 16 - nomin == nmins
				kind = KIND_NORMAL_RLE;
				total -= nmins - (mcnt + 2u);
				minbytes = mcnt + 2u;
			}
			if (total > 256u) { // raw plane (:1200-1204)
				kind = KIND_RAW;
				total = 256u;
			}
			if ((same >> p) & 1u) { // only this half's plane is all-same
				kind = KIND_SAME;
				total = 1u;
			}
			kinds |= kind << (4 * p);
			psize[p] = total;
			rowoff[p] = 8u + minbytes + (incl - pay);
			mmask[p] = mmb;
			if (kind == KIND_NORMAL_RLE)
				minoff[p] = ((mmb >> r) & 1u) ? 0xFFFFu : 8u + 2u + (uint32_t)__popc(~mmb & below);
			else if (needmin)
				minoff[p] = 8u + (uint32_t)__popc(nmb & below);
		}
		uint32_t full = 0;
#pragma unroll
		for (int p = 0; p < T; ++p)
			full += psize[p];
		uint32_t size = HS + full;

		// ---- LZ attempt (block_compress.h:1210-1223): blocks whose plane coding ratio is < 3.  Rare on
		// compressible data; runs with the lane-per-half-row matcher of sb_encode.cuh, one block at a time.
		bool lz_done = false;
		if ((T % 4) == 0) {
			const uint32_t want = __ballot_sync(FULL, full * 3u > (uint32_t)T * 256u);
			if (want) {
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					if (!((want >> (16 * hh)) & 1u) || (hh && !second))
						continue;
					const uint8_t* gs = hh ? blk1 : blk0;
					uint8_t* sl = tmp0 + (size_t)hh * tmp_stride;
					const uint32_t fmax = __shfl_sync(FULL, full, 16 * hh);
					uint32_t w[2 * T];
					load_lane_words<T>(gs, lane, w);
					__syncwarp();
					const uint32_t lr = lz_encode_block<T>(gs, w, sl + 1, fmax, lz_scratch, lane);
					if (lr) {
						if (lane == 0)
							sl[0] = (uint8_t)MARK_LZ;
						if (hb == hh) {
							lz_done = true;
							size = lr + 1u;
						}
					}
					__syncwarp();
				}
			}
		}

		// ---- placement: the sizes of both blocks are final; the caller says where the lane's block goes
		// (nullptr: the bytes are not wanted).  LZ streams were built in the temporary and move now.
		uint8_t* slot = place(size);
		if ((T % 4) == 0) {
			const uint32_t lzm = __ballot_sync(FULL, lz_done);
			if (lzm) {
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					if (!((lzm >> (16 * hh)) & 1u))
						continue;
					uint8_t* to = reinterpret_cast<uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)slot, 16 * hh));
					const uint8_t* from = tmp0 + (size_t)hh * tmp_stride;
					const uint32_t n = __shfl_sync(FULL, size, 16 * hh);
					if (to && to != from)
						for (uint32_t i = lane; i < n; i += 32)
							to[i] = from[i];
				}
				__syncwarp();
			}
		}

		// ---- emission (encode16x16_generic, :739-806)
		const bool emit = writer && !lz_done && slot != nullptr;
		if (emit && r == 0) {
#pragma unroll
			for (uint32_t i = 0; i < HS; ++i)
				slot[i] = (uint8_t)(kinds >> (8 * i));
		}
		uint32_t pos = HS;
#pragma unroll
		for (int p = 0; p < T; ++p) {
			uint8_t* o = slot + pos;
			pos += psize[p];
			if ((both_same >> p) & 1u) {
				if (emit && r == 0)
					o[0] = (uint8_t)pw[p][0];
				continue;
			}
			const uint32_t kind = (kinds >> (4 * p)) & 0xFu;
			const uint32_t h = info[p] & 0xFu, minv = (info[p] >> 8) & 0xFFu;
			const uint32_t hn = __shfl_down_sync(FULL, h, 1, 16);
			if (!emit)
				continue;
			if (kind == KIND_SAME) {
				if (r == 0)
					o[0] = (uint8_t)pw[p][0];
				continue;
			}
			if (kind == KIND_RAW) {
				store_bytes16(o + 16 * r, pw[p], 16);
				continue;
			}
			// row headers: two nibbles per byte, row 2i in the low nibble
			if ((r & 1) == 0)
				o[r >> 1] = (uint8_t)(h | (hn << 4));
			// mins: raw (rows that use one) or RLE coded [mask:2][values that differ from the previous min]
			if (kind == KIND_NORMAL_RLE && r == 0) {
				o[8] = (uint8_t)mmask[p];
				o[9] = (uint8_t)(mmask[p] >> 8);
			}
			if (minoff[p] != 0xFFFFu)
				o[minoff[p]] = (uint8_t)minv;
			uint8_t* rowp = o + rowoff[p];
			if (h == 15u) { // raw row (:635-637)
				store_bytes16(rowp, pw[p], 16);
			}
			else if (h == 7u) { // [mask:2][non repeated values] (:258-265)
				uint32_t d[4];
				row_deltas(pw[p], prevw[p], d);
				const uint32_t m = zero_mask16(d);
				rowp[0] = (uint8_t)m;
				rowp[1] = (uint8_t)(m >> 8);
				store_kept_bytes16(rowp + 2, pw[p], ~m & 0xFFFFu);
			}
			else if (h == 6u) { // the same on the deltas (:266-293)
				uint32_t d[4], x[4];
				row_deltas(pw[p], prevw[p], d);
				delta_repeats(d, x);
				const uint32_t m = zero_mask16(x);
				rowp[0] = (uint8_t)m;
				rowp[1] = (uint8_t)(m >> 8);
				store_kept_bytes16(rowp + 2, d, ~m & 0xFFFFu);
			}
			else {
				const uint32_t bits = h & 7u;
				if (bits) {
					// bit packing of (value - min) or (delta - min), two groups of 8 (:540-602).  With both
					// sides biased to unsigned order no byte of the word subtraction borrows.
					uint32_t v[4];
					if (h & 8u)
						row_deltas(pw[p], prevw[p], v);
					else {
#pragma unroll
						for (int j = 0; j < 4; ++j)
							v[j] = pw[p][j];
					}
					const uint32_t mb = splat(minv) ^ 0x80808080u;
					const uint32_t mul = 1u << bits;
					uint32_t pk[4];
#pragma unroll
					for (int j = 0; j < 4; ++j)
						pk[j] = pack4_mul((v[j] ^ 0x80808080u) - mb, mul); // 4*bits bits each
					const uint32_t s = 4u * bits;                           // 4..24
					const uint32_t g0l = pk[0] | (pk[1] << s), g0h = pk[1] >> (32u - s);
					const uint32_t g1l = pk[2] | (pk[3] << s), g1h = pk[3] >> (32u - s);
					store_bytes8(rowp, g0l, g0h, bits);
					store_bytes8(rowp + bits, g1l, g1h, bits);
				}
			}
		}
		return size;
	}
}
