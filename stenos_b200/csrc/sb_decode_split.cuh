// sb_decode_split.cuh -- decode_split_kernel: the level-1 frame decoder as PARSER threads + DECODER warps.
//
// The block stream of a superblock is not seekable: block k's offset is known only
// once blocks 0..k-1 have been walked, and inside a block every plane's length depends on its row
// headers and on the masks of its RLE rows (block_compress.h:2088-2175, :2046-2084, :1939-1968).
// decode_pairs_kernel pays that serial chain once per block inside the decoding warp: the frame
// decoder is then bound by the LATENCY of one stream (measured: a 128-block stream takes the same
// ~310 us whether 4 or 24 warps share the SM), and by the quantisation of 55 streams per SM.
//
// Here the chain is walked by scalar threads that do nothing else (one thread per superblock: 32
// independent chains per warp, a few loads and ~50 integer instructions per plane), and everything
// that costs instructions -- unpacking, prefix sums, carries, the byte transpose, the stores -- is done
// by half-warps that take single BLOCKS as work items, in block-major order (block k of every
// superblock, then block k + 1, ...), so the decoders follow closely behind the parsers' front and
// never wait for a whole stream.  Roles are handed out by ticket, parsers first: a decoder only waits
// for entries whose parser already runs, whatever the order in which the hardware starts CTAs.
//
// Block entry (one u64 per block, zero = not parsed yet):  flags:2 | encoded length:14 | offset in src:48.
// With the length in hand a decoder first asks for exactly the block's lines (one instruction, all misses in
// flight together) and then runs its dependent loads against L1.  The decoders re-check everything they
// read, within [offset, offset + length); a parser that meets a malformed stream sets the error bits and
// marks the rest of the superblock SKIP, so nobody waits for ever.
#pragma once
#include "sb_decode_rows.cuh"

namespace sb
{
	constexpr unsigned long long ENT_RAWCOPY = 1ull << 62; // the block is raw bytes of a COPY superblock (stenos.cpp:743-751)
	constexpr unsigned long long ENT_SKIP = 1ull << 63;    // nothing to decode (error already reported, or the host decodes it)
	constexpr unsigned long long ENT_OFFSET = (1ull << 48) - 1ull;
	constexpr uint32_t ENT_LEN_MAX = (1u << 14) - 1u;
	__device__ __forceinline__ unsigned long long make_entry(unsigned long long flags, uint32_t len, uint64_t off)
	{
		return flags | ((unsigned long long)min(len, ENT_LEN_MAX) << 48) | (unsigned long long)off;
	}
	constexpr uint32_t PARSE_BAD = 0xFFFFFFFFu;

	struct SplitParams
	{
		DecodeParams d;
		unsigned long long* entries; // [n_sb x kmax], zero initialised
		uint32_t kmax;        // blocks (full + partial) of a full superblock of this launch
		uint32_t n_parse_cta; // CTAs whose ticket makes them parsers
		uint32_t parse_only;  // experiments: 1 decoders leave at once, 2 parsers leave at once (entries already there)
	};

	// sum of the eight nibbles of x
	__device__ __forceinline__ uint32_t nibble_sum(uint32_t x)
	{
		x = (x & 0x0F0F0F0Fu) + ((x >> 4) & 0x0F0F0F0Fu);
		return (x * 0x01010101u) >> 24;
	}

	// 8 row headers (one per nibble of H) -> per nibble: payload bytes / 2 of the bit packed rows (0 for raw and RLE
	// rows), the raw rows (header 15) and the RLE rows (headers 6, 7) as 0/1 nibbles
	__device__ __forceinline__ void classify_headers(uint32_t H, uint32_t& half_pay, uint32_t& raw, uint32_t& rle)
	{
		const uint32_t b0 = H & 0x11111111u, b1 = (H >> 1) & 0x11111111u, b2 = (H >> 2) & 0x11111111u, b3 = (H >> 3) & 0x11111111u;
		raw = b0 & b1 & b2 & b3;
		rle = b2 & b1 & ~b3;
		half_pay = H & 0x77777777u & ~((raw | rle) * 15u);
	}

	// Length of a NORMAL / NORMAL_RLE plane at q (decode_block_flat / decode_block_flat_rle, block_compress.h:2046-2084):
	// 8 header bytes, the mins, 16 rows.  Every byte that decides the length is read below `end`.
	__device__ __forceinline__ uint32_t parse_plane_len(const uint8_t* q, const uint8_t* end, uint32_t kind)
	{
		if (q + 8 + (kind == (uint32_t)KIND_NORMAL_RLE ? 2 : 0) > end)
			return PARSE_BAD;
		uint32_t H0, H1;
		rd64_unaligned(q, H0, H1);
		uint32_t a0, a1, raw0, raw1, rle0, rle1;
		classify_headers(H0, a0, raw0, rle0);
		classify_headers(H1, a1, raw1, rle1);
		const uint32_t nomin = __popc(raw0 | rle0) + __popc(raw1 | rle1);
		const uint32_t mins_len = kind == (uint32_t)KIND_NORMAL_RLE ? 2u + __popc(~rd16(q + 8) & 0xFFFFu) : 16u - nomin;
		const uint8_t* rows = q + 8 + mins_len;
		uint32_t consumed = 2u * (nibble_sum(a0) + nibble_sum(a1)) + 16u * (__popc(raw0) + __popc(raw1));
		// RLE rows ([mask:2][non repeated bytes], :1939-1968) in row order: each starts after everything before it
		uint32_t acc = 0;
		uint32_t todo = rle0;
#pragma unroll 1
		for (int w = 0; w < 2; ++w) {
			while (todo) {
				const uint32_t sh = (uint32_t)__ffs((int)todo) - 1u; // 4 * row (within this word)
				todo &= todo - 1u;
				const uint32_t below = (1u << sh) - 1u;
				uint32_t at = 2u * nibble_sum((w ? a1 : a0) & below) + 16u * __popc((w ? raw1 : raw0) & below) + acc;
				if (w)
					at += 2u * nibble_sum(a0) + 16u * __popc(raw0);
				const uint8_t* mp = rows + at;
				if (mp + 2 > end)
					return PARSE_BAD;
				acc += 2u + __popc(~rd16(mp) & 0xFFFFu);
			}
			todo = rle1;
		}
		consumed += acc;
		if (rows + consumed > end)
			return PARSE_BAD;
		return 8u + mins_len + consumed;
	}

	// Length of an LZ stream (lz_decompress, lz_compress.h:234-277): 32 groups of [anchor][8 items]
	template<int B>
	__device__ __noinline__ uint32_t parse_lz_len(const uint8_t* p, const uint8_t* end)
	{
		const uint8_t* s = p;
		for (uint32_t g = 0; g < 32; ++g) {
			if (s + 2 > end)
				return PARSE_BAD;
			const uint32_t anchor = *s++;
			if (!anchor) {
				if (s + 8 * B > end)
					return PARSE_BAD;
				s += 8 * B;
				continue;
			}
			for (uint32_t k = 0; k < 8; ++k) {
				if ((anchor >> k) & 1u) {
					if (s >= end)
						return PARSE_BAD;
					if (*s++ > 127u) {
						if (s >= end)
							return PARSE_BAD;
						++s;
					}
				}
				else {
					if (s + B > end)
						return PARSE_BAD;
					s += B;
				}
			}
		}
		return (uint32_t)(s - p);
	}

	// Length of one full block at p (block_decompress_sse, block_compress.h:2088-2157)
	template<int T>
	__device__ __forceinline__ uint32_t parse_block_len(const uint8_t* p, const uint8_t* end)
	{
		constexpr uint32_t HS = (T + 1) / 2;
		if (p + HS >= end) // :2114-2116
			return PARSE_BAD;
		const uint32_t marker = p[0];
		if (marker == (uint32_t)MARK_COPY)
			return p + 1 + T * 256 > end ? PARSE_BAD : 1u + T * 256u;
		if (marker == (uint32_t)MARK_LZ) {
			if ((T % 4) != 0)
				return PARSE_BAD;
			const uint32_t r = parse_lz_len<T>(p + 1, end);
			return r == PARSE_BAD ? r : r + 1u;
		}
		uint32_t kinds = 0;
#pragma unroll
		for (uint32_t i = 0; i < HS; ++i)
			kinds |= (uint32_t)p[i] << (8 * i);
		// kinds: one nibble per plane, 0 SAME (1 byte), 1 RAW (256 bytes), 2 / 3 NORMAL / NORMAL_RLE (walked below)
		constexpr uint32_t ONES = T == 8 ? 0x11111111u : ((1u << (4 * T)) - 1u) / 15u;
		if (kinds & (ONES * 12u))
			return PARSE_BAD;
		const uint32_t normal = (kinds >> 1) & ONES, raw = kinds & ~(kinds >> 1) & ONES;
		// fixed part: every SAME plane 1 byte, every RAW plane 256; the NORMAL planes add their lengths in plane order
		uint32_t q = HS + (uint32_t)__popc(ONES & ~normal & ~raw) + 256u * (uint32_t)__popc(raw);
		const uint32_t room = (uint32_t)(end - p);
		if (q > room)
			return PARSE_BAD;
		uint32_t todo = normal;
		while (todo) {
			const uint32_t sh = (uint32_t)__ffs((int)todo) - 1u; // 4 * plane
			todo &= todo - 1u;
			// the plane starts after the planes before it: SAME and RAW ones by count, NORMAL ones already walked.
			// q holds: HS + all SAME + all RAW + NORMAL planes walked so far; take back the SAME / RAW planes that come later.
			const uint32_t later = ~((2u << sh) - 1u);
			const uint32_t at = q - (uint32_t)__popc(ONES & ~normal & ~raw & later) - 256u * (uint32_t)__popc(raw & later);
			const uint32_t r = parse_plane_len(p + at, end, (kinds >> sh) & 15u);
			if (r == PARSE_BAD)
				return r;
			q += r;
			if (q > room)
				return PARSE_BAD;
		}
		return q;
	}

	// one thread walks one superblock (stenos.cpp:1124-1143, :681-753) and publishes its block entries
	template<int T>
	__device__ __forceinline__ void parse_superblock(const SplitParams& S, uint32_t i, void* scratch_word)
	{
		constexpr uint32_t BLOCK = T * 256u;
		const DecodeParams& P = S.d;
		const uint32_t s = P.first_sb + i;
		const uint64_t doff = (uint64_t)s * P.sb_bytes;
		const uint32_t dsize = (uint32_t)min((uint64_t)P.sb_bytes, P.total - doff); // remainder 0 = full superblock (appendix C1)
		const bool last = (doff + dsize == P.total);
		const uint32_t nfull = dsize / BLOCK, rem = dsize - nfull * BLOCK;
		const uint32_t nblk = nfull + (rem ? 1u : 0u);
		unsigned long long* ent = S.entries + (uint64_t)i * S.kmax;
		uint32_t err = 0, k = 0;
		const uint64_t at = P.sb_offsets[s];
		if (at + 4 > P.src_size)
			err = DEV_ERR_SRC_OVERFLOW; // stenos.cpp:1126-1127
		else {
			const uint8_t* p = P.src + at;
			const uint32_t code = p[0], csize = rd24(p + 1);
			const uint8_t* end = p + 4 + csize;
			if (at + 4 + csize > P.src_size)
				err = DEV_ERR_INVALID_INPUT; // stenos.cpp:1133-1134
			else if (code == (uint32_t)CODE_COPY) {
				if (csize != dsize)
					err = DEV_ERR_INVALID_INPUT; // stenos.cpp:743-744
				else
					for (; k < nblk; ++k)
						st_volatile_u64(&ent[k], make_entry(ENT_RAWCOPY, min(BLOCK, dsize - k * BLOCK), at + 4u + k * BLOCK));
			}
			else if (code == (uint32_t)CODE_BLOCK) {
				if (csize == 0 && dsize)
					err = DEV_ERR_INVALID_INPUT;
				else {
					uint32_t q = 4;
					for (; k < nfull; ++k) {
						// the walk should meet L1 on every new line: far ahead the lines start towards L2, nearer towards L1
						if (p + q + 2048 < end)
							prefetch_l2(p + q + 2048);
						if (p + q + 512 < end) {
							prefetch_l1_async(p + q + 384, scratch_word);
							prefetch_l1_async(p + q + 512, scratch_word);
						}
						const uint32_t len = parse_block_len<T>(p + q, end);
						if (len == PARSE_BAD) {
							err = DEV_ERR_INVALID_INPUT;
							break;
						}
						st_volatile_u64(&ent[k], make_entry(0ull, len, at + q));
						q += len;
					}
					if (!err && rem) {
						// [254] + partial block (block_compress.h:2158-2172), checked by its decoder
						st_volatile_u64(&ent[k], make_entry(0ull, csize + 4u - q, at + q));
						++k;
					}
				}
			}
			else if (code == (uint32_t)CODE_ZSTD && P.skip_zstd_tail && last && dsize < 128u) {
				// tiny final superblock (stenos.cpp:435-437): decoded by the host layer through libzstd
			}
			else
				err = DEV_ERR_INVALID_INPUT; // codes 3,4,5 carry Zstd payloads (levels >= 2): out of scope
		}
		for (; k < nblk; ++k)
			st_volatile_u64(&ent[k], ENT_SKIP);
		if (err)
			atomicOr(&P.result[1], (unsigned long long)err);
	}

	// one half-warp copies n raw bytes (n <= T * 256) from q (any alignment) to out (16-byte aligned)
	__device__ __forceinline__ void half_copy_raw(const uint8_t* q, const uint8_t* lim, uint8_t* out, uint32_t n, int r)
	{
		for (uint32_t o = 16u * (uint32_t)r; o < n; o += 256u) {
			if (o + 16u <= n && q + o + 20 <= lim) {
				uint32_t v[4];
				load16_unaligned(q + o, v);
				*reinterpret_cast<uint4*>(out + o) = make_uint4(v[0], v[1], v[2], v[3]);
			}
			else {
				for (uint32_t j = o; j < min(o + 16u, n); ++j)
					out[j] = q[j];
			}
		}
	}

	template<int T>
	__global__ void __launch_bounds__(DECODE2_WARPS * 32, DECODE2_MIN_CTAS) decode_split_kernel(SplitParams S)
	{
		constexpr uint32_t BLOCK = T * 256u;
		constexpr uint32_t HS = (T + 1) / 2;
		STENOS_DYN_SMEM(uint8_t, smem);
		const DecodeParams& P = S.d;
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		uint16_t* lz_scratch = reinterpret_cast<uint16_t*>(smem) + 256 * warp;
		uint32_t* role = reinterpret_cast<uint32_t*>(smem + DECODE2_WARPS * 512);
		void* scratch_word = smem + DECODE2_WARPS * 512 + 16 + 4 * threadIdx.x; // target of the L1 prefetches, never read
		if (threadIdx.x == 0)
			*role = (uint32_t)atomicAdd(P.ticket, 1ull);
		__syncthreads();
		const uint32_t cta = *role;
		if (cta < S.n_parse_cta) {
			const uint32_t i = cta * (uint32_t)(DECODE2_WARPS * 32) + threadIdx.x;
			if (i < P.n_sb && S.parse_only != 2u)
				parse_superblock<T>(S, i, scratch_word);
			return;
		}

		if (S.parse_only == 1u)
			return;
		const int half = lane >> 4, r = lane & 15, hsh = 16 * half;
		const uint8_t* lim = P.src + P.src_size;
		const unsigned long long units = (unsigned long long)P.n_sb * S.kmax;
		// Decoder warps take unit pairs round robin (warp w: pairs w, w + W, ...) and run a three stage software
		// pipeline over them, so that no decode waits for memory it could have asked for earlier:
		//   pair n + 2: the entry load is issued        pair n + 1: the entry is resolved, the block's lines are requested
		//   pair n    : decoded
		const unsigned long long n_warps = (unsigned long long)(gridDim.x - S.n_parse_cta) * DECODE2_WARPS;
		const unsigned long long w0 = (unsigned long long)(cta - S.n_parse_cta) * DECODE2_WARPS + warp;

		struct Unit
		{
			uint32_t k, i, want; // block k of superblock i; decoded bytes (0: nothing to do)
			unsigned long long e; // entry (0: not read / not published yet)
		};
		// which block pair n of this warp is, for this half; issues the entry load
		auto locate = [&](unsigned long long n, Unit& U) {
			const unsigned long long u = 2ull * (w0 + n * n_warps) + (unsigned long long)half;
			U.k = U.i = U.want = 0;
			U.e = 0ull;
			if (u >= units)
				return;
			if (units <= 0xFFFFFFFFull) {
				U.k = (uint32_t)u / P.n_sb;
				U.i = (uint32_t)u - U.k * P.n_sb;
			}
			else {
				U.k = (uint32_t)(u / P.n_sb);
				U.i = (uint32_t)(u - (unsigned long long)U.k * P.n_sb);
			}
			const uint64_t doff = (uint64_t)(P.first_sb + U.i) * P.sb_bytes;
			const uint32_t dsize = (uint32_t)min((uint64_t)P.sb_bytes, P.total - doff); // remainder 0 = full superblock (appendix C1)
			if ((uint64_t)U.k * BLOCK < dsize) {
				U.want = min(BLOCK, dsize - U.k * BLOCK);
				if (r == 0)
					U.e = ld_volatile_u64(S.entries + (uint64_t)U.i * S.kmax + U.k);
			}
		};
		// waits for the parser if it has to, then asks for exactly the block's lines, all misses in flight together:
		// the dependent loads of the decode meet L1
		auto resolve = [&](Unit& U) {
			for (;;) {
				if (U.want && U.e == 0ull && r == 0)
					U.e = ld_volatile_u64(S.entries + (uint64_t)U.i * S.kmax + U.k);
				U.e = __shfl_sync(FULL, U.e, 0, 16);
				if (!__any_sync(FULL, U.want != 0u && U.e == 0ull))
					break;
				STENOS_SPIN_HINT();
			}
			if (U.e & ENT_SKIP)
				U.want = 0;
			if (U.want) {
				const uint8_t* q = P.src + (U.e & ENT_OFFSET);
				const uint8_t* end = q + ((uint32_t)(U.e >> 48) & ENT_LEN_MAX);
				const uint8_t* l0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)127);
				for (const uint8_t* l = l0 + 128 * r; l < end && l < lim; l += 128 * 16)
					prefetch_l1_async(l < P.src ? P.src : l, scratch_word);
			}
		};

		Unit U0, U1, U2;
		locate(0, U1);
		resolve(U1);
		locate(1, U2);
		for (unsigned long long n = 0;; ++n) {
			if (2ull * (w0 + n * n_warps) >= units)
				break;
			U0 = U1;
			U1 = U2;
			resolve(U1);
			locate(n + 2, U2);
			bool act = U0.want != 0u;
			if (!__any_sync(FULL, act))
				continue;
			const unsigned long long e = U0.e;
			const uint32_t k = U0.k, want = U0.want;
			const uint64_t doff = (uint64_t)(P.first_sb + U0.i) * P.sb_bytes;
			const uint8_t* q = P.src + (e & ENT_OFFSET);
			const uint32_t len = (uint32_t)(e >> 48) & ENT_LEN_MAX;
			const uint8_t* end = q + len;
			uint8_t* out = P.dst + (doff - P.dst_origin) + (size_t)k * BLOCK;
			uint32_t err = 0;
			if (act && (e & ENT_RAWCOPY)) {
				half_copy_raw(q, lim, out, want, r);
				act = false;
			}
			// full blocks: the row-per-lane decoder unless the block is LZ / COPY coded or too close to the end of the buffer
			const bool full = act && want == BLOCK;
			bool fast = full && q + HS < end && q + WorstBlock<T>::READ <= lim;
			const uint32_t marker = fast ? (uint32_t)q[0] : 0u;
			fast = fast && marker < (uint32_t)MARK_COPY;
			const bool slow = act && !fast;
			if (__any_sync(FULL, fast)) {
				const uint32_t c = decode_block_rows<T>(q, fast, r, hsh, out);
				if (fast && (c == 0xFFFFFFFFu || c > (uint32_t)(end - q)))
					err = DEV_ERR_INVALID_INPUT;
			}
			if (__any_sync(FULL, slow)) {
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					if (!__shfl_sync(FULL, (int)slow, 16 * hh))
						continue;
					const uint8_t* q_h = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)q, 16 * hh));
					const uint8_t* end_h = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)end, 16 * hh));
					uint8_t* o_h = reinterpret_cast<uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)out, 16 * hh));
					const uint32_t want_h = __shfl_sync(FULL, want, 16 * hh);
					uint32_t bad = 0;
					if (want_h == BLOCK) {
						const uint32_t c = decode_block<T>(q_h, end_h, lim, o_h, lz_scratch, lane);
						bad = (c == 0xFFFFFFFFu || c > (uint32_t)(end_h - q_h));
					}
					else if (q_h >= end_h || *q_h != (uint8_t)MARK_PARTIAL) // block_compress.h:2160-2166
						bad = 1;
					else
						bad = decode_partial_block<T>(q_h + 1, end_h, lim, o_h, want_h, lane) == 0xFFFFFFFFu;
					if (half == hh && bad)
						err = DEV_ERR_INVALID_INPUT;
				}
			}
			if (err && r == 0)
				atomicOr(&P.result[1], (unsigned long long)err);
		}
	}
}
