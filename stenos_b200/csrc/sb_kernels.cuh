// sb_kernels.cuh -- frame-level kernels of the level-1 path (device side) and their launchers.
//
//   encode_frame_kernel   stenos_compress_generic's superblock loop (stenos.cpp:893-904) +
//                         compress_generic_superblock level-1 branch (:403-450, :606-615, :363-374):
//                         persistent CTAs, one superblock at a time, every 256-element block encoded
//                         by one warp into a shared-memory slot, in-CTA scan of the block sizes,
//                         decoupled look-back across superblocks for the frame offsets, then one
//                         coalesced copy-out.  HBM traffic = N (read) + C (write).
//   frame_index_kernel    the serial walk over [code][csize:3] headers (stenos.cpp:1124-1143)
//   decode_frame_kernel   decompress_generic_superblock codes 1 and 6 (:681-753): one warp per
//                         superblock, blocks in sequence (the stream is not seekable)
//   shuffle / delta       filters (see sb_filters.cuh)
#pragma once
#include "sb_common.cuh"
#include "sb_encode.cuh"
#include "sb_encode_rows.cuh"
#include "sb_decode.cuh"

namespace sb
{
	// ------------------------------------------------------------------------------------------
	// byte copies with arbitrary alignment on both sides
	// ------------------------------------------------------------------------------------------

	// one warp copies n bytes src -> dst (dst global, src shared or global).  4-byte stores on the
	// destination's natural alignment, bytes at both ends.
	__device__ __forceinline__ void warp_copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, int lane)
	{
		const uint32_t head = min((uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u), n);
		if ((uint32_t)lane < head)
			dst[lane] = src[lane];
		const uint32_t body = (n - head) >> 2;
		const uint8_t* s = src + head;
		uint32_t* d = reinterpret_cast<uint32_t*>(dst + head);
		const uintptr_t sa = (uintptr_t)s;
		const uint32_t sh = (uint32_t)(sa & 3u) * 8u;
		const uint32_t* sw = reinterpret_cast<const uint32_t*>(sa & ~(uintptr_t)3);
		if (sh == 0) {
			for (uint32_t j = lane; j < body; j += 32)
				d[j] = sw[j];
		}
		else {
			for (uint32_t j = lane; j < body; j += 32)
				d[j] = __funnelshift_r(sw[j], sw[j + 1], sh);
		}
		const uint32_t done = head + (body << 2);
		if (done + (uint32_t)lane < n)
			dst[done + lane] = src[done + lane];
	}

	// the whole CTA copies n bytes global -> global
	__device__ __forceinline__ void cta_copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, int tid, int nthreads)
	{
		const uint32_t head = min((uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u), n);
		if ((uint32_t)tid < head)
			dst[tid] = src[tid];
		const uint32_t body = (n - head) >> 4;
		const uint8_t* s = src + head;
		uint4* d = reinterpret_cast<uint4*>(dst + head);
		const uintptr_t sa = (uintptr_t)s;
		const uint32_t sh = (uint32_t)(sa & 3u) * 8u;
		const uint32_t* sw = reinterpret_cast<const uint32_t*>(sa & ~(uintptr_t)3);
		if ((sa & 15u) == 0) {
			const uint4* s4 = reinterpret_cast<const uint4*>(s);
			for (uint32_t j = tid; j < body; j += nthreads)
				d[j] = s4[j];
		}
		else {
			for (uint32_t j = tid; j < body; j += nthreads) {
				const uint32_t a = sw[4 * j], b = sw[4 * j + 1], c = sw[4 * j + 2], e = sw[4 * j + 3];
				const uint32_t f = sh ? sw[4 * j + 4] : 0u;
				d[j] = make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh), __funnelshift_r(c, e, sh), __funnelshift_r(e, f, sh));
			}
		}
		const uint32_t done = head + (body << 4);
		if (done + (uint32_t)tid < n)
			dst[done + tid] = src[done + tid];
	}

	// ------------------------------------------------------------------------------------------
	// encoder
	// ------------------------------------------------------------------------------------------
	struct EncodeParams
	{
		const uint8_t* src;  // uncompressed input (16-byte aligned)
		uint64_t bytes;      // bytes handled on the device (whole superblocks >= 128 bytes)
		uint8_t* dst;        // frame output
		uint64_t dst_size;   // capacity of dst
		uint32_t sb_bytes;   // superblock size
		uint32_t n_sb;       // superblocks handled by this launch
		uint32_t header_len; // frame header bytes preceding the first superblock (8 or 12); 0 = do not write it
		uint32_t shift_byte; // first header byte
		uint64_t frame_bytes; // decompressed size written in the frame header (may exceed `bytes` when the host adds a Zstd tail)
		int level;           // 0 (all COPY) or 1
		unsigned long long* state; // [n_sb] look-back words, zero initialised
		uint32_t* ticket;          // zero initialised
		unsigned long long* result; // [0] offset past the last superblock written, [1] device error bits
		unsigned long long* sb_offsets; // optional [n_sb + 1]: offset of every superblock header in dst
		uint64_t base_offset; // offset in dst of this launch's first superblock when header_len == 0 (multi-GPU segments)
		uint32_t first_sb;    // encode_frame_kernel: superblocks [first_sb, n_sb) (the ones before were placed by encode_stream_kernel)
		uint32_t n_stream;    // encode_stream_kernel / encode_flow_kernel: superblocks [0, n_stream)
		uint8_t* spill;       // encode_flow_kernel: HBM spill slots (FlowLayout::spill_bytes(grid))
		uint32_t ring_cap;    // encode_flow_kernel: 0, or (tests) bytes of each staging ring to use
		// encode_frame_kernel in BUCKET mode (stenos_b200_compress_buckets_async, cvector.hpp:1394-1420): superblock s is
		// the bare, independent superblock of bucket ids[s] (or s) written at dst + s * bucket_stride with
		// dst_size = bucket_stride, as one stenos_private_compress_block call per bucket would; no chain, no frame header
		uint32_t bucket_stride = 0;            // 0: frame mode
		const uint32_t* bucket_ids = nullptr;  // optional [n_sb]
		uint32_t* bucket_sizes = nullptr;      // [n_sb]: 4 + csize, or 0 when the slot is too small
	};

	constexpr unsigned long long LB_AGGREGATE = 1ull << 62;
	constexpr unsigned long long LB_INCLUSIVE = 2ull << 62;
	constexpr unsigned long long LB_VALUE = (1ull << 62) - 1;

	// MB: most full blocks of a superblock (the shared-memory slots are sized for it)
	template<int T, uint32_t MB = DEFAULT_SUPERBLOCK / (T * 256u)>
	struct EncodeLayout
	{
		static constexpr uint32_t BLOCK = T * 256u;
		static constexpr uint32_t HS = (T + 1) / 2;
		// worst full block BLOCK + HS (LZ: 1 + BLOCK); worst partial block 1 + HS + 8 * T + (BLOCK - 1); 16-byte aligned
		static constexpr uint32_t STRIDE = (BLOCK + HS + 8u * T + 1u + 15u) & ~15u;
		static constexpr uint32_t MAX_BLOCKS = MB;
		static constexpr uint32_t SLOTS_BYTES = MAX_BLOCKS * STRIDE;
		static constexpr uint32_t NENT = MAX_BLOCKS + 1; // + partial
		// [slots][sizes u32 x NENT][offsets u32 x (NENT+1)][misc 16 x u64][lz scratch per warp]
		static constexpr uint32_t SIZES_OFF = SLOTS_BYTES;
		static constexpr uint32_t OFFS_OFF = SIZES_OFF + ((NENT * 4u + 15u) & ~15u);
		static constexpr uint32_t MISC_OFF = OFFS_OFF + (((NENT + 1) * 4u + 15u) & ~15u);
		static constexpr uint32_t LZ_OFF = MISC_OFF + 128u;
		static constexpr uint32_t LZ_STRIDE = (LZ_SCRATCH_BYTES + 15u) & ~15u;
		static uint32_t smem_bytes(int nwarps) { return LZ_OFF + LZ_STRIDE * (uint32_t)nwarps; }
	};

	// worst case length of a superblock's block stream
	template<int T>
	__device__ __forceinline__ uint32_t worst_stream(uint32_t nfull, uint32_t rem)
	{
		return nfull * (T * 256u + (T + 1) / 2) + (rem ? 1u + (T + 1) / 2 + 8u * T + rem : 0u);
	}

	template<int T, int NT, uint32_t MB = DEFAULT_SUPERBLOCK / (T * 256u)>
	__global__ void __launch_bounds__(NT, MB == DEFAULT_SUPERBLOCK / (T * 256u) ? 1 : 8) encode_frame_kernel(EncodeParams P)
	{
		using L = EncodeLayout<T, MB>;
		const bool buckets = P.bucket_stride != 0u;
		STENOS_DYN_SMEM(uint8_t, smem);
		uint32_t* sizes = reinterpret_cast<uint32_t*>(smem + L::SIZES_OFF);
		uint32_t* offs = reinterpret_cast<uint32_t*>(smem + L::OFFS_OFF);
		unsigned long long* misc = reinterpret_cast<unsigned long long*>(smem + L::MISC_OFF);
		const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
		const int nwarps = blockDim.x >> 5;
		uint32_t* lz_scratch = reinterpret_cast<uint32_t*>(smem + L::LZ_OFF + L::LZ_STRIDE * warp);
		const uint64_t first_off = P.header_len ? (uint64_t)P.header_len : P.base_offset;

		// tickets are fetched one superblock ahead so the atomic's latency hides behind the encoding
		uint32_t next_ticket = 0;
		if (tid == 0)
			next_ticket = P.first_sb + atomicAdd(P.ticket, 1u);
		for (;;) {
			// ---- next superblock (ticket order == look-back order: predecessors are always running or done)
			__syncthreads();
			if (tid == 0)
				misc[0] = next_ticket;
			__syncthreads();
			const uint32_t s = (uint32_t)misc[0];
			if (s >= P.n_sb)
				break;
			if (tid == 0)
				next_ticket = P.first_sb + atomicAdd(P.ticket, 1u);
			const uint64_t in_at = (uint64_t)((buckets && P.bucket_ids) ? P.bucket_ids[s] : s) * P.sb_bytes;
			const uint8_t* in = P.src + in_at;
			const uint32_t in_bytes = in_at < P.bytes ? (uint32_t)min((uint64_t)P.sb_bytes, P.bytes - in_at) : 0u;
			const uint32_t nfull = in_bytes / L::BLOCK, rem = in_bytes - nfull * L::BLOCK;
			const uint32_t nent = nfull + (rem ? 1u : 0u);

			// ---- can the reference's dst room checks change anything for this superblock?
			// lower bound of the room: every earlier superblock stored as COPY
			const uint32_t need = worst_stream<T>(nfull, rem) + 8u * T + 32u;
			const long long room_lb = buckets ? (long long)P.bucket_stride - 4 : (long long)P.dst_size - (long long)(first_off + (uint64_t)s * (4ull + P.sb_bytes)) - 4;
			bool have_base = false, exact = false;
			uint64_t base = 0; // offset of this superblock's header in dst
			long long room = room_lb;
			if (buckets) {
				// the bucket's slot is its whole dst (stenos.cpp:768-778): base and room are known, nothing to wait for
				base = (uint64_t)s * P.bucket_stride;
				have_base = true;
				exact = P.level != 0 && room < (long long)need;
			}
			else if (P.level != 0 && room_lb < (long long)need) {
				// wait for the true offset first
				if (tid == 0) {
					uint64_t excl = 0;
					if (s > 0) {
						unsigned long long v;
						while (((v = ld_volatile_u64(&P.state[s - 1])) & LB_INCLUSIVE) == 0)
							STENOS_SPIN_HINT();
						excl = v & LB_VALUE;
					}
					misc[1] = first_off + excl;
				}
				__syncthreads();
				base = misc[1];
				have_base = true;
				room = (long long)P.dst_size - (long long)base - 4;
				exact = room < (long long)need;
			}

			bool err = false;
			if (P.level == 0) {
			}
			else if (!exact && !(T == 2 || T == 4 || T == 8)) {
				// other element sizes: one warp per block, one lane per half row (sb_encode.cuh)
				for (uint32_t b = warp; b < nfull; b += nwarps) {
					bool e = false;
					const uint32_t sz = encode_block<T, false>(in + (size_t)b * L::BLOCK, smem + b * L::STRIDE, lz_scratch, lane, 0xFFFFFFFFu, e);
					if (lane == 0)
						sizes[b] = sz;
				}
				if (rem && warp == (int)(nfull % nwarps)) {
					bool e = false;
					const uint32_t sz = encode_partial_block<T, false>(in + (size_t)nfull * L::BLOCK, rem, smem + nfull * L::STRIDE, lane, 0xFFFFFFFFu, e);
					if (lane == 0)
						sizes[nfull] = sz;
				}
			}
			else if (!exact) {
				// full blocks in pairs: one warp per pair, one lane per 16-element row (sb_encode_rows.cuh)
				const uint32_t npairs = (nfull + 1u) >> 1;
				for (uint32_t pr = warp; pr < npairs; pr += nwarps) {
					const uint32_t b = 2u * pr;
					const bool second = b + 1u < nfull;
					uint8_t* slot0 = smem + b * L::STRIDE;
					const uint32_t sz = encode_block_pair<T>(in + (size_t)b * L::BLOCK, second, slot0, L::STRIDE, lz_scratch, lane,
										 [&](uint32_t) -> uint8_t* { return ((lane >> 4) && second) ? slot0 + L::STRIDE : slot0; });
					if ((lane & 15) == 0 && (lane == 0 || second))
						sizes[b + (lane >> 4)] = sz;
				}
				if (rem && warp == (int)(npairs % nwarps)) {
					bool e = false;
					const uint32_t sz = encode_partial_block<T, false>(in + (size_t)nfull * L::BLOCK, rem, smem + nfull * L::STRIDE, lane, 0xFFFFFFFFu, e);
					if (lane == 0)
						sizes[nfull] = sz;
				}
			}
			else if (warp == 0) {
				// exact mode: the blocks of the superblock in sequence with the reference's room arithmetic
				uint32_t off = 0;
				for (uint32_t b = 0; b < nfull && !err; ++b) {
					const uint32_t r = room > (long long)off ? (uint32_t)((room - (long long)off) > 0x7FFFFFFFll ? 0x7FFFFFFFll : (room - (long long)off)) : 0u;
					const uint32_t sz = encode_block<T, true>(in + (size_t)b * L::BLOCK, smem + b * L::STRIDE, lz_scratch, lane, r, err);
					if (lane == 0)
						sizes[b] = sz;
					off += sz;
				}
				if (rem && !err) {
					const uint32_t r = room > (long long)off ? (uint32_t)((room - (long long)off) > 0x7FFFFFFFll ? 0x7FFFFFFFll : (room - (long long)off)) : 0u;
					const uint32_t sz = encode_partial_block<T, true>(in + (size_t)nfull * L::BLOCK, rem, smem + nfull * L::STRIDE, lane, r, err);
					if (lane == 0)
						sizes[nfull] = sz;
				}
				if (lane == 0)
					misc[2] = err ? 1ull : 0ull;
			}
			__syncthreads();
			if (exact)
				err = misc[2] != 0ull;

			// ---- in-CTA exclusive scan of the block sizes (warp 0)
			if (warp == 0 && P.level != 0 && !err) {
				const uint32_t per = (nent + 31u) / 32u;
				const uint32_t lo_i = (uint32_t)lane * per;
				uint32_t sum = 0;
				for (uint32_t i = lo_i; i < min(lo_i + per, nent); ++i)
					sum += sizes[i];
				uint32_t incl = sum;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					uint32_t t = __shfl_up_sync(FULL, incl, d);
					if (lane >= d)
						incl += t;
				}
				uint32_t run = incl - sum;
				for (uint32_t i = lo_i; i < min(lo_i + per, nent); ++i) {
					offs[i] = run;
					run += sizes[i];
				}
				if (lane == 31)
					offs[nent] = incl;
			}
			__syncthreads();
			const uint32_t csize = (P.level != 0 && !err) ? offs[nent] : 0xFFFFFFFFu;
			const bool copy = csize > in_bytes; // stenos.cpp:609-610 (error or larger than the input -> MEMCPY)
			const uint32_t out_size = 4u + (copy ? in_bytes : csize);

			// ---- frame offset: publish the size, then a warp-wide decoupled look-back (32 predecessors per probe)
			if (warp == 0) {
				uint64_t excl = 0;
				if (have_base)
					excl = base - first_off;
				else {
					if (lane == 0)
						st_volatile_u64(&P.state[s], LB_AGGREGATE | (unsigned long long)out_size);
					long long top = (long long)s - 1;
					while (top >= 0) {
						const long long idx = top - lane;
						// entries before the first superblock act as an inclusive prefix of 0
						const unsigned long long v = idx >= 0 ? ld_volatile_u64(&P.state[idx]) : LB_INCLUSIVE;
						const uint32_t inc = __ballot_sync(FULL, (v & LB_INCLUSIVE) != 0ull);
						const uint32_t inv = __ballot_sync(FULL, (v >> 62) == 0ull);
						const int fi = inc ? (__ffs((int)inc) - 1) : 32; // nearest predecessor that already knows its prefix
						const uint32_t upto = fi >= 31 ? 0xFFFFFFFFu : ((2u << fi) - 1u);
						if (inv & upto) { // a predecessor in the window has not published yet
							STENOS_SPIN_HINT();
							continue;
						}
						unsigned long long val = ((upto >> lane) & 1u) ? (v & LB_VALUE) : 0ull;
#pragma unroll
						for (int d = 16; d >= 1; d >>= 1)
							val += __shfl_xor_sync(FULL, val, d);
						excl += val;
						if (fi < 32)
							break;
						top -= 32;
					}
				}
				if (lane == 0 && buckets)
					misc[1] = base;
				else if (lane == 0) {
					st_volatile_u64(&P.state[s], LB_INCLUSIVE | (unsigned long long)(excl + out_size));
					misc[1] = first_off + excl;
					if (P.sb_offsets) {
						P.sb_offsets[s] = first_off + excl;
						if (s == P.n_sb - 1)
							P.sb_offsets[s + 1] = first_off + excl + out_size;
					}
					if (s == P.n_sb - 1)
						P.result[0] = first_off + excl + out_size;
				}
			}
			__syncthreads();
			base = misc[1];
			if (buckets ? (out_size > P.bucket_stride || in_bytes == 0u) : (base + out_size > P.dst_size)) {
				// stenos.cpp:366-367 / :611-612: the caller's buffer is too small
				if (tid == 0) {
					atomicOr(&P.result[1], (unsigned long long)DEV_ERR_DST_OVERFLOW);
					if (buckets)
						P.bucket_sizes[s] = 0u;
				}
				continue;
			}
			if (buckets && tid == 0)
				P.bucket_sizes[s] = out_size;

			// ---- copy out: [code][csize:3] + payload
			uint8_t* out = P.dst + base;
			if (tid == 0) {
				const uint32_t len = copy ? in_bytes : csize;
				out[0] = (uint8_t)(copy ? CODE_COPY : CODE_BLOCK);
				out[1] = (uint8_t)len;
				out[2] = (uint8_t)(len >> 8);
				out[3] = (uint8_t)(len >> 16);
				if (s == 0 && P.header_len) {
					// frame header (stenos.cpp:862-874): [shift][decompressed bytes:7]([superblock bytes:4])
					P.dst[0] = (uint8_t)P.shift_byte;
					for (int i = 0; i < 7; ++i)
						P.dst[1 + i] = (uint8_t)(P.frame_bytes >> (8 * i));
					if (P.header_len == 12)
						for (int i = 0; i < 4; ++i)
							P.dst[8 + i] = (uint8_t)(P.sb_bytes >> (8 * i));
				}
			}
			if (copy)
				cta_copy_bytes(out + 4, in, in_bytes, tid, blockDim.x);
			else
				for (uint32_t b = warp; b < nent; b += nwarps)
					warp_copy_bytes(out + 4 + offs[b], smem + b * L::STRIDE, sizes[b], lane);
		}
	}

	// ------------------------------------------------------------------------------------------
	// cvector buckets of ONE 256-element block, two per warp (stenos_b200_compress_buckets_async): the lane-per-row pair
	// encoder first -- it assumes ample room -- and its result stands whenever the reference's room checks could not have
	// fired for it: every check of block_compress.h:1214-1248 compares (bytes so far) + 16 or full + 8 T + 2 with the room,
	// so "size + 8 T + 32 <= room" and "not an LZ block" (LZ is what the room decides about, :1214) make the exact and the
	// ample-room encodings the same bytes.  Everything else -- incompressible buckets next to their slot's end, LZ blocks,
	// the array's partial last bucket -- is redone by the room-exact encoder, one bucket per warp, as in the bucket mode of
	// encode_frame_kernel.  Output as there: [code][csize:3][payload] at dst + i * bucket_stride, sizes[i] = 4 + csize.
	// ------------------------------------------------------------------------------------------
	constexpr int BUCKET_WARPS = 4;
	// (launch bounds of 5 / 6 CTAs per SM -- 96 / 80 registers, spills -- were measured: 1.113 / 1.115 ms against 1.108 ms)
	template<int T>
	__global__ void __launch_bounds__(BUCKET_WARPS * 32) encode_bucket_pairs_kernel(EncodeParams P)
	{
		using L = EncodeLayout<T, 1>;
		constexpr uint32_t PER_WARP = 2u * L::STRIDE + L::LZ_STRIDE;
		STENOS_DYN_SMEM(uint8_t, smem);
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, hb = lane >> 4;
		uint8_t* slots = smem + (uint32_t)warp * PER_WARP;
		uint32_t* lz_scratch = reinterpret_cast<uint32_t*>(slots + 2u * L::STRIDE);
		const uint32_t n_pairs = (P.n_sb + 1u) / 2u;
		const long long room = (long long)P.bucket_stride - 4;
		for (uint32_t pr = blockIdx.x * BUCKET_WARPS + warp; pr < n_pairs; pr += gridDim.x * BUCKET_WARPS) {
			uint64_t in_at[2];
			uint32_t in_bytes[2];
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const uint32_t i = 2u * pr + h;
				in_at[h] = i < P.n_sb ? (uint64_t)(P.bucket_ids ? P.bucket_ids[i] : i) * P.sb_bytes : 0ull;
				in_bytes[h] = (i < P.n_sb && in_at[h] < P.bytes) ? (uint32_t)min((uint64_t)P.sb_bytes, P.bytes - in_at[h]) : 0u;
			}
			const bool full0 = in_bytes[0] == L::BLOCK, full1 = in_bytes[1] == L::BLOCK;
			uint32_t size[2] = { 0u, 0u };
			bool done[2] = { false, false };
			if (P.level != 0 && full0) {
				const uint32_t sz = encode_block_pair<T>(P.src + in_at[0], full1, slots, L::STRIDE, lz_scratch, lane,
									 [&](uint32_t) -> uint8_t* { return (hb && full1) ? slots + L::STRIDE : slots; }, P.src + in_at[1]);
				__syncwarp();
				const uint32_t s0 = __shfl_sync(FULL, sz, 0), s1 = __shfl_sync(FULL, sz, 16);
				size[0] = s0;
				size[1] = s1;
				done[0] = (long long)s0 + 8 * T + 32 <= room && slots[0] != (uint8_t)MARK_LZ;
				done[1] = full1 && (long long)s1 + 8 * T + 32 <= room && slots[L::STRIDE] != (uint8_t)MARK_LZ;
			}
#pragma unroll 1
			for (int h = 0; h < 2; ++h) {
				const uint32_t i = 2u * pr + h;
				if (i >= P.n_sb)
					break;
				uint8_t* slot = slots + (uint32_t)h * L::STRIDE;
				const uint8_t* in = P.src + in_at[h];
				bool err = false;
				uint32_t csize = size[h];
				if (P.level == 0 || in_bytes[h] == 0u)
					err = true; // level 0: COPY (stenos.cpp:432); an id past the array: reported below
				else if (!done[h]) {
					// the room-exact encoder (one warp per bucket)
					__syncwarp();
					const uint32_t r = room > 0 ? (uint32_t)room : 0u;
					if (in_bytes[h] == L::BLOCK)
						csize = encode_block<T, true>(in, slot, lz_scratch, lane, r, err);
					else
						csize = encode_partial_block<T, true>(in, in_bytes[h], slot, lane, r, err);
					__syncwarp();
				}
				const bool copy = err || csize > in_bytes[h]; // stenos.cpp:609-610
				const uint32_t len = copy ? in_bytes[h] : csize;
				const uint32_t out_size = 4u + len;
				if (out_size > P.bucket_stride || in_bytes[h] == 0u) {
					if (lane == 0) {
						atomicOr(&P.result[1], (unsigned long long)DEV_ERR_DST_OVERFLOW);
						P.bucket_sizes[i] = 0u;
					}
					continue;
				}
				uint8_t* out = P.dst + (uint64_t)i * P.bucket_stride;
				if (lane == 0) {
					out[0] = (uint8_t)(copy ? CODE_COPY : CODE_BLOCK);
					out[1] = (uint8_t)len;
					out[2] = (uint8_t)(len >> 8);
					out[3] = (uint8_t)(len >> 16);
					P.bucket_sizes[i] = out_size;
				}
				if (copy)
					warp_copy_bytes(out + 4, in, len, lane);
				else
					warp_copy_bytes(out + 4, slot, len, lane);
			}
			__syncwarp();
		}
	}

	// ------------------------------------------------------------------------------------------
	// frame index: offsets of the superblock headers of a frame that lives in device memory
	// ------------------------------------------------------------------------------------------
	struct IndexParams
	{
		const uint8_t* src;
		uint64_t src_size;
		uint64_t first;  // offset of the first superblock header
		uint32_t n_sb;
		unsigned long long* sb_offsets; // [n_sb + 1]
		unsigned long long* result;     // [1] error bits
	};

	__global__ void frame_index_kernel(IndexParams P)
	{
		if (threadIdx.x != 0 || blockIdx.x != 0)
			return;
		uint64_t at = P.first;
		for (uint32_t i = 0; i < P.n_sb; ++i) {
			P.sb_offsets[i] = at;
			if (at + 4 > P.src_size) {
				atomicOr(&P.result[1], (unsigned long long)DEV_ERR_SRC_OVERFLOW);
				for (uint32_t j = i; j <= P.n_sb; ++j)
					P.sb_offsets[j] = P.src_size;
				return;
			}
			at += 4ull + rd24(P.src + at + 1);
		}
		P.sb_offsets[P.n_sb] = at;
		if (at > P.src_size)
			atomicOr(&P.result[1], (unsigned long long)DEV_ERR_INVALID_INPUT);
	}

	// ------------------------------------------------------------------------------------------
	// Parallel frame index.  The header chain is a linked list (each [code][csize:3] gives the next
	// header), which the reference walks serially (stenos.cpp:1124-1143): 8192 dependent DRAM reads
	// per GiB.  Here the frame is cut in segments; one warp per segment RE-SYNCHRONISES on the chain
	// by scanning for the first position that looks like a level-1 header and keeps looking like one
	// for INDEX_HOPS hops, then walks (counts) only its own segment.  Plausible is not proof (payload
	// bytes can imitate headers), so a merge kernel accepts the result only if every segment's walk
	// lands exactly on the next segment's start, the count equals the superblock count and the chain
	// ends inside the frame; segment 0 starts at the true first header, so by induction every accepted
	// header is a true one.  A third kernel re-walks the segments and writes the offsets.  Any
	// inconsistency falls back to the serial walk on the device.
	// ------------------------------------------------------------------------------------------
	struct FastIndexParams
	{
		const uint8_t* src;
		uint64_t src_size;
		uint64_t first;      // offset of the first superblock header
		uint32_t n_sb;
		uint32_t max_csize;  // largest plausible payload (the superblock size)
		uint32_t seg_bytes;
		uint32_t n_seg;
		unsigned long long* seg_start; // [n_seg]
		unsigned long long* seg_end;   // [n_seg]
		uint32_t* seg_count;           // [n_seg]
		uint32_t* seg_base;            // [n_seg] exclusive prefix of the counts
		uint32_t* ok;                  // [1] set by the merge kernel when the parallel result is accepted
		unsigned long long* sb_offsets; // [n_sb + 1]
		unsigned long long* result;
	};
	constexpr unsigned long long IDX_NONE = ~0ull;
	constexpr unsigned long long IDX_BROKEN = ~0ull - 1;
	constexpr int INDEX_HOPS = 6;

	// the walk: next header after the one at `at` (bounds only), or IDX_BROKEN
	__device__ __forceinline__ unsigned long long index_next(const uint8_t* src, uint64_t size, uint64_t at)
	{
		if (at + 4 > size)
			return IDX_BROKEN;
		const uint64_t end = at + 4 + rd24(src + at + 1);
		return end > size ? IDX_BROKEN : end;
	}
	// the re-synchronisation filter: could `at` be a level-1 superblock header?  (A.2: code 1 with
	// 1 <= csize <= superblock, code 6 with csize == superblock or ending the frame, code 2 only as
	// the tiny final superblock.)  Frames holding Zstd codes 3-5 never synchronise and take the serial walk.
	__device__ __forceinline__ bool index_plausible(const uint8_t* src, uint64_t size, uint64_t at, uint32_t max_csize, unsigned long long& next)
	{
		if (at + 4 > size)
			return false;
		const uint32_t code = src[at];
		const uint32_t csize = rd24(src + at + 1);
		const uint64_t end = at + 4 + csize;
		next = end;
		if (end > size)
			return false;
		if (code == (uint32_t)CODE_BLOCK)
			return csize != 0u && csize <= max_csize;
		if (code == (uint32_t)CODE_COPY)
			return csize == max_csize || end == size;
		if (code == (uint32_t)CODE_ZSTD)
			return end == size && csize <= 256u;
		return false;
	}
	__device__ __forceinline__ bool index_chain_plausible(const uint8_t* src, uint64_t size, uint64_t at, uint32_t max_csize)
	{
		unsigned long long nx = 0;
		for (int hop = 0; hop < INDEX_HOPS; ++hop) {
			if (!index_plausible(src, size, at, max_csize, nx))
				return false;
			if (nx == size)
				return true; // the chain ends exactly at the end of the frame
			at = nx;
		}
		return true;
	}

	constexpr int INDEX_WARPS = 4;

	__global__ void __launch_bounds__(INDEX_WARPS * 32) index_scan_kernel(FastIndexParams P)
	{
		const int lane = threadIdx.x & 31;
		const uint32_t k = blockIdx.x * INDEX_WARPS + (threadIdx.x >> 5);
		if (k >= P.n_seg)
			return;
		const uint64_t lo = P.first + (uint64_t)k * P.seg_bytes;
		const uint64_t hi = min(lo + (uint64_t)P.seg_bytes, P.src_size); // walk while pos < hi
		unsigned long long start = IDX_NONE;
		if (k == 0)
			start = P.first;
		else {
			// a true header lies within one superblock (4 + max_csize bytes) of any position inside the chain
			const uint64_t scan_end = min(P.src_size, (uint64_t)(lo + 4ull + P.max_csize));
			for (uint64_t x0 = lo; x0 < scan_end; x0 += 32 * 16) {
				// lane scans 16 consecutive positions; candidates first by their code byte (SWAR), then the chain
				const uint64_t xb = x0 + 16ull * lane;
				unsigned long long found = IDX_NONE;
				uint32_t cand = 0;
				uint32_t w[5] = { 0u, 0u, 0u, 0u, 0u };
				if (xb < scan_end) {
					const uint8_t* p = P.src + xb;
					if (xb + 16 <= P.src_size && (((uintptr_t)p) & 3u) == 0) {
#pragma unroll
						for (int j = 0; j < 4; ++j)
							w[j] = reinterpret_cast<const uint32_t*>(p)[j];
					}
					else {
#pragma unroll
						for (int j = 0; j < 4; ++j) {
							uint32_t v = 0;
#pragma unroll
							for (int b = 0; b < 4; ++b)
								v |= (xb + 4 * j + b < P.src_size ? (uint32_t)p[4 * j + b] : 0u) << (8 * b);
							w[j] = v;
						}
					}
				}
				// A header's csize is at most max_csize, so its top byte (3 bytes after the code) is small: a second SWAR
				// filter on the bytes already loaded (the next lane holds the 3 bytes past mine) removes nearly all of the
				// code-byte look-alikes before the chain check, which costs dependent loads.
				w[4] = __shfl_down_sync(FULL, w[0], 1);
				if (lane == 31)
					w[4] = 0u; // unknown: lets the last three positions pass
				if (xb < scan_end) {
					const uint32_t hi_max = P.max_csize >> 16;
					const uint32_t kadd = hi_max < 0x7Fu ? (0x7Fu - hi_max) * 0x01010101u : 0u;
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						uint32_t z = zero_bytes(w[j] ^ 0x01010101u) | zero_bytes(w[j] ^ 0x06060606u) | zero_bytes(w[j] ^ 0x02020202u);
						if (hi_max < 0x7Fu) {
							const uint32_t t3 = __funnelshift_r(w[j], w[j + 1], 24); // byte b + 3 at position b
							z &= ~(((t3 & 0x7F7F7F7Fu) + kadd) | t3);
						}
						cand |= flags_to_mask4(z) << (4 * j);
					}
				}
				while (cand) {
					const uint32_t b = (uint32_t)__ffs((int)cand) - 1u;
					cand &= cand - 1u;
					const uint64_t x = xb + b;
					if (x >= scan_end)
						break;
					if (index_chain_plausible(P.src, P.src_size, x, P.max_csize)) {
						found = x;
						break;
					}
				}
				const uint32_t m = __ballot_sync(FULL, found != IDX_NONE);
				if (m) {
					start = __shfl_sync(FULL, found, __ffs((int)m) - 1);
					break;
				}
			}
		}
		// count the headers of my own segment
		if (lane == 0) {
			unsigned long long pos = start;
			uint32_t cnt = 0;
			while (pos != IDX_NONE && pos < hi) {
				++cnt;
				pos = index_next(P.src, P.src_size, pos);
				if (pos == IDX_BROKEN)
					break;
			}
			P.seg_start[k] = start;
			P.seg_end[k] = pos;
			P.seg_count[k] = cnt;
		}
	}

	__global__ void __launch_bounds__(1024) index_merge_kernel(FastIndexParams P)
	{
		STENOS_DYN_SMEM(uint32_t, sm); // [0] ok flag, [2..2+32) warp sums
		const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
		if (tid == 0)
			sm[0] = 1u;
		__syncthreads();
		// ---- consistency of the hand-offs.  My walk must land exactly on the next segment's start; once the
		// chain has reached the end of the frame the remaining segments must have found nothing.  A segment
		// that re-synchronised on an imitation (structured payloads do contain plausible 6-hop chains) is
		// REPAIRED: it is walked again from where its predecessor landed.  Segment 0 is always right, so each
		// round fixes at least the first wrong segment; the rounds are bounded, then the serial walk takes over.
		for (int round = 0; round < 64; ++round) {
			if (tid == 0)
				sm[1] = 0u;
			__syncthreads();
			for (uint32_t k = tid; k + 1 < P.n_seg; k += blockDim.x) {
				const unsigned long long e = P.seg_end[k];
				const unsigned long long nx = P.seg_start[k + 1];
				const unsigned long long want = (e == P.src_size || e == IDX_NONE) ? IDX_NONE : e;
				const bool bad = e != IDX_BROKEN && nx != want;
				P.seg_base[k + 1] = bad ? 1u : 0u;
				if (bad)
					sm[1] = 1u;
			}
			__syncthreads();
			if (!sm[1])
				break;
			for (uint32_t k = tid; k + 1 < P.n_seg; k += blockDim.x) {
				if (!P.seg_base[k + 1])
					continue;
				const unsigned long long e = P.seg_end[k];
				const uint64_t hi = min(P.first + (uint64_t)(k + 2) * P.seg_bytes, P.src_size);
				unsigned long long pos = (e == P.src_size || e == IDX_NONE) ? IDX_NONE : e;
				P.seg_start[k + 1] = pos;
				uint32_t cnt = 0;
				while (pos != IDX_NONE && pos < hi) {
					++cnt;
					pos = index_next(P.src, P.src_size, pos);
					if (pos == IDX_BROKEN)
						break;
				}
				P.seg_end[k + 1] = pos;
				P.seg_count[k + 1] = cnt;
			}
			__syncthreads();
		}
		bool ok = sm[1] == 0u;
		for (uint32_t k = tid; k < P.n_seg; k += blockDim.x)
			if (P.seg_end[k] == IDX_BROKEN)
				ok = false;
		if (!ok)
			sm[0] = 0u;
		__syncthreads();
		// ---- exclusive scan of the counts, 1024 segments per pass with a running base
		uint32_t base = 0;
		if (sm[0]) {
			for (uint32_t k0 = 0; k0 < P.n_seg; k0 += blockDim.x) {
				const uint32_t k = k0 + tid;
				const uint32_t c = k < P.n_seg ? P.seg_count[k] : 0u;
				uint32_t incl = c;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					uint32_t t = __shfl_up_sync(FULL, incl, d);
					if (lane >= d)
						incl += t;
				}
				if (lane == 31)
					sm[2 + warp] = incl;
				__syncthreads();
				uint32_t wpre = 0, tot = 0;
				for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
					const uint32_t ws = sm[2 + i];
					if (i < warp)
						wpre += ws;
					tot += ws;
				}
				if (k < P.n_seg)
					P.seg_base[k] = min(base + wpre + incl - c, P.n_sb); // clamped: a wrong total is rejected below
				base += tot;
				__syncthreads();
			}
			if (tid == 0) {
				unsigned long long last_end = P.seg_end[P.n_seg - 1];
				if (last_end == IDX_NONE)
					last_end = P.src_size;
				if (base != P.n_sb || last_end > P.src_size)
					sm[0] = 0u;
				else
					P.sb_offsets[P.n_sb] = last_end;
			}
		}
		__syncthreads();
		if (tid == 0)
			*P.ok = sm[0];
		if (!sm[0] && tid == 0) {
			// fallback: the serial walk (stenos.cpp:1124-1143)
			uint64_t at = P.first;
			for (uint32_t i = 0; i < P.n_sb; ++i) {
				P.sb_offsets[i] = at;
				if (at + 4 > P.src_size) {
					atomicOr(&P.result[1], (unsigned long long)DEV_ERR_SRC_OVERFLOW);
					for (uint32_t j = i; j <= P.n_sb; ++j)
						P.sb_offsets[j] = P.src_size;
					return;
				}
				at += 4ull + rd24(P.src + at + 1);
			}
			P.sb_offsets[P.n_sb] = at;
			if (at > P.src_size)
				atomicOr(&P.result[1], (unsigned long long)DEV_ERR_INVALID_INPUT);
		}
	}

	// accepted: every segment walks its headers again (L2 hits) and writes them at its base
	__global__ void __launch_bounds__(128) index_fill_kernel(FastIndexParams P)
	{
		const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
		if (k >= P.n_seg || *P.ok == 0u)
			return;
		const uint32_t cnt = P.seg_count[k];
		uint32_t at_i = P.seg_base[k];
		unsigned long long pos = P.seg_start[k];
		for (uint32_t i = 0; i < cnt && at_i < P.n_sb; ++i, ++at_i) {
			P.sb_offsets[at_i] = pos;
			pos = index_next(P.src, P.src_size, pos);
		}
	}

	// ------------------------------------------------------------------------------------------
	// decoder: one warp per superblock
	// ------------------------------------------------------------------------------------------
	struct DecodeParams
	{
		const uint8_t* src;
		uint64_t src_size;
		uint8_t* dst;        // 16-byte aligned
		uint64_t total;      // decompressed bytes of the frame
		uint32_t sb_bytes;
		uint32_t n_sb;
		uint32_t first_sb;   // superblocks [first_sb, first_sb + n_sb) are decoded by this launch
		const unsigned long long* sb_offsets; // [>= first_sb + n_sb + 1] offsets of superblock headers in src
		unsigned long long* result;           // [1] error bits
		uint32_t skip_zstd_tail;              // 1: a code-2 (Zstd) final superblock is left to the host
		uint64_t dst_origin; // decompressed offset that dst[0] corresponds to (multi-GPU segments)
		unsigned long long* ticket; // decode_pairs_kernel: zero initialised work counter
	};

	constexpr int DECODE_WARPS = 4;

	// One warp decodes one superblock [code][csize:3][payload] found at offset `at` of src into
	// dsize bytes at `out` (16-byte aligned).  Returns 0 or device error bits.
	template<int T>
	__device__ __forceinline__ uint32_t decode_superblock_warp(const uint8_t* src, uint64_t src_size, uint64_t at, uint32_t dsize, uint8_t* out, uint16_t* lz_scratch,
								   int lane, bool allow_zstd_tail)
	{
		const uint8_t* lim = src + src_size;
		if (at + 4 > src_size)
			return DEV_ERR_SRC_OVERFLOW; // stenos.cpp:1126-1127
		const uint8_t* p = src + at;
		const uint32_t code = p[0], csize = rd24(p + 1);
		const uint8_t* q = p + 4;
		const uint8_t* end = q + csize;
		if (at + 4 + csize > src_size)
			return DEV_ERR_INVALID_INPUT; // stenos.cpp:1133-1134
		if (code == (uint32_t)CODE_COPY) {
			if (csize != dsize)
				return DEV_ERR_INVALID_INPUT; // stenos.cpp:743-744
			// aligned destination, arbitrary source
			const uint32_t nw = dsize >> 2;
			uint32_t* d = reinterpret_cast<uint32_t*>(out);
			const uintptr_t sa = (uintptr_t)q;
			const uint32_t sh = (uint32_t)(sa & 3u) * 8u;
			const uint32_t* sw = reinterpret_cast<const uint32_t*>(sa & ~(uintptr_t)3);
			for (uint32_t j = lane; j < nw; j += 32)
				d[j] = sh ? __funnelshift_r(sw[j], sw[j + 1], sh) : sw[j];
			for (uint32_t j = (nw << 2) + lane; j < dsize; j += 32)
				out[j] = q[j];
			return 0;
		}
		if (code == (uint32_t)CODE_BLOCK) {
			constexpr uint32_t BLOCK = T * 256u;
			const uint32_t nfull = dsize / BLOCK, rem = dsize - nfull * BLOCK;
			if (csize == 0 && dsize)
				return DEV_ERR_INVALID_INPUT;
			for (uint32_t b = 0; b < nfull; ++b) {
				const uint32_t r = decode_block<T>(q, end, lim, out + (size_t)b * BLOCK, lz_scratch, lane);
				if (r == 0xFFFFFFFFu)
					return DEV_ERR_INVALID_INPUT;
				q += r;
			}
			if (rem) {
				if (q >= end || *q != (uint8_t)MARK_PARTIAL) // block_compress.h:2160-2166
					return DEV_ERR_INVALID_INPUT;
				const uint32_t r = decode_partial_block<T>(q + 1, end, lim, out + (size_t)nfull * BLOCK, rem, lane);
				if (r == 0xFFFFFFFFu)
					return DEV_ERR_INVALID_INPUT;
			}
			return 0;
		}
		if (code == (uint32_t)CODE_ZSTD && allow_zstd_tail && dsize < 128u)
			return 0; // tiny final superblock (stenos.cpp:435-437): decoded by the host layer through libzstd
		return DEV_ERR_INVALID_INPUT; // codes 3,4,5 carry Zstd payloads (levels >= 2): out of scope
	}

	template<int T>
	__global__ void __launch_bounds__(DECODE_WARPS * 32) decode_frame_kernel(DecodeParams P)
	{
		STENOS_DYN_SMEM(uint8_t, smem);
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		uint16_t* lz_scratch = reinterpret_cast<uint16_t*>(smem) + 256 * warp;
		const uint32_t i = blockIdx.x * DECODE_WARPS + warp;
		if (i >= P.n_sb)
			return;
		const uint32_t s = P.first_sb + i;
		const uint64_t doff = (uint64_t)s * P.sb_bytes;
		const uint32_t dsize = (uint32_t)min((uint64_t)P.sb_bytes, P.total - doff); // remainder 0 = full superblock (appendix C1)
		const bool last = (doff + dsize == P.total);
		const uint32_t bad = decode_superblock_warp<T>(P.src, P.src_size, P.sb_offsets[s], dsize, P.dst + (doff - P.dst_origin), lz_scratch, lane, P.skip_zstd_tail && last);
		if (bad && lane == 0)
			atomicOr(&P.result[1], (unsigned long long)bad);
	}

	// stenos::cvector random access: one warp per requested bucket (cvector.hpp:2879 -> :1862-1883)
	struct GatherParams
	{
		const uint8_t* src;
		uint64_t src_size;
		uint8_t* dst;          // [n x bucket_bytes]
		uint64_t total;        // decompressed bytes of the whole container
		uint32_t bucket_bytes; // superblock size of the frame
		uint32_t n_buckets;
		const unsigned long long* sb_offsets;
		const uint32_t* ids;
		uint32_t n;
		unsigned long long* result;
	};

	template<int T>
	__global__ void __launch_bounds__(DECODE_WARPS * 32) gather_decode_kernel(GatherParams P)
	{
		STENOS_DYN_SMEM(uint8_t, smem);
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		uint16_t* lz_scratch = reinterpret_cast<uint16_t*>(smem) + 256 * warp;
		const uint32_t i = blockIdx.x * DECODE_WARPS + warp;
		if (i >= P.n)
			return;
		const uint32_t id = P.ids[i];
		uint32_t bad;
		if (id >= P.n_buckets)
			bad = DEV_ERR_INVALID_INPUT;
		else {
			const uint64_t doff = (uint64_t)id * P.bucket_bytes;
			const uint32_t dsize = (uint32_t)min((uint64_t)P.bucket_bytes, P.total - doff);
			bad = decode_superblock_warp<T>(P.src, P.src_size, P.sb_offsets[id], dsize, P.dst + (uint64_t)i * P.bucket_bytes, lz_scratch, lane, false);
		}
		if (bad && lane == 0)
			atomicOr(&P.result[1], (unsigned long long)bad);
	}
}
