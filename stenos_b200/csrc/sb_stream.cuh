// sb_stream.cuh -- encode_stream_kernel: the level-1 frame encoder as a barrier-free warp pipeline.
//
// Same job as encode_frame_kernel (stenos_compress_generic's superblock loop, stenos.cpp:893-904, and
// compress_generic_superblock's level-1 branch, :403-450, :606-615, :363-374) for the superblocks
// whose dst-room checks are provably inert; the few that are not are left to
// encode_frame_kernel, which shares the look-back words.
//
// Persistent CTAs; a CTA's q-th superblock is blockIdx.x for q = 0 and then whatever the global
// ticket counter hands out (fetched one superblock ahead), so CTAs that meet cheap superblocks simply
// take more of them.  Inside a CTA nothing ever waits for the whole CTA:
//
//   * warps take TICKETS from a shared-memory counter; ticket t = (q, j) is pair j of two
//     256-element blocks of the CTA's q-th superblock (the slot after the last pair is the partial
//     tail block, if any);
//   * a pair is analysed in registers (sb_encode_rows.cuh), its size published in a per-superblock
//     chain in shared memory, and its offset in the superblock's block stream obtained by a
//     decoupled look-back over that chain (32 predecessors per probe);
//   * the bytes are emitted straight at their final position of the COMPACT stream, which lives in
//     a shared-memory ring shared by consecutive superblocks (a block never wraps: the ring has a
//     linear tail, and the flush knows where a superblock's stream jumps back to the ring start);
//   * the warp that completes a superblock's last ticket flushes it: decoupled look-back across
//     CTAs for the frame offset, [code][csize:3], one copy of the stream with 16-byte stores (or of
//     the raw input when the stream is larger: COPY, stenos.cpp:609-610), and releases the ring space.
//
// HBM traffic: N read + C written, nothing else.
#pragma once
#include "sb_kernels.cuh"

namespace sb
{
	constexpr uint32_t CH_AGG = 1u << 30;
	constexpr uint32_t CH_INC = 2u << 30;
	constexpr uint32_t CH_VALUE = (1u << 30) - 1u;
	constexpr unsigned long long START_READY = 1ull << 63;

	struct StreamSlot
	{
		unsigned long long start; // START_READY | ring position (absolute, monotonic) of the superblock's stream
		uint32_t done;            // tickets completed
		uint32_t cross_end;       // absolute end of the block that ran into the ring's linear tail (0: none)
		uint32_t sb_plus1;        // 1 + the superblock this slot works on (0: not assigned yet)
		uint32_t pad[3];
	};
	struct StreamCtl
	{
		uint32_t ticket;
		uint32_t flushed_q;   // superblocks [0, flushed_q) of this CTA are flushed, their slots reusable
		uint32_t flushed_abs; // ring bytes below this absolute position are free
		uint32_t finished_q;  // first q of this CTA whose superblock is past the end (0xFFFFFFFF: not met yet)
	};

	template<int T, int NT>
	struct StreamLayout
	{
		static constexpr uint32_t BLOCK = T * 256u;
		static constexpr uint32_t HS = (T + 1) / 2;
		static constexpr uint32_t NWARPS = NT / 32;
		static constexpr uint32_t TMP = (BLOCK + HS + 8u * T + 1u + 15u) & ~15u; // worst block / partial block, 16-byte aligned
		static constexpr uint32_t NSLOT = 8;
		static constexpr uint32_t PMAX = DEFAULT_SUPERBLOCK / BLOCK / 2 + 1; // tickets per superblock: pairs + the partial block
		static constexpr uint32_t HAS_LZ = (T % 4) == 0 ? 1u : 0u;
		static constexpr uint32_t LZ_STRIDE = (LZ_SCRATCH_BYTES + 15u) & ~15u;
		static constexpr uint32_t CTL_OFF = 0;
		static constexpr uint32_t SLOT_OFF = 16;
		static constexpr uint32_t CHAIN_OFF = SLOT_OFF + NSLOT * 32u;
		static constexpr uint32_t TMP_OFF = (CHAIN_OFF + NSLOT * PMAX * 4u + 15u) & ~15u;
		static constexpr uint32_t LZ_OFF = TMP_OFF + NWARPS * 2u * TMP;
		static constexpr uint32_t RING_OFF = LZ_OFF + HAS_LZ * NWARPS * LZ_STRIDE;
		static constexpr uint32_t SMEM_TOTAL = 227u * 1024u;
		static constexpr uint32_t TAIL = TMP + 32u; // linear tail of the ring (+ over-read slack of the flush copy)
		static constexpr uint32_t RING = (SMEM_TOTAL - RING_OFF - TAIL) & ~15u;
		// a superblock's stream allocation never exceeds (sb / BLOCK) * (BLOCK + HS) + partial
		static constexpr uint32_t MAX_ALLOC = (DEFAULT_SUPERBLOCK / BLOCK) * (BLOCK + HS) + TMP;
		static_assert(RING >= MAX_ALLOC + TAIL + 4096u, "ring too small");
		static constexpr uint32_t smem_bytes() { return RING_OFF + RING + TAIL; }
	};

	// decoupled look-back over one superblock's chain (shared memory): exclusive prefix of ticket j
	__device__ __forceinline__ uint32_t chain_lookback(const uint32_t* chain, uint32_t j, int lane)
	{
		const volatile uint32_t* ch = chain;
		uint32_t excl = 0;
		int top = (int)j - 1;
		while (top >= 0) {
			const int idx = top - lane;
			const uint32_t v = idx >= 0 ? ch[idx] : CH_INC; // before the first ticket: inclusive prefix 0
			const uint32_t inc = __ballot_sync(FULL, (v & CH_INC) != 0u);
			const uint32_t inv = __ballot_sync(FULL, (v >> 30) == 0u);
			const int fi = inc ? (__ffs((int)inc) - 1) : 32;
			const uint32_t upto = fi >= 31 ? 0xFFFFFFFFu : ((2u << fi) - 1u);
			if (inv & upto) {
				STENOS_SPIN_HINT();
				continue;
			}
			const uint32_t val = ((upto >> lane) & 1u) ? (v & CH_VALUE) : 0u;
			excl += __reduce_add_sync(FULL, val);
			if (fi < 32)
				break;
			top -= 32;
		}
		return excl;
	}

	// one warp copies n bytes src (global, 16-byte aligned) -> dst (global, any alignment), several loads in flight
	__device__ __forceinline__ void warp_copy_global(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, int lane)
	{
		const uint32_t head = min((uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u), n);
		if ((uint32_t)lane < head)
			dst[lane] = src[lane];
		const uint32_t body = (n - head) >> 4;
		uint4* d = reinterpret_cast<uint4*>(dst + head);
		const uint32_t sh = (head & 3u) * 8u;
		const uint32_t* sw = reinterpret_cast<const uint32_t*>(src + (head & ~3u));
		if (head == 0) {
			const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll 8
			for (uint32_t j = lane; j < body; j += 32)
				d[j] = s4[j];
		}
		else {
#pragma unroll 4
			for (uint32_t j = lane; j < body; j += 32) {
				const uint32_t a = sw[4 * j], b = sw[4 * j + 1], c = sw[4 * j + 2], e = sw[4 * j + 3];
				const uint32_t f = sh ? sw[4 * j + 4] : 0u;
				d[j] = make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh), __funnelshift_r(c, e, sh), __funnelshift_r(e, f, sh));
			}
		}
		const uint32_t done = head + (body << 4);
		if (done + (uint32_t)lane < n)
			dst[done + lane] = src[done + lane];
	}

	template<int T, int NT>
	__global__ void __launch_bounds__(NT, 1) encode_stream_kernel(EncodeParams P)
	{
		using L = StreamLayout<T, NT>;
		STENOS_DYN_SMEM(uint8_t, smem);
		StreamCtl* ctl = reinterpret_cast<StreamCtl*>(smem + L::CTL_OFF);
		StreamSlot* slots = reinterpret_cast<StreamSlot*>(smem + L::SLOT_OFF);
		uint32_t* chains = reinterpret_cast<uint32_t*>(smem + L::CHAIN_OFF);
		uint8_t* ring = smem + L::RING_OFF;
		const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
		uint8_t* tmp = smem + L::TMP_OFF + (uint32_t)warp * 2u * L::TMP;
		uint32_t* lz_scratch = reinterpret_cast<uint32_t*>(smem + L::LZ_OFF + L::LZ_STRIDE * (L::HAS_LZ ? warp : 0));
		const uint64_t first_off = P.header_len ? (uint64_t)P.header_len : P.base_offset;
		volatile uint32_t* v_flushed_q = &ctl->flushed_q;
		volatile uint32_t* v_flushed_abs = &ctl->flushed_abs;

		for (uint32_t i = tid; i < (L::TMP_OFF - L::CTL_OFF) / 4u; i += blockDim.x)
			reinterpret_cast<uint32_t*>(smem)[i] = 0u;
		__syncthreads();
		if (tid == 0) {
			ctl->finished_q = 0xFFFFFFFFu;
			slots[0].start = START_READY;
			slots[0].sb_plus1 = blockIdx.x + 1u;
		}
		__syncthreads();

		// tickets per superblock: the pairs of a full superblock + one for a partial tail block
		const uint32_t nfull_sb = P.sb_bytes / L::BLOCK;
		const uint32_t PT = (nfull_sb + 1u) / 2u + 1u;
		const uint32_t PT_RCP = (uint32_t)((0x100000000ull + PT - 1u) / PT); // t / PT == umulhi(t, PT_RCP) for t < 2^32 / PT

		for (;;) {
			uint32_t t = 0;
			if (lane == 0)
				t = atomicAdd(&ctl->ticket, 1u);
			t = __shfl_sync(FULL, t, 0);
			const uint32_t q = __umulhi(t, PT_RCP), j = t - q * PT;
			// the slot of q and the one of q + 1 (whose superblock and start are published from q) must be free
			bool finished = false;
			while ((int)(q - *v_flushed_q) >= (int)L::NSLOT - 1) {
				if (q >= *reinterpret_cast<volatile uint32_t*>(&ctl->finished_q)) { // past this CTA's last superblock: the slot will never come
					finished = true;
					break;
				}
				STENOS_SPIN_WAIT();
			}
			if (finished)
				break;
			StreamSlot* slot = &slots[q % L::NSLOT];
			uint32_t* chain = chains + (q % L::NSLOT) * L::PMAX;
			uint32_t s;
			while ((s = *reinterpret_cast<volatile uint32_t*>(&slot->sb_plus1)) == 0u) {
				if (q >= *reinterpret_cast<volatile uint32_t*>(&ctl->finished_q)) { // only tickets past the end may give up
					s = 0xFFFFFFFFu;
					break;
				}
				STENOS_SPIN_HINT();
			}
			s -= 1u;
			if (s >= P.n_stream) {
				// superblocks are handed out in increasing order: every ticket of this CTA from q on is past the end too
				// (earlier ones are not, and must never give up)
				if (lane == 0)
					atomicMin(&ctl->finished_q, q);
				break;
			}
			if (j == 0 && lane == 0) {
				// the next superblock of this CTA, one superblock ahead of its first use.  Asked for only once this
				// one is known, so that a CTA's superblocks are numbered in increasing order of q.
				const uint32_t nx = gridDim.x + atomicAdd(P.ticket + 1, 1u); // [1]: [0] belongs to encode_frame_kernel
				*reinterpret_cast<volatile uint32_t*>(&slots[(q + 1u) % L::NSLOT].sb_plus1) = nx + 1u;
			}

			const uint8_t* in = P.src + (uint64_t)s * P.sb_bytes;
			const uint32_t in_bytes = (uint32_t)min((uint64_t)P.sb_bytes, P.bytes - (uint64_t)s * P.sb_bytes);
			const uint32_t nfull = in_bytes / L::BLOCK, rem = in_bytes - nfull * L::BLOCK;
			const uint32_t npairs = (nfull + 1u) >> 1;

#ifndef STREAM_NO_PREFETCH
			// The pair NWARPS tickets ahead (roughly this warp's next one) starts its way from HBM to L2 now: its
			// loads then meet L2 latency instead of DRAM latency.  A wrong guess costs nothing but the request.
			{
				uint32_t jj = j + L::NWARPS;
				const uint8_t* pf = nullptr;
				if (jj < npairs)
					pf = in + (size_t)(2u * jj) * L::BLOCK;
				else {
					jj -= PT;
					const uint32_t s1 = *reinterpret_cast<volatile uint32_t*>(&slots[(q + 1u) % L::NSLOT].sb_plus1);
					if (s1 != 0u && s1 - 1u < P.n_stream && jj < PT)
						pf = P.src + (uint64_t)(s1 - 1u) * P.sb_bytes + (size_t)(2u * jj) * L::BLOCK;
				}
				if (pf != nullptr && lane < 4 * T) {
					const uint8_t* a = pf + 128u * (uint32_t)lane;
					if (a < P.src + P.bytes)
						prefetch_l2(a);
				}
			}
#endif

			// sizes are final -> stream offset (look-back), ring position, room in the ring
			uint32_t abs0 = 0, excl = 0;
			auto place_pair = [&](uint32_t sz2) {
				// (a serial chain -- wait for the predecessor's inclusive prefix only -- was measured at 1.45 ms against
				// 0.75 ms: the links' store -> load round trips add up across the warps in flight)
				if (lane == 0)
					*reinterpret_cast<volatile uint32_t*>(&chain[j]) = (j == 0 ? CH_INC : CH_AGG) | sz2;
				excl = chain_lookback(chain, j, lane);
				if (lane == 0 && j != 0)
					*reinterpret_cast<volatile uint32_t*>(&chain[j]) = CH_INC | (excl + sz2);
				unsigned long long st;
				while (((st = *reinterpret_cast<volatile unsigned long long*>(&slot->start)) & START_READY) == 0ull)
					STENOS_SPIN_HINT();
				abs0 = (uint32_t)st + excl;
				if (j == PT - 1u && lane == 0)
					*reinterpret_cast<volatile unsigned long long*>(&slots[(q + 1u) % L::NSLOT].start) = START_READY | (unsigned long long)(abs0 + sz2);
				while ((int)(abs0 + sz2 - *v_flushed_abs) > (int)(L::RING - L::TAIL))
					STENOS_SPIN_HINT();
			};
			// where a block of `sz` bytes at stream offset `off` goes (nullptr: the superblock is COPY anyway)
			auto block_ptr = [&](uint32_t off, uint32_t sz) -> uint8_t* {
				if (off + sz > in_bytes)
					return nullptr;
				const uint32_t a = abs0 + (off - excl);
				const uint32_t p = a % L::RING;
				if (p + sz > L::RING && (lane & 15) == 0)
					slot->cross_end = a + sz;
				return ring + p;
			};

			if (j < npairs) {
				const uint32_t b = 2u * j;
				const bool second = b + 1u < nfull;
				encode_block_pair<T>(in + (size_t)b * L::BLOCK, second, tmp, L::TMP, lz_scratch, lane, [&](uint32_t size) -> uint8_t* {
					const uint32_t szA = __shfl_sync(FULL, size, 0);
					const uint32_t szB = second ? __shfl_sync(FULL, size, 16) : 0u;
					place_pair(szA + szB);
					uint8_t* pa = block_ptr(excl, szA);
					uint8_t* pb = second ? block_ptr(excl + szA, szB) : pa;
					return (lane >> 4) ? pb : pa;
				});
			}
			else if (j == npairs && rem) {
				// partial tail block: built in the temporary (its size is only known once it exists), then moved
				bool e = false;
				const uint32_t sz = encode_partial_block<T, false>(in + (size_t)nfull * L::BLOCK, rem, tmp, lane, 0xFFFFFFFFu, e);
				__syncwarp();
				place_pair(sz);
				uint8_t* to = block_ptr(excl, sz);
				if (to)
					for (uint32_t i = lane; i < sz; i += 32)
						to[i] = tmp[i];
			}
			else
				place_pair(0u);

			// ---- ticket done; the last one flushes the superblock
			__syncwarp();
			__threadfence_block();
			uint32_t nd = 0;
			if (lane == 0)
				nd = atomicAdd(&slot->done, 1u);
			nd = __shfl_sync(FULL, nd, 0);
			if (nd != PT - 1u)
				continue;
			__threadfence_block();

			const uint32_t start_abs = (uint32_t)*reinterpret_cast<volatile unsigned long long*>(&slot->start);
			const uint32_t csize = *reinterpret_cast<volatile uint32_t*>(&chain[PT - 1u]) & CH_VALUE;
			const uint32_t cross_end = *reinterpret_cast<volatile uint32_t*>(&slot->cross_end);
			const bool copy = csize > in_bytes; // stenos.cpp:609-610
			const uint32_t len = copy ? in_bytes : csize;
			const uint32_t out_size = 4u + len;

			// frame offset: decoupled look-back across superblocks (other CTAs), 32 predecessors per probe
			uint64_t fexcl = 0;
			{
				if (lane == 0)
					st_volatile_u64(&P.state[s], LB_AGGREGATE | (unsigned long long)out_size);
				long long top = (long long)s - 1;
				while (top >= 0) {
					const long long idx = top - lane;
					const unsigned long long v = idx >= 0 ? ld_volatile_u64(&P.state[idx]) : LB_INCLUSIVE;
					const uint32_t inc = __ballot_sync(FULL, (v & LB_INCLUSIVE) != 0ull);
					const uint32_t inv = __ballot_sync(FULL, (v >> 62) == 0ull);
					const int fi = inc ? (__ffs((int)inc) - 1) : 32;
					const uint32_t upto = fi >= 31 ? 0xFFFFFFFFu : ((2u << fi) - 1u);
					if (inv & upto) {
						STENOS_SPIN_HINT();
						continue;
					}
					unsigned long long val = ((upto >> lane) & 1u) ? (v & LB_VALUE) : 0ull;
#pragma unroll
					for (int d = 16; d >= 1; d >>= 1)
						val += __shfl_xor_sync(FULL, val, d);
					fexcl += val;
					if (fi < 32)
						break;
					top -= 32;
				}
			}
			const uint64_t base = first_off + fexcl;
			if (lane == 0) {
				st_volatile_u64(&P.state[s], LB_INCLUSIVE | (unsigned long long)(fexcl + out_size));
				if (P.sb_offsets) {
					P.sb_offsets[s] = base;
					if (s == P.n_sb - 1)
						P.sb_offsets[s + 1] = base + out_size;
				}
				if (s == P.n_sb - 1)
					P.result[0] = base + out_size;
			}
			if (base + out_size > P.dst_size) {
				// stenos.cpp:366-367 / :611-612: the caller's buffer is too small
				if (lane == 0)
					atomicOr(&P.result[1], (unsigned long long)DEV_ERR_DST_OVERFLOW);
			}
			else {
				uint8_t* out = P.dst + base;
				if (lane == 0) {
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
					warp_copy_global(out + 4, in, in_bytes, lane);
				else {
					// the stream is linear in the ring from start_abs up to the block that ran into the tail (or
					// up to the ring's end), then continues near the ring's start
					const uint32_t p0 = start_abs % L::RING;
					uint32_t first_len = csize;
					if (cross_end != 0u && (int)(cross_end - start_abs) > 0 && cross_end - start_abs <= csize)
						first_len = cross_end - start_abs;
					else if (p0 + csize > L::RING)
						first_len = L::RING - p0;
					cta_copy_bytes(out + 4, ring + p0, first_len, lane, 32);
					if (first_len < csize)
						cta_copy_bytes(out + 4 + first_len, ring + (start_abs + first_len) % L::RING, csize - first_len, lane, 32);
				}
			}
			// release, in superblock order: chain + slot of q, ring space up to the start of q + 1
			const uint32_t next_abs = (uint32_t)*reinterpret_cast<volatile unsigned long long*>(&slots[(q + 1u) % L::NSLOT].start);
			while (*v_flushed_q != q)
				STENOS_SPIN_HINT();
			__syncwarp();
			for (uint32_t i = lane; i < PT; i += 32)
				chain[i] = 0u;
			if (lane == 0) {
				slot->start = 0ull;
				slot->done = 0u;
				slot->cross_end = 0u;
				slot->sb_plus1 = 0u;
			}
			__syncwarp();
			__threadfence_block();
			if (lane == 0) {
				*v_flushed_abs = next_abs;
				__threadfence_block();
				*v_flushed_q = q + 1u;
			}
			__syncwarp();
		}
	}
}
