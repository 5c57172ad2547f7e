// sb_filters.cuh -- byte shuffle (transpose) and byte delta filters, the pre-Zstd stages of
// levels >= 2 (stenos.cpp:513, :646, :709, :722-724).
//
//   shuffle      stenos::shuffle   (shuffle.cpp:82-90, semantics shuffle-generic.h:33-74)
//   unshuffle    stenos::unshuffle (shuffle.cpp:94-102, shuffle-generic.h:83-125)
//   delta        stenos::delta     (delta.cpp:30-71: 4 independent quarter streams when bytes > 2048)
//   delta_inv    stenos::delta_inv (delta.cpp:230-267)
//
// All kernels are batched: the buffer is cut in `chunk` byte pieces (the superblock size of the
// level: 128 KiB << shift) and every piece is filtered on its own, exactly as the reference does
// per superblock.  Lane mapping of the transposes: a lane owns 16 consecutive elements, so every
// byte plane receives / provides one 16-byte vector per lane and a warp moves 512 contiguous
// elements with 128-bit accesses on both sides.
#pragma once
#include "sb_common.cuh"
#include "sb_decode_rows.cuh" // sad4_acc, prefix16
#include "sb_flow.cuh"        // cp_async16, lds_u128, smem_addr32

namespace sb
{
	struct FilterParams
	{
		const uint8_t* src;
		uint8_t* dst;
		uint64_t bytes; // total
		uint64_t chunk; // bytes per independently filtered piece (last one may be shorter)
		uint32_t with_delta; // shuffle: also apply the byte delta (fused); unshuffle: unused
		uint32_t chunk_in_y; // grid layout of the transposes: 0: x = chunk, y = groups of a chunk; 1: the other way round
	};

	// position i of a chunk of `bytes` bytes starts a delta stream? (delta.cpp:42-70)
	__device__ __forceinline__ bool delta_stream_start(uint64_t i, uint64_t bytes)
	{
		if (i == 0)
			return true;
		if (bytes <= 2048)
			return false;
		const uint64_t q = bytes / 4;
		return i == q || i == 2 * q || i == 3 * q;
	}

	// 16 elements of T bytes (4*T words) -> T vectors of 16 plane bytes
	template<int T>
	__device__ __forceinline__ void transpose16(const uint32_t (&w)[4 * T], uint4 (&pl)[T])
	{
		if (T == 2) {
			uint32_t a[4], b[4];
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				a[i] = __byte_perm(w[2 * i], w[2 * i + 1], 0x6420);
				b[i] = __byte_perm(w[2 * i], w[2 * i + 1], 0x7531);
			}
			pl[0] = make_uint4(a[0], a[1], a[2], a[3]);
			pl[1 % T] = make_uint4(b[0], b[1], b[2], b[3]);
		}
		else if (T == 4) {
			uint32_t p[4][4];
#pragma unroll
			for (int i = 0; i < 4; ++i)
				transpose4(w[(4 * i) % (4 * T)], w[(4 * i + 1) % (4 * T)], w[(4 * i + 2) % (4 * T)], w[(4 * i + 3) % (4 * T)], p[0][i], p[1][i], p[2][i], p[3][i]);
#pragma unroll
			for (int k = 0; k < 4; ++k)
				pl[k % T] = make_uint4(p[k][0], p[k][1], p[k][2], p[k][3]);
		}
		else {
			uint32_t p[8][4];
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				transpose4(w[(8 * i) % (4 * T)], w[(8 * i + 2) % (4 * T)], w[(8 * i + 4) % (4 * T)], w[(8 * i + 6) % (4 * T)], p[0][i], p[1][i], p[2][i], p[3][i]);
				transpose4(w[(8 * i + 1) % (4 * T)], w[(8 * i + 3) % (4 * T)], w[(8 * i + 5) % (4 * T)], w[(8 * i + 7) % (4 * T)], p[4][i], p[5][i], p[6][i], p[7][i]);
			}
#pragma unroll
			for (int k = 0; k < 8; ++k)
				pl[k % T] = make_uint4(p[k][0], p[k][1], p[k][2], p[k][3]);
		}
	}
	template<int T>
	__device__ __forceinline__ void untranspose16(const uint4 (&pl)[T], uint32_t (&w)[4 * T])
	{
		if (T == 2) {
			const uint32_t a[4] = { pl[0].x, pl[0].y, pl[0].z, pl[0].w };
			const uint32_t b[4] = { pl[1 % T].x, pl[1 % T].y, pl[1 % T].z, pl[1 % T].w };
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				w[(2 * i) % (4 * T)] = __byte_perm(a[i], b[i], 0x5140);
				w[(2 * i + 1) % (4 * T)] = __byte_perm(a[i], b[i], 0x7362);
			}
		}
		else if (T == 4) {
			uint32_t p[4][4];
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				p[k][0] = pl[k % T].x;
				p[k][1] = pl[k % T].y;
				p[k][2] = pl[k % T].z;
				p[k][3] = pl[k % T].w;
			}
#pragma unroll
			for (int i = 0; i < 4; ++i)
				transpose4(p[0][i], p[1][i], p[2][i], p[3][i], w[(4 * i) % (4 * T)], w[(4 * i + 1) % (4 * T)], w[(4 * i + 2) % (4 * T)], w[(4 * i + 3) % (4 * T)]);
		}
		else {
			uint32_t p[8][4];
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				p[k][0] = pl[k % T].x;
				p[k][1] = pl[k % T].y;
				p[k][2] = pl[k % T].z;
				p[k][3] = pl[k % T].w;
			}
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				transpose4(p[0][i], p[1][i], p[2][i], p[3][i], w[(8 * i) % (4 * T)], w[(8 * i + 2) % (4 * T)], w[(8 * i + 4) % (4 * T)], w[(8 * i + 6) % (4 * T)]);
				transpose4(p[4][i], p[5][i], p[6][i], p[7][i], w[(8 * i + 1) % (4 * T)], w[(8 * i + 3) % (4 * T)], w[(8 * i + 5) % (4 * T)], w[(8 * i + 7) % (4 * T)]);
			}
		}
	}

	// a thread's 16 elements (16 * T contiguous bytes, 16-byte aligned): whole 32-byte sectors per instruction when aligned
	template<int T>
	__device__ __forceinline__ void store_elements16(uint8_t* d, const uint32_t (&w)[4 * T])
	{
		if ((reinterpret_cast<uintptr_t>(d) & 31u) == 0) {
#pragma unroll
			for (int i = 0; i < T; i += 2)
				st_global_256(d + 16 * i, make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]), make_uint4(w[4 * i + 4], w[4 * i + 5], w[4 * i + 6], w[4 * i + 7]));
		}
		else {
			uint4* d4 = reinterpret_cast<uint4*>(d);
#pragma unroll
			for (int i = 0; i < T; ++i)
				d4[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
		}
	}

	// byte delta of a 16-byte vector given the byte that precedes it
	__device__ __forceinline__ uint4 delta16(uint4 v, uint32_t before)
	{
		uint4 r;
		r.x = __vsub4(v.x, prev_bytes(before << 24, v.x));
		r.y = __vsub4(v.y, prev_bytes(v.x, v.y));
		r.z = __vsub4(v.z, prev_bytes(v.y, v.z));
		r.w = __vsub4(v.w, prev_bytes(v.z, v.w));
		return r;
	}
	__device__ __forceinline__ uint32_t byte16(const uint4& v, uint32_t idx)
	{
		const uint32_t a = idx < 4 ? v.x : idx < 8 ? v.y : idx < 12 ? v.z : v.w;
		return (a >> ((idx & 3u) * 8u)) & 0xFFu;
	}
	__device__ __forceinline__ void set_byte16(uint4& v, uint32_t idx, uint32_t b)
	{
		uint32_t* a = idx < 4 ? &v.x : idx < 8 ? &v.y : idx < 12 ? &v.z : &v.w;
		const uint32_t sh = (idx & 3u) * 8u;
		*a = (*a & ~(0xFFu << sh)) | (b << sh);
	}

	constexpr int FILTER_THREADS = 256;

	// groups of 16 elements a thread of shuffle_kernel handles, FILTER_THREADS groups apart: small elements make small CTAs
	// (T = 2: 8 KiB per CTA with one group per thread; the fused delta sat at 0.78 of HBM bandwidth)
	template<int T>
	struct ShuffleGroups
	{
		static constexpr int N = T == 2 ? 4 : (T == 4 ? 2 : 1);
	};

	template<int T>
	__device__ __forceinline__ void shuffle_group(const FilterParams& P, const uint8_t* src, uint8_t* dst, uint64_t cb, uint64_t n, uint64_t j0);

	// shuffle (+ optional fused delta).  grid.x covers groups of 16 elements of every chunk.
	template<int T>
	__global__ void __launch_bounds__(FILTER_THREADS) shuffle_kernel(FilterParams P)
	{
		// grid: x = chunk, (y * N + u) * FILTER_THREADS + thread = group of 16 elements inside the chunk (+ one group for the leftover bytes)
		const uint64_t c = P.chunk_in_y ? blockIdx.y : blockIdx.x;
		const uint64_t cb = min(P.chunk, P.bytes - c * P.chunk); // bytes of this chunk
		const uint64_t n = cb / T;                                // elements
		const uint8_t* src = P.src + c * P.chunk;
		uint8_t* dst = P.dst + c * P.chunk;
#pragma unroll 1
		for (int u = 0; u < ShuffleGroups<T>::N; ++u)
			shuffle_group<T>(P, src, dst, cb, n, (((uint64_t)(P.chunk_in_y ? blockIdx.x : blockIdx.y) * ShuffleGroups<T>::N + u) * FILTER_THREADS + threadIdx.x) * 16);
	}
	template<int T>
	__device__ __forceinline__ void shuffle_group(const FilterParams& P, const uint8_t* src, uint8_t* dst, uint64_t cb, uint64_t n, uint64_t j0)
	{
		if (j0 >= n) {
			// the thread after the last group copies the leftover bytes (shuffle-generic.h:73)
			if (j0 < n + 16) {
				for (uint64_t i = n * T; i < cb; ++i) {
					uint32_t b = src[i];
					if (P.with_delta && !delta_stream_start(i, cb))
						b = (b - src[i == n * T ? (n - 1) * T + (T - 1) : i - 1]) & 0xFFu;
					dst[i] = (uint8_t)b;
				}
			}
			return;
		}
		const bool vec = (j0 + 16 <= n) && ((n & 15u) == 0) && ((((uintptr_t)src) & 15u) == 0) && ((((uintptr_t)dst) & 15u) == 0);
		if (vec) {
			uint32_t w[4 * T];
			const uint4* s4 = reinterpret_cast<const uint4*>(src + j0 * T);
#pragma unroll
			for (int i = 0; i < T; ++i) {
				const uint4 v = s4[i];
				w[4 * i] = v.x;
				w[4 * i + 1] = v.y;
				w[4 * i + 2] = v.z;
				w[4 * i + 3] = v.w;
			}
			uint4 pl[T];
			transpose16<T>(w, pl);
			// fused delta (delta.cpp:30-71 on the transposed chunk): the byte before plane k's 16 bytes is byte k of the
			// element before j0 -- or, for the first group, byte k - 1 of the chunk's last element.  One element load
			// (the neighbour thread has just fetched the line), not T byte loads.
			uint32_t pe[2] = { 0u, 0u };
			uint64_t q1 = ~0ull, q2 = ~0ull, q3 = ~0ull; // starts of the quarter streams 1..3 (chunk offsets)
			if (P.with_delta) {
				const uint8_t* e = src + (j0 ? j0 - 1 : n - 1) * T;
				if (T == 2)
					pe[0] = *reinterpret_cast<const uint16_t*>(e);
				else if (T == 4)
					pe[0] = *reinterpret_cast<const uint32_t*>(e);
				else {
					const uint2 t = *reinterpret_cast<const uint2*>(e);
					pe[0] = t.x;
					pe[1] = t.y;
				}
				if (j0 == 0) { // byte k - 1 of the last element; nothing before plane 0
					pe[1] = T == 8 ? __funnelshift_l(pe[0], pe[1], 8) : 0u;
					pe[0] <<= 8;
				}
				if (cb > 2048) {
					q1 = cb / 4;
					q2 = 2u * q1;
					q3 = 3u * q1;
				}
			}
#pragma unroll
			for (int k = 0; k < T; ++k) {
				uint4 v = pl[k];
				if (P.with_delta) {
					const uint64_t i0 = (uint64_t)k * n + j0; // chunk offset of the plane's 16 bytes
					const uint32_t before = (pe[k >> 2] >> (8 * (k & 3))) & 0xFFu;
					const uint4 raw = v;
					v = delta16(v, before);
					// stream starts keep the raw byte
					if (i0 == 0)
						set_byte16(v, 0, raw.x & 0xFFu);
					if (q1 - i0 < 16ull)
						set_byte16(v, (uint32_t)(q1 - i0), byte16(raw, (uint32_t)(q1 - i0)));
					if (q2 - i0 < 16ull)
						set_byte16(v, (uint32_t)(q2 - i0), byte16(raw, (uint32_t)(q2 - i0)));
					if (q3 - i0 < 16ull)
						set_byte16(v, (uint32_t)(q3 - i0), byte16(raw, (uint32_t)(q3 - i0)));
				}
				*reinterpret_cast<uint4*>(dst + (uint64_t)k * n + j0) = v;
			}
		}
		else {
			// ragged / unaligned group: plain byte accesses
			const uint64_t j1 = min(j0 + 16, n);
			for (int k = 0; k < T; ++k)
				for (uint64_t j = j0; j < j1; ++j) {
					const uint64_t i = (uint64_t)k * n + j;
					uint32_t b = src[j * T + k];
					if (P.with_delta && !delta_stream_start(i, cb)) {
						const uint32_t pb = j ? src[(j - 1) * T + k] : src[(n - 1) * T + (k - 1)];
						b = (b - pb) & 0xFFu;
					}
					dst[i] = (uint8_t)b;
				}
		}
	}

	template<int T>
	__global__ void __launch_bounds__(FILTER_THREADS) unshuffle_kernel(FilterParams P)
	{
		const uint64_t c = P.chunk_in_y ? blockIdx.y : blockIdx.x;
		const uint64_t cb = min(P.chunk, P.bytes - c * P.chunk);
		const uint64_t n = cb / T;
		const uint64_t j0 = ((uint64_t)(P.chunk_in_y ? blockIdx.x : blockIdx.y) * FILTER_THREADS + threadIdx.x) * 16;
		const uint8_t* src = P.src + c * P.chunk;
		uint8_t* dst = P.dst + c * P.chunk;
		if (j0 >= n) {
			if (j0 < n + 16)
				for (uint64_t i = n * T; i < cb; ++i)
					dst[i] = src[i];
			return;
		}
		const bool vec = (j0 + 16 <= n) && ((n & 15u) == 0) && ((((uintptr_t)src) & 15u) == 0) && ((((uintptr_t)dst) & 15u) == 0);
		if (vec) {
			uint4 pl[T];
#pragma unroll
			for (int k = 0; k < T; ++k)
				pl[k] = *reinterpret_cast<const uint4*>(src + (uint64_t)k * n + j0);
			uint32_t w[4 * T];
			untranspose16<T>(pl, w);
			store_elements16<T>(dst + j0 * T, w);
		}
		else {
			const uint64_t j1 = min(j0 + 16, n);
			for (uint64_t j = j0; j < j1; ++j)
				for (int k = 0; k < T; ++k)
					dst[j * T + k] = src[(uint64_t)k * n + j];
		}
	}

	// plain delta (no transpose): a thread handles DELTA_GROUPS groups of 16 output bytes, FILTER_THREADS groups apart (one CTA:
	// 16 KiB -- with one group per thread, 4 GiB were a million 4 KiB CTAs and 0.58 of HBM bandwidth)
	constexpr int DELTA_GROUPS = 4;
	__global__ void __launch_bounds__(FILTER_THREADS) delta_kernel(FilterParams P)
	{
		const uint64_t nchunks = (P.bytes + P.chunk - 1) / P.chunk;
		const uint64_t groups_per_chunk = (P.chunk + 15) / 16;
		// (chunks are powers of two in practice -- 128 KiB << shift: a shift instead of a 64-bit division)
		const bool pow2 = (groups_per_chunk & (groups_per_chunk - 1)) == 0;
		const int sh = 63 - __clzll((long long)groups_per_chunk);
#pragma unroll
		for (int u = 0; u < DELTA_GROUPS; ++u) {
			const uint64_t g = ((uint64_t)blockIdx.x * DELTA_GROUPS + u) * FILTER_THREADS + threadIdx.x;
			const uint64_t c = pow2 ? g >> sh : g / groups_per_chunk;
			if (c >= nchunks)
				continue;
			const uint64_t cb = min(P.chunk, P.bytes - c * P.chunk);
			const uint64_t i0 = (g - c * groups_per_chunk) * 16;
			if (i0 >= cb)
				continue;
			const uint8_t* src = P.src + c * P.chunk;
			uint8_t* dst = P.dst + c * P.chunk;
			const uint64_t i1 = min(i0 + 16, cb);
			// whole 16-byte groups with no stream start inside: one 128-bit load, the byte before the group, four SIMD byte
			// subtractions, one 128-bit store
			const uint64_t q = cb > 2048 ? cb / 4 : ~0ull;
			const bool start_inside = i0 == 0 || (q != ~0ull && ((q - i0 < 16ull) || (2 * q - i0 < 16ull) || (3 * q - i0 < 16ull)));
			if (i1 - i0 == 16 && !start_inside && (((uintptr_t)(src + i0) | (uintptr_t)(dst + i0)) & 15u) == 0) {
				const uint4 v = *reinterpret_cast<const uint4*>(src + i0);
				const uint32_t a[4] = { v.x, v.y, v.z, v.w };
				uint32_t d[4];
				row_deltas(a, (uint32_t)src[i0 - 1] << 24, d);
				*reinterpret_cast<uint4*>(dst + i0) = make_uint4(d[0], d[1], d[2], d[3]);
				continue;
			}
			for (uint64_t i = i0; i < i1; ++i)
				dst[i] = delta_stream_start(i, cb) ? src[i] : (uint8_t)(src[i] - src[i - 1]);
		}
	}

	// inverse delta (delta.cpp:230-267): one CTA per (chunk, stream); the stream is scanned tile by tile (16 bytes per
	// thread, one 128-bit load and store each) with a running carry.  Per tile: byte sums of the 16-byte groups
	// (VABSDIFF4), one CTA-wide scan of them, then the in-group prefix sums on top of the group's carry
	// (prefix16: 16-bit lanes, one IMAD per word).
#ifndef DELTA_INV_NT
#define DELTA_INV_NT 128
#endif
	constexpr int DELTA_INV_THREADS = DELTA_INV_NT; // (1024 until round 2: four barriers of 32 warps per 16 KiB tile)
	__global__ void __launch_bounds__(DELTA_INV_THREADS) delta_inv_kernel(FilterParams P)
	{
		STENOS_DYN_SMEM(uint32_t, warp_sums);
		const uint64_t c = blockIdx.x >> 2;
		const uint32_t stream = blockIdx.x & 3u;
		const uint64_t cb = min(P.chunk, P.bytes - c * P.chunk);
		uint64_t lo, hi; // byte range of this stream inside the chunk
		if (cb > 2048) {
			const uint64_t q = cb / 4;
			lo = stream * q;
			hi = stream == 3 ? cb : lo + q;
		}
		else {
			if (stream)
				return;
			lo = 0;
			hi = cb;
		}
		const uint8_t* src = P.src + c * P.chunk;
		uint8_t* dst = P.dst + c * P.chunk;
		const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
		const bool aligned = (((uintptr_t)(src + lo) | (uintptr_t)(dst + lo)) & 15u) == 0;
		uint32_t carry = 0;
		for (uint64_t t0 = lo; t0 < hi; t0 += (uint64_t)DELTA_INV_THREADS * 16) {
			const uint64_t i0 = t0 + (uint64_t)tid * 16;
			const bool full = aligned && i0 + 16 <= hi;
			uint32_t x[4] = { 0u, 0u, 0u, 0u };
			if (full) {
				const uint4 v = *reinterpret_cast<const uint4*>(src + i0);
				x[0] = v.x;
				x[1] = v.y;
				x[2] = v.z;
				x[3] = v.w;
			}
			else {
#pragma unroll
				for (int k = 0; k < 16; ++k)
					if (i0 + k < hi)
						x[k >> 2] |= (uint32_t)src[i0 + k] << (8 * (k & 3));
			}
			uint32_t sum = 0;
#pragma unroll
			for (int j = 0; j < 4; ++j)
				sum = sad4_acc(x[j], 0u, sum);
			// CTA-wide exclusive scan of the per-thread sums (mod 256)
			uint32_t incl = sum;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				uint32_t t = __shfl_up_sync(FULL, incl, d);
				if (lane >= d)
					incl += t;
			}
			if (lane == 31)
				warp_sums[warp] = incl;
			__syncthreads();
			const uint32_t ws = lane < DELTA_INV_THREADS / 32 ? warp_sums[lane] : 0u;
			const uint32_t wpre = __reduce_add_sync(FULL, lane < warp ? ws : 0u);
			const uint32_t tot = __reduce_add_sync(FULL, ws);
			__syncthreads();
			const uint32_t add = (carry + wpre + (incl - sum)) & 0xFFu;
			uint32_t o[4];
			prefix16(x, 0u, add, o);
			if (full)
				*reinterpret_cast<uint4*>(dst + i0) = make_uint4(o[0], o[1], o[2], o[3]);
			else {
#pragma unroll
				for (int k = 0; k < 16; ++k)
					if (i0 + k < hi)
						dst[i0 + k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
			}
			carry = (carry + tot) & 0xFFu;
		}
	}

	// ------------------------------------------------------------------------------------------
	// unshuffle with the inverse delta fused in (stenos.cpp:711-725: delta_inv, then unshuffle): one CTA per chunk
	// walks it tile by tile (16 elements per thread); every byte plane carries its own running prefix sum, since for a
	// fixed plane the positions k * n + j of the transposed chunk grow with j.  HBM traffic 2N instead of the 4N of
	// delta_inv into scratch + unshuffle.  Where a quarter stream starts (delta.cpp:42-70) the sum restarts: a group
	// is an affine map last = a * prev + c with a = 0 if it holds a restart, and the CTA scans maps, two planes per
	// register (compose_maps2).  Requires cb % (16 * T) == 0 and 16-byte aligned chunks (the host checks).
	// ------------------------------------------------------------------------------------------
	// threads of a CTA (one CTA per chunk).  Measured with the next tile staged by cp.async, 2N/t over HBM bandwidth at
	// 256 KiB chunks: 256 threads 0.83 / 0.82 / 0.59 (T = 2 / 4 / 8), 128: 0.85 / 0.89 / 0.63, 64: 0.86 / 0.92 / 0.63,
	// 32: 0.81 / 0.91 / 0.65 -- small CTAs: more of them per SM, cheap barriers.
#ifdef UNSHUFFLE_DELTA_NT
	template<int T>
	struct UnshuffleDeltaThreads
	{
		static constexpr int N = UNSHUFFLE_DELTA_NT;
	};
#else
	template<int T>
	struct UnshuffleDeltaThreads
	{
		static constexpr int N = T == 8 ? 32 : 64;
	};
#endif
	template<int T>
	__global__ void __launch_bounds__(UnshuffleDeltaThreads<T>::N) unshuffle_delta_kernel(FilterParams P)
	{
		constexpr int UNSHUFFLE_DELTA_THREADS = UnshuffleDeltaThreads<T>::N;
		constexpr int NW = UNSHUFFLE_DELTA_THREADS / 32;
		constexpr int NP = T / 2; // plane pairs
		STENOS_DYN_SMEM(uint32_t, warp_maps_raw); // [NP][NW] (256 bytes reserved), then the staging slots of the next tile
		uint32_t (*warp_maps)[NW] = reinterpret_cast<uint32_t (*)[NW]>(warp_maps_raw);
		const uint64_t c = blockIdx.x;
		const uint64_t cb = min(P.chunk, P.bytes - c * P.chunk);
		const uint64_t n = cb / T;
		const uint8_t* src = P.src + c * P.chunk;
		uint8_t* dst = P.dst + c * P.chunk;
		const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
		// The 16 bytes a thread needs of every plane for the NEXT tile travel to shared memory with cp.async while the
		// current tile is scanned and written (with plain loads at the top of a tile the kernel waited 6.7 cycles per
		// instruction for them: one tile of loads in flight per thread, 46 % issue utilisation).  A thread reads back only
		// what it copied itself: no barrier; slot (k, tid) at (k * NT + tid) * 16, conflict free.
		const uint32_t stage32 = smem_addr32(warp_maps_raw) + 256u + 16u * (uint32_t)threadIdx.x;
		auto fetch_tile = [&](uint64_t j0) {
			if (j0 < (cb / T)) {
#pragma unroll
				for (int k = 0; k < T; ++k)
					cp_async16(stage32 + (uint32_t)k * 16u * UNSHUFFLE_DELTA_THREADS, P.src + blockIdx.x * P.chunk + (uint64_t)k * (cb / T) + j0);
			}
			cp_async_commit();
		};
		const uint64_t q1 = cb > 2048 ? cb / 4 : ~0ull; // starts of the quarter streams 1..3
		const uint64_t q2 = cb > 2048 ? 2 * (cb / 4) : ~0ull, q3 = cb > 2048 ? 3 * (cb / 4) : ~0ull;
		uint32_t run[NP]; // running carries (c only), two planes per register
#pragma unroll
		for (int pp = 0; pp < NP; ++pp)
			run[pp] = 0u;
		fetch_tile((uint64_t)tid * 16); // tile 0
		if (T == 8) {
			// n = q / 2: an odd plane starts in the middle of a quarter stream, on top of the sum of the plane before it.
			// One extra read of the even planes (they are read again below, from L2).
			uint32_t s[NP];
#pragma unroll
			for (int e = 0; e < NP; ++e)
				s[e] = 0u;
			constexpr uint64_t STEP = (uint64_t)UNSHUFFLE_DELTA_THREADS * 16;
			uint64_t j = (uint64_t)tid * 16;
			for (; j + STEP < n; j += 2 * STEP) { // two tiles of loads in flight
				uint4 v[NP], w[NP];
#pragma unroll
				for (int e = 0; e < NP; ++e) {
					v[e] = *reinterpret_cast<const uint4*>(src + (uint64_t)(2 * e) * n + j);
					w[e] = *reinterpret_cast<const uint4*>(src + (uint64_t)(2 * e) * n + j + STEP);
				}
#pragma unroll
				for (int e = 0; e < NP; ++e) {
					s[e] = sad4_acc(v[e].x, 0u, sad4_acc(v[e].y, 0u, sad4_acc(v[e].z, 0u, sad4_acc(v[e].w, 0u, s[e]))));
					s[e] = sad4_acc(w[e].x, 0u, sad4_acc(w[e].y, 0u, sad4_acc(w[e].z, 0u, sad4_acc(w[e].w, 0u, s[e]))));
				}
			}
			for (; j < n; j += STEP) {
#pragma unroll
				for (int e = 0; e < NP; ++e) {
					const uint4 v = *reinterpret_cast<const uint4*>(src + (uint64_t)(2 * e) * n + j);
					s[e] = sad4_acc(v.x, 0u, sad4_acc(v.y, 0u, sad4_acc(v.z, 0u, sad4_acc(v.w, 0u, s[e]))));
				}
			}
#pragma unroll
			for (int e = 0; e < NP; ++e) {
				const uint32_t ws = __reduce_add_sync(FULL, s[e]);
				if (lane == 0)
					warp_maps[e][warp] = ws;
			}
			__syncthreads();
#pragma unroll
			for (int e = 0; e < NP; ++e) {
				const uint32_t tot = __reduce_add_sync(FULL, lane < NW ? warp_maps[e][lane] : 0u);
				run[e] = (tot & 0xFFu) << 16; // planes 2e (starts a stream: 0) and 2e + 1
			}
			__syncthreads();
		}
		for (uint64_t t0 = 0; t0 < n; t0 += (uint64_t)UNSHUFFLE_DELTA_THREADS * 16) {
			const uint64_t j0 = t0 + (uint64_t)tid * 16;
			const bool live = j0 < n; // n is a multiple of 16
			uint4 pl[T];
			uint32_t rst[T]; // offset (0..15) of a stream start inside the plane's 16 bytes, 16: none
			uint32_t z[NP];
#pragma unroll
			for (int pp = 0; pp < NP; ++pp)
				z[pp] = 0x01000100u; // identity maps
			cp_async_wait_all(); // my 16 bytes of every plane of this tile are in my slots
#pragma unroll
			for (int k = 0; k < T; ++k) {
				pl[k] = make_uint4(0u, 0u, 0u, 0u);
				rst[k] = 16u;
				if (live) {
					const uint64_t i0 = (uint64_t)k * n + j0;
					pl[k] = lds_u128(stage32 + (uint32_t)k * 16u * UNSHUFFLE_DELTA_THREADS);
					if (q1 - i0 < 16ull)
						rst[k] = (uint32_t)(q1 - i0);
					if (q2 - i0 < 16ull)
						rst[k] = (uint32_t)(q2 - i0);
					if (q3 - i0 < 16ull)
						rst[k] = (uint32_t)(q3 - i0);
					uint32_t m;
					if (rst[k] == 16u)
						m = 0x100u | ((sad4_acc(pl[k].x, 0u, 0u) + sad4_acc(pl[k].y, 0u, 0u) + sad4_acc(pl[k].z, 0u, 0u) + sad4_acc(pl[k].w, 0u, 0u)) & 0xFFu);
					else {
						uint32_t sum = 0; // bytes from the restart on
						for (uint32_t b = rst[k]; b < 16u; ++b)
							sum += byte16(pl[k], b);
						m = sum & 0xFFu;
					}
					z[k >> 1] = (z[k >> 1] & ~(0x1FFu << (16 * (k & 1)))) | (m << (16 * (k & 1)));
				}
			}
			fetch_tile(j0 + (uint64_t)UNSHUFFLE_DELTA_THREADS * 16); // the next tile, into the slots just read
			// inclusive scan of the maps over the CTA: warp, then the warps' totals
			uint32_t incl[NP];
#pragma unroll
			for (int pp = 0; pp < NP; ++pp) {
				uint32_t x = z[pp];
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t y = __shfl_up_sync(FULL, x, d);
					if (lane >= d)
						x = compose_maps2(x, y);
				}
				incl[pp] = x;
				if (lane == 31)
					warp_maps[pp][warp] = x;
			}
			__syncthreads();
#pragma unroll
			for (int pp = 0; pp < NP; ++pp) {
				// maps of the warps before mine, composed; then everything before this thread
				uint32_t w = lane < NW ? warp_maps[pp][lane] : 0x01000100u;
#pragma unroll
				for (int d = 1; d < NW; d <<= 1) {
					const uint32_t y = __shfl_up_sync(FULL, w, d);
					if (lane >= d)
						w = compose_maps2(w, y);
				}
				const uint32_t tile_total = __shfl_sync(FULL, w, NW - 1);
				uint32_t before_warp = __shfl_sync(FULL, w, (warp + NW - 1) % NW);
				if (warp == 0)
					before_warp = 0x01000100u;
				uint32_t excl = __shfl_up_sync(FULL, incl[pp], 1); // threads before me in my warp
				if (lane == 0)
					excl = 0x01000100u;
				// carry into this thread = (excl o before_warp)(run)
				const uint32_t m = compose_maps2(excl, before_warp);
				const uint32_t carry2 = compose_maps2(m, run[pp]) & 0x00FF00FFu;
				run[pp] = compose_maps2(tile_total, run[pp]) & 0x00FF00FFu;
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					const int k = 2 * pp + h;
					const uint32_t carry = (carry2 >> (16 * h)) & 0xFFu;
					uint32_t x[4] = { pl[k].x, pl[k].y, pl[k].z, pl[k].w }, o[4];
					if (rst[k] == 16u)
						prefix16(x, 0u, carry, o);
					else {
						// a stream starts inside the group: byte by byte
						uint32_t acc = carry;
						o[0] = o[1] = o[2] = o[3] = 0u;
						for (uint32_t b = 0; b < 16u; ++b) {
							if (b == rst[k])
								acc = 0u;
							acc = (acc + byte16(pl[k], b)) & 0xFFu;
							o[b >> 2] |= acc << (8 * (b & 3));
						}
					}
					pl[k] = make_uint4(o[0], o[1], o[2], o[3]);
				}
			}
			__syncthreads();
			if (live) {
				uint32_t w[4 * T];
				untranspose16<T>(pl, w);
				store_elements16<T>(dst + j0 * T, w);
			}
		}
	}
}
