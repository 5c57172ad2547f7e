// sb_common.cuh -- shared definitions for the sm_100a level-1 stenos path.
//
// Format constants restate the reference's stream format (file:line relative to /root/reference):
//   plane kinds           stenos/internal/block_compress.h:52-55
//   block markers         stenos/internal/block_compress.h:57-60
//   superblock codes      stenos/internal/stenos.cpp:34-39
//   error codes           stenos/stenos.h:75-84
#pragma once

#ifdef STENOS_EMU
#include "cuda_emu.h"
#define STENOS_SPIN_HINT() emu::yield()
#define STENOS_SPIN_WAIT() emu::yield()
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#define STENOS_DYN_SMEM(type, name)                       \
	extern __shared__ __align__(16) uint8_t stenos_dyn_smem_raw[]; \
	type* name = reinterpret_cast<type*>(stenos_dyn_smem_raw)
#define STENOS_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#ifndef STENOS_SPIN_HINT_NS
#define STENOS_SPIN_HINT_NS 20
#endif
#define STENOS_SPIN_HINT() __nanosleep(STENOS_SPIN_HINT_NS)
#define STENOS_SPIN_WAIT() __nanosleep(200) // waits that are expected to be long
#endif

namespace sb
{
	constexpr uint32_t FULL = 0xffffffffu;

	enum : int { KIND_SAME = 0, KIND_RAW = 1, KIND_NORMAL = 2, KIND_NORMAL_RLE = 3 };
	enum : int { MARK_COPY = 252, MARK_LZ = 253, MARK_PARTIAL = 254 };
	enum : int { CODE_BLOCK = 1, CODE_ZSTD = 2, CODE_COPY = 6 };

	constexpr uint64_t ERR_UNDEFINED = (uint64_t)-1;
	constexpr uint64_t ERR_SRC_OVERFLOW = (uint64_t)-2;
	constexpr uint64_t ERR_ALLOC = (uint64_t)-3;
	constexpr uint64_t ERR_INVALID_INPUT = (uint64_t)-4;
	constexpr uint64_t ERR_INVALID_INSTRUCTION_SET = (uint64_t)-5;
	constexpr uint64_t ERR_DST_OVERFLOW = (uint64_t)-6;
	constexpr uint64_t ERR_INVALID_BYTESOFTYPE = (uint64_t)-7;
	constexpr uint64_t ERR_ZSTD_INTERNAL = (uint64_t)-8;
	constexpr uint64_t ERR_INVALID_PARAMETER = (uint64_t)-9;
	constexpr uint64_t LAST_ERROR_CODE = (uint64_t)-100;

	constexpr uint32_t DEFAULT_SUPERBLOCK = 131072u; // STENOS_BLOCK_SIZE, stenos/stenos.h:57

	// Device error bits accumulated by kernels (mapped to the codes above by the host layer)
	enum : uint32_t { DEV_ERR_DST_OVERFLOW = 1u, DEV_ERR_SRC_OVERFLOW = 2u, DEV_ERR_INVALID_INPUT = 4u };

	// ------------------------------------------------------------------------------------------
	// byte-SIMD helpers on packed 4 x u8 words
	// ------------------------------------------------------------------------------------------

	// PTX prmt.b32 in its default mode: selector nibble bit 3 replicates the SIGN of the selected byte.
	// (__byte_perm masks the selector with 0x7777, so it cannot express sign replication.)
	__device__ __forceinline__ uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t sel)
	{
#ifdef STENOS_EMU
		return emu::prmt_b32(a, b, sel);
#else
		uint32_t r;
		asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
		return r;
#endif
	}

	// 0x80 in every byte of x that is zero, 0 elsewhere (exact, no cross-byte borrow)
	__device__ __forceinline__ uint32_t zero_bytes(uint32_t x)
	{
		uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
		return ~(t | x | 0x7F7F7F7Fu);
	}
	// gathers the four 0x80 flags of z into bits 0..3
	__device__ __forceinline__ uint32_t flags_to_mask4(uint32_t z) { return (((z >> 7) * 0x00204081u) >> 21) & 0xFu; }

	// replicate a byte into the four byte lanes
	__device__ __forceinline__ uint32_t splat(uint32_t b) { return (b & 0xFFu) * 0x01010101u; }

	// [hi byte 3 of `before`, bytes 0..2 of `cur`] : the word of "previous bytes" of cur
	__device__ __forceinline__ uint32_t prev_bytes(uint32_t before, uint32_t cur) { return __byte_perm(before, cur, 0x6543); }

	// inclusive prefix sum over the 4 bytes of x (mod 256 per byte)
	__device__ __forceinline__ uint32_t prefix4(uint32_t x)
	{
		x = __vadd4(x, x << 8);
		x = __vadd4(x, x << 16);
		return x;
	}

	// 64-bit flag+value words of the decoupled look-back: single-copy atomic, never cached in L1
	__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
	__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }

	__device__ __forceinline__ uint32_t lanemask_lt(int lane) { return (1u << lane) - 1u; }

	// warms L1 with the line holding p: a load whose result is never used (no scoreboard wait follows it)
	__device__ __forceinline__ void touch_l1(const void* p)
	{
#ifndef STENOS_EMU
		uint32_t unused;
		asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(unused) : "l"(p));
#else
		(void)p;
#endif
	}
	// Asks for the line holding p to be brought into L1 without tying up a register or a scoreboard: an asynchronous
	// 4-byte copy (LDGSTS, cached at all levels) into a scratch word of shared memory that nobody reads.  A load into
	// an unused register (touch_l1) is only half asynchronous: the register's next writer waits for it.
	__device__ __forceinline__ void prefetch_l1_async(const void* p, void* smem_scratch_word)
	{
#ifndef STENOS_EMU
		const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_scratch_word);
		const uintptr_t g = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3;
		asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(g) : "memory");
#else
		(void)p;
		(void)smem_scratch_word;
#endif
	}
	// 32-byte store (STG.E.256, sm_100): a thread that owns 32 contiguous bytes writes a whole sector in one instruction
	// instead of two half-sector writes.  p must be 32-byte aligned.
	__device__ __forceinline__ void st_global_256(void* p, uint4 a, uint4 b)
	{
#ifndef STENOS_EMU
		asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
#else
		reinterpret_cast<uint4*>(p)[0] = a;
		reinterpret_cast<uint4*>(p)[1] = b;
#endif
	}
	// 32-byte load (LDG.E.256, sm_100).  p must be 32-byte aligned.
	__device__ __forceinline__ void ld_global_256(const void* p, uint4& a, uint4& b)
	{
#ifndef STENOS_EMU
		asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
#else
		a = reinterpret_cast<const uint4*>(p)[0];
		b = reinterpret_cast<const uint4*>(p)[1];
#endif
	}
	__device__ __forceinline__ void prefetch_l2(const void* p)
	{
#ifndef STENOS_EMU
		asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
		(void)p;
#endif
	}

	// 4x4 byte transpose: words a,b,c,d (one per element) -> p0..p3 (one per byte plane)
	__device__ __forceinline__ void transpose4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t& p0, uint32_t& p1, uint32_t& p2, uint32_t& p3)
	{
		uint32_t t0 = __byte_perm(a, b, 0x5140); // a0 b0 a1 b1
		uint32_t t1 = __byte_perm(c, d, 0x5140); // c0 d0 c1 d1
		uint32_t t2 = __byte_perm(a, b, 0x7362); // a2 b2 a3 b3
		uint32_t t3 = __byte_perm(c, d, 0x7362);
		p0 = __byte_perm(t0, t1, 0x5410);
		p1 = __byte_perm(t0, t1, 0x7632);
		p2 = __byte_perm(t2, t3, 0x5410);
		p3 = __byte_perm(t2, t3, 0x7632);
	}

	// byte i of the 8 * T bytes a lane holds in w[0 .. 2T)
	template<int T>
	__device__ __forceinline__ uint32_t lane_byte(const uint32_t (&w)[2 * T], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu; }

	// Lane-local view of one 256-element block: lane l owns elements 8l..8l+7, i.e. for every byte
	// plane p the 8 bytes lo[p] (elements 0..3) and hi[p] (elements 4..7).  W = 2*T input words.
	template<int T>
	__device__ __forceinline__ void words_to_planes(const uint32_t (&w)[2 * T], uint32_t (&lo)[T], uint32_t (&hi)[T])
	{
		if (T == 2) {
			lo[0] = __byte_perm(w[0], w[1], 0x6420);
			lo[1] = __byte_perm(w[0], w[1], 0x7531);
			hi[0] = __byte_perm(w[2], w[3], 0x6420);
			hi[1] = __byte_perm(w[2], w[3], 0x7531);
		}
		else if (T == 4) {
			transpose4(w[0], w[1], w[2], w[3], lo[0], lo[1], lo[2], lo[3]);
			transpose4(w[4], w[5], w[6], w[7], hi[0], hi[1], hi[2], hi[3]);
		}
		else if (T == 8) {
			transpose4(w[0], w[2], w[4], w[6], lo[0], lo[1], lo[2], lo[3]);
			transpose4(w[1], w[3], w[5], w[7], lo[4], lo[5], lo[6], lo[7]);
			transpose4(w[8], w[10], w[12], w[14], hi[0], hi[1], hi[2], hi[3]);
			transpose4(w[9], w[11], w[13], w[15], hi[4], hi[5], hi[6], hi[7]);
		}
		else {
			// any other element size (3, 6: SURVEY 8 f3): byte p of element j is byte j * T + p of the lane's 8 * T bytes
#pragma unroll
			for (int p = 0; p < T; ++p) {
				uint32_t l = 0, h = 0;
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					l |= lane_byte<T>(w, j * T + p) << (8 * j);
					h |= lane_byte<T>(w, (4 + j) * T + p) << (8 * j);
				}
				lo[p] = l;
				hi[p] = h;
			}
		}
	}
	template<int T>
	__device__ __forceinline__ void planes_to_words(const uint32_t (&lo)[T], const uint32_t (&hi)[T], uint32_t (&w)[2 * T])
	{
		// transpose4 is an involution on the 4x4 byte matrix
		if (T == 2) {
			w[0] = __byte_perm(lo[0], lo[1], 0x5140);
			w[1] = __byte_perm(lo[0], lo[1], 0x7362);
			w[2] = __byte_perm(hi[0], hi[1], 0x5140);
			w[3] = __byte_perm(hi[0], hi[1], 0x7362);
		}
		else if (T == 4) {
			transpose4(lo[0], lo[1], lo[2], lo[3], w[0], w[1], w[2], w[3]);
			transpose4(hi[0], hi[1], hi[2], hi[3], w[4], w[5], w[6], w[7]);
		}
		else if (T == 8) {
			transpose4(lo[0], lo[1], lo[2], lo[3], w[0], w[2], w[4], w[6]);
			transpose4(lo[4], lo[5], lo[6], lo[7], w[1], w[3], w[5], w[7]);
			transpose4(hi[0], hi[1], hi[2], hi[3], w[8], w[10], w[12], w[14]);
			transpose4(hi[4], hi[5], hi[6], hi[7], w[9], w[11], w[13], w[15]);
		}
		else {
#pragma unroll
			for (int i = 0; i < 2 * T; ++i) {
				uint32_t v = 0;
#pragma unroll
				for (int b = 0; b < 4; ++b) {
					const int idx = 4 * i + b, j = idx / T, p = idx % T; // byte p of element j
					v |= (((j < 4 ? lo[p] : hi[p]) >> (8 * (j & 3))) & 0xFFu) << (8 * b);
				}
				w[i] = v;
			}
		}
	}

	// 16-byte vector load/store of a lane's 8 elements (block start must be 16-byte aligned)
	template<int T>
	__device__ __forceinline__ void load_lane_words(const uint8_t* __restrict__ block, int lane, uint32_t (&w)[2 * T])
	{
		if (T % 2 != 0 || ((8 * T) % 16) != 0) {
			// 8 * T bytes per lane that are not a multiple of 16: 8-byte loads (the lane's bytes start 8-byte aligned)
			const uint2* q = reinterpret_cast<const uint2*>(block + (size_t)lane * 8 * T);
#pragma unroll
			for (int i = 0; i < T; ++i) {
				const uint2 v = q[i];
				w[2 * i] = v.x;
				w[2 * i + 1] = v.y;
			}
			return;
		}
		const uint4* p = reinterpret_cast<const uint4*>(block + (size_t)lane * 8 * T);
#pragma unroll
		for (int i = 0; i < T / 2; ++i) {
			uint4 v = p[i];
			w[4 * i + 0] = v.x;
			w[4 * i + 1] = v.y;
			w[4 * i + 2] = v.z;
			w[4 * i + 3] = v.w;
		}
	}
	template<int T>
	__device__ __forceinline__ void store_lane_words(uint8_t* __restrict__ block, int lane, const uint32_t (&w)[2 * T])
	{
		if (T % 2 != 0 || ((8 * T) % 16) != 0) {
			uint2* q = reinterpret_cast<uint2*>(block + (size_t)lane * 8 * T);
#pragma unroll
			for (int i = 0; i < T; ++i)
				q[i] = make_uint2(w[2 * i], w[2 * i + 1]);
			return;
		}
		uint4* p = reinterpret_cast<uint4*>(block + (size_t)lane * 8 * T);
#pragma unroll
		for (int i = 0; i < T / 2; ++i)
			p[i] = make_uint4(w[4 * i + 0], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
	}

	// unaligned little-endian reads from a byte stream (global or shared)
	__device__ __forceinline__ uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
	__device__ __forceinline__ uint32_t rd24(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }
	// reads 8 bytes at an arbitrary address through three aligned 32-bit loads
	__device__ __forceinline__ void rd64_unaligned(const uint8_t* p, uint32_t& lo, uint32_t& hi)
	{
		uintptr_t a = reinterpret_cast<uintptr_t>(p);
		const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
		uint32_t sh = (uint32_t)(a & 3) * 8;
		uint32_t w0 = q[0], w1 = q[1];
		uint32_t w2 = sh ? q[2] : 0u;
		lo = __funnelshift_r(w0, w1, sh);
		hi = __funnelshift_r(w1, w2, sh);
	}
}
