// sb_flow.cuh -- encode_flow_kernel: the level-1 frame encoder of the fast path (round 2).
//
// Same job and same contract as encode_stream_kernel (sb_stream.cuh), which it replaces as the default:
// stenos_compress_generic's superblock loop (stenos.cpp:893-904) and compress_generic_superblock's level-1
// branch (:403-450, :606-615, :363-374) for the superblocks whose dst-room checks are provably inert
// (SURVEY.md appendix C2).  What changed is where the instructions go.  The round-1 pipeline spent a third
// of its warp instructions on bookkeeping (a ticket per block pair, a decoupled look-back per pair to place
// it in a shared ring, spinning on late predecessors) and analysed every plane twice (sizes first, bytes
// after the placement).  Here:
//
//   * A HALF-WARP owns a run of K consecutive 256-element blocks of the superblock (a PIECE) and a private
//     staging ring in shared memory.  The position of a block's bytes is the half's own cursor: known
//     before the block is analysed, so every byte plane is analysed and emitted in ONE pass, and nothing
//     inside a superblock waits for anything else (no chain, no per-pair atomics).
//   * One extra warp per CTA, the PLACER, does nothing but wait for a superblock's pieces: it scans the 2*NW
//     piece sizes, does the decoupled look-back across superblocks (the same words encode_frame_kernel uses),
//     writes [code][csize:3] and publishes the frame offset of every piece; each half-warp then copies its own
//     piece to HBM the next time it passes a superblock boundary (or needs ring space).  COPY superblocks copy
//     their input instead.  (Giving this duty to the warp that arrives last was measured first: the slowest
//     warp got slower every turn and the others spun on it, 1.75 ms instead of 0.74 ms.)
//   * Superblocks come from the global ticket for every CTA, including the first one, so every superblock a
//     look-back waits for belongs to a CTA that is running (no co-residency assumption).
//   * A warp never waits in the middle of a task: when its ring is full of pieces whose superblocks are not placed
//     yet (they may be waiting for this very task), the oldest one moves to a spill slot in HBM and is copied to the
//     frame from there.  Rare (incompressible data); it is what makes the hand-out of tasks deadlock free.
//
// HBM traffic: N read + C written, nothing else.
#pragma once
#include "sb_kernels.cuh"
#include "sb_stream.cuh"

namespace sb
{
	// How the input reaches the registers.  2 (default): the rows of a half-warp's NEXT block are copied to shared memory
	// with cp.async (LDGSTS, 16 bytes per lane and instruction, conflict-free layout) while the current block is
	// encoded, and read back with LDS.128 when their turn comes.  1: the same look-ahead into registers (16 * T / 4 more
	// live registers: spills at 96).  0: loads at the top of a block behind an L2 prefetch (the first version: the loads'
	// wait was 12 % of the warps' time).
#ifndef FLOW_STAGING
#define FLOW_STAGING 2
#endif
	// waits of the flow kernel that last microseconds (ring space, a free slot, a placement): the polling warp must not
	// compete for issue slots with the warps that encode (__nanosleep returns well before its nominal time)
#ifdef STENOS_EMU
#define FLOW_LONG_WAIT() emu::yield()
#else
#ifndef FLOW_LONG_WAIT_NS
#define FLOW_LONG_WAIT_NS 1000
#endif
#define FLOW_LONG_WAIT() __nanosleep(FLOW_LONG_WAIT_NS)
#endif
	// Row decision table (find_pack_bits_params, block_compress.h:334-352, :420-435, :499-502): indexed by the row's value
	// range (clipped to 32: everything above needs 8 bits) and delta range (clipped to 64), an entry holds
	// header nibble | payload bytes << 4 | row size << 9 | plain << 14 -- one LDS instead of ~25 integer instructions.
	constexpr uint32_t FLOW_HLUT_RB = 33, FLOW_HLUT_RD = 65, FLOW_HLUT_N = FLOW_HLUT_RB * FLOW_HLUT_RD;
	__device__ __forceinline__ uint32_t flow_hlut_entry(uint32_t rb, uint32_t rd)
	{
		const uint32_t b0 = rb >= 32u ? 8u : (32u - (uint32_t)__clz((int)rb)); // 7 -> 8, and 6 -> 8 for the plain type
		const uint32_t b1 = rd >= 64u ? 8u : (32u - (uint32_t)__clz((int)rd));
		const uint32_t bits = min(b0, b1);
		const bool plain = (b0 == bits); // ties -> plain
		const uint32_t sz = 2u * bits + (bits != 8u ? 1u : 0u);
		const uint32_t h = plain ? (bits == 8u ? 15u : bits) : (8u + bits);
		const uint32_t pay = (bits == 8u) ? 16u : 2u * bits;
		return h | (pay << 4) | (sz << 9) | ((plain ? 1u : 0u) << 14);
	}
	constexpr int FLOW_NS = 8;  // superblocks in flight per CTA
	constexpr int FLOW_PQ = 8;   // pieces a half-warp may have waiting for their placement
	constexpr int FLOW_SBQ = 16; // ring of superblock numbers handed to the CTA (> FLOW_NS + 1)

	// T: element size; NT: threads of the encoder warps; KB: blocks per piece.  Pieces and tasks are independent of the
	// number of warps: NPC pieces (of KB blocks) and NTASK = NPC / 2 tasks per superblock, taken by whichever warp is free.
	template<int T, int NT, int KB>
	struct FlowLayout
	{
		static constexpr uint32_t BLOCK = T * 256u;
		static constexpr uint32_t HS = (T + 1) / 2;
		static constexpr uint32_t MAXB = BLOCK + HS; // worst full block (LZ: 1 + BLOCK)
		// free bytes a half's ring must have at the position of a block: the row stores of flow_store_row run up to
		// FLOW_STORE_SLACK bytes past the end of what they emit (overwritten by whatever comes next, never read)
		static constexpr uint32_t ROOM = MAXB + 16u;
		static constexpr uint32_t NW = NT / 32;
		static constexpr uint32_t NH = 2 * NW; // half-warps
		static constexpr uint32_t NBLK = DEFAULT_SUPERBLOCK / BLOCK;
		static constexpr uint32_t KMAX = KB;
		static constexpr uint32_t NTASK = (NBLK + 2u * KB - 1u) / (2u * KB); // tasks per superblock
		static constexpr uint32_t NPC = 2u * NTASK;                          // pieces per superblock
		static constexpr uint32_t TMP = (BLOCK + HS + 8u * T + 1u + 15u) & ~15u; // worst partial block
		static constexpr uint32_t HAS_LZ = (T % 4) == 0 ? 1u : 0u;
		static constexpr uint32_t LZ_STRIDE = (LZ_SCRATCH_BYTES + 15u) & ~15u;
		static constexpr uint32_t SLOT_BYTES = 64u + NPC * 8u;
		static constexpr uint32_t CTL_OFF = 0;                 // sbq[FLOW_SBQ] u64
		static constexpr uint32_t SLOT_OFF = FLOW_SBQ * 8u;
		static constexpr uint32_t PINFO_OFF = SLOT_OFF + FLOW_NS * SLOT_BYTES; // [NH][FLOW_PQ] x 16 bytes: pieces waiting for their placement
		static constexpr uint32_t PMETA_OFF = PINFO_OFF + NH * FLOW_PQ * 16u;   // [NW][FLOW_PQ] x 4 bytes: their task numbers
		static constexpr uint32_t TASK_OFF = PMETA_OFF + NW * FLOW_PQ * 4u;     // the CTA's task counter
		static constexpr uint32_t LUT_OFF = TASK_OFF + 16u;
		static constexpr uint32_t HLUT_OFF = LUT_OFF + 1024u; // row decision table, FLOW_HLUT_N x u16
		static constexpr uint32_t TAIL_OFF = HLUT_OFF + ((FLOW_HLUT_N * 2u + 15u) & ~15u);
		static constexpr uint32_t LZ_OFF = TAIL_OFF + TMP;
		static constexpr uint32_t IN_OFF = LZ_OFF + HAS_LZ * NW * LZ_STRIDE; // per warp: the rows of the block pair that comes next (cp.async)
		static constexpr uint32_t IN_STRIDE = FLOW_STAGING == 2 ? 32u * 16u * T : 0u;
		static constexpr uint32_t STAGE_OFF = IN_OFF + NW * IN_STRIDE;
		static constexpr uint32_t SMEM_TOTAL = 227u * 1024u;
		// per-half staging ring: at least two worst-case blocks (a piece that outgrows its ring continues in its spill slot)
		static constexpr uint32_t REG_MIN = 2u * ROOM;
		static constexpr uint32_t REG = ((SMEM_TOTAL - STAGE_OFF - 32u) / NH) & ~15u;
		static_assert(REG >= REG_MIN, "staging ring too small for this block size / warp count");
		static constexpr uint32_t smem_bytes() { return STAGE_OFF + NH * REG + 32u; }
		// HBM spill area (EncodeParams::spill): one slot per half-warp and waiting piece, worst case of a piece each
		static constexpr uint32_t SPILL_SLOT = (KB * MAXB + 15u) & ~15u;
		static constexpr size_t spill_bytes(size_t grid) { return grid * NH * (size_t)FLOW_PQ * SPILL_SLOT; }
	};

	// slot of one superblock in flight (shared memory); sizes / offs follow at +64
	struct FlowSlot
	{
		unsigned long long base; // frame offset of the superblock's header
		uint32_t owner;          // q + 1 of the superblock the slot is assigned to
		uint32_t started;        // warps that began it (the first one fetches the CTA's next superblock)
		uint32_t arrived;        // warps that published their pieces
		uint32_t consumed;       // warps that copied their pieces out
		uint32_t ready;          // q + 1 once base / offs / mode are valid
		uint32_t mode;           // 0 block stream, 1 COPY, 2 nothing to write (dst overflow)
		uint32_t sb;             // superblock number
		uint32_t agg;            // q + 1 once the piece offsets, csize and the look-back AGGREGATE word are published
		uint32_t csize;          // length of the block stream (partial tail block included)
		uint32_t tail_off;       // offset of the partial tail block in the stream
		uint32_t tail_sz;        // its size (0: none)
		uint32_t pad[3];
	};
	static_assert(sizeof(FlowSlot) == 64, "FlowSlot layout");

	// ------------------------------------------------------------------------------------------
	// Byte stores into the staging rings go through 32-bit shared-window addresses (st.shared): with generic
	// pointers the compiler rebuilt the window base (S2R SR_CgaCtaId + MOV + LEA + IADD) in front of every
	// predicated byte store -- 64 of the 110 instructions of a row store.
	// ------------------------------------------------------------------------------------------
#ifdef STENOS_EMU
	__device__ __forceinline__ uint32_t smem_addr32(const void* p) { return (uint32_t)(reinterpret_cast<const uint8_t*>(p) - emu::st().dyn_smem); }
	__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { emu::st().dyn_smem[a] = (uint8_t)v; }
	__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { emu::st().dyn_smem[a] = (uint8_t)v, emu::st().dyn_smem[a + 1] = (uint8_t)(v >> 8); }
#else
	__device__ __forceinline__ uint32_t smem_addr32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
	__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
	__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; st.shared.b16 [%0], lo; }" ::"r"(a), "r"(v)); }
#endif

#ifdef STENOS_EMU
	__device__ __forceinline__ void cp_async16(uint32_t a, const void* g) { memcpy(emu::st().dyn_smem + a, g, 16); }
	__device__ __forceinline__ void cp_async_commit() {}
	__device__ __forceinline__ void cp_async_wait_all() {}
	__device__ __forceinline__ uint4 lds_u128(uint32_t a) { return *reinterpret_cast<const uint4*>(emu::st().dyn_smem + a); }
#else
	__device__ __forceinline__ void cp_async16(uint32_t a, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g) : "memory"); }
	__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
	__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
	__device__ __forceinline__ uint4 lds_u128(uint32_t a)
	{
		uint4 v;
		asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
		return v;
	}
#endif
	// the lane's row of a block (16 * T bytes at g) -> the warp's input buffer; chunk k of lane l sits at (k * 32 + l) * 16,
	// so every instruction of the warp, cp.async or LDS.128, touches 512 consecutive bytes
	template<int T>
	__device__ __forceinline__ void flow_fetch_row(uint32_t in32, const uint8_t* __restrict__ g, int lane)
	{
#pragma unroll
		for (int k = 0; k < T; ++k)
			cp_async16(in32 + (uint32_t)(k * 32 + lane) * 16u, g + 16 * k);
	}
	template<int T>
	__device__ __forceinline__ void flow_read_row(uint32_t in32, int lane, uint32_t (&e)[4 * T])
	{
#pragma unroll
		for (int k = 0; k < T; ++k) {
			const uint4 v = lds_u128(in32 + (uint32_t)(k * 32 + lane) * 16u);
			e[4 * k + 0] = v.x, e[4 * k + 1] = v.y, e[4 * k + 2] = v.z, e[4 * k + 3] = v.w;
		}
	}

	// one step of an inclusive scan inside 16-lane segments: v += (v of the lane `delta` below, if it is in my segment);
	// the shuffle's own in-range predicate guards the add (no compare, no select)
	__device__ __forceinline__ uint32_t scan_up16_step(uint32_t v, int delta, int r)
	{
#ifdef STENOS_EMU
		const uint32_t t = __shfl_up_sync(FULL, v, delta, 16);
		return r >= delta ? v + t : v;
#else
		(void)r;
		asm("{ .reg .pred p; .reg .u32 t; shfl.sync.up.b32 t|p, %0, %1, 0x1000, 0xffffffff; @p add.u32 %0, %0, t; }" : "+r"(v) : "r"(delta));
		return v;
#endif
	}

	// ------------------------------------------------------------------------------------------
	// row payload helpers
	// ------------------------------------------------------------------------------------------

	// Stores the n bytes (0, or 2..16: no row payload is one byte long) of a row, w[0..3], at the shared-window address
	// `at` -- WITHOUT a predicate per byte (16 compares + 16 predicated byte stores + 12 shifts per row before).  A row
	// with n != 0 always stores 16 (17 when `at` is odd) bytes: eight aligned halfwords in DESCENDING address order, then
	// its first byte.  What runs past the row's n bytes lands on bytes that are written later: the rows after it (they
	// start >= 2 bytes further on, so their halfword of the same step is a different one, and the halfword that holds
	// valid data for a clobbered address belongs to an earlier step index = a LATER instruction; a row's first byte,
	// which an odd-aligned predecessor's last halfword may touch, goes last of all), the next plane, the next block
	// (FlowLayout::ROOM keeps that many bytes free).  Shared-memory stores of one warp are performed in program order.
	__device__ __forceinline__ void flow_store_row(uint32_t at, const uint32_t (&w)[4], uint32_t n)
	{
		const uint32_t odd = at & 1u, s = odd * 8u, R = at + odd;
		const uint32_t v0 = __funnelshift_r(w[0], w[1], s), v1 = __funnelshift_r(w[1], w[2], s), v2 = __funnelshift_r(w[2], w[3], s), v3 = w[3] >> s;
#ifdef STENOS_EMU
		// the emulator runs the lanes of a warp one after the other between two collectives: a rendez-vous after every
		// store gives the stores the order they have on the GPU (called by all 32 lanes)
#define FLOW_ROW_STEP(stmt) \
	do { \
		if (n != 0u) \
			stmt; \
		__syncwarp(); \
	} while (0)
		FLOW_ROW_STEP(sts_u16(R + 14u, v3 >> 16));
		FLOW_ROW_STEP(sts_u16(R + 12u, v3));
		FLOW_ROW_STEP(sts_u16(R + 10u, v2 >> 16));
		FLOW_ROW_STEP(sts_u16(R + 8u, v2));
		FLOW_ROW_STEP(sts_u16(R + 6u, v1 >> 16));
		FLOW_ROW_STEP(sts_u16(R + 4u, v1));
		FLOW_ROW_STEP(sts_u16(R + 2u, v0 >> 16));
		FLOW_ROW_STEP(sts_u16(R, v0));
		FLOW_ROW_STEP(sts_u8(at, w[0]));
#undef FLOW_ROW_STEP
#else
		if (n != 0u) {
			sts_u16(R + 14u, v3 >> 16);
			sts_u16(R + 12u, v3);
			sts_u16(R + 10u, v2 >> 16);
			sts_u16(R + 8u, v2);
			sts_u16(R + 6u, v1 >> 16);
			sts_u16(R + 4u, v1);
			sts_u16(R + 2u, v0 >> 16);
			sts_u16(R, v0);
			sts_u8(at, w[0]);
		}
#endif
	}

	// (lo,hi) << (8 * nbytes) as a 128-bit value OR-ed into w[0..3]; nbytes in 0..10, the value has at most 8 bytes
	__device__ __forceinline__ void flow_or_shifted(uint32_t (&w)[4], uint32_t lo, uint32_t hi, uint32_t nbytes)
	{
		const uint32_t s = (nbytes & 3u) * 8u;
		const uint32_t t0 = lo << s;
		const uint32_t t1 = __funnelshift_l(lo, hi, s);
		const uint32_t t2 = __funnelshift_l(hi, 0u, s);
		const uint32_t k = nbytes >> 2; // 0..2
		w[0] |= k == 0u ? t0 : 0u;
		w[1] |= k == 0u ? t1 : (k == 1u ? t0 : 0u);
		w[2] |= k == 0u ? t2 : (k == 1u ? t1 : t0);
		w[3] |= k == 0u ? 0u : (k == 1u ? t2 : t1);
	}

	// hides from the compiler that x is a power of two: x * y + z stays ONE IMAD (fma pipe) instead of SHF + IADD on
	// the alu pipe, which is the busy one in this kernel
	__device__ __forceinline__ uint32_t flow_opaque(uint32_t x)
	{
#ifndef STENOS_EMU
		asm("" : "+r"(x));
#endif
		return x;
	}

	// bit packing of 16 values (< 2^bits, one per byte of v[0..3]) LSB first: 2 * bits bytes in w (:540-602; two groups
	// of 8 values of `bits` bytes each are one contiguous string of 16 * bits bits).  bits in 1..7.  Every shift by a
	// variable amount is a multiplication by a power of two (IMAD / IMAD.WIDE).
	__device__ __forceinline__ void flow_pack_row(const uint32_t (&v)[4], uint32_t bits, uint32_t (&w)[4])
	{
		const uint32_t mul = flow_opaque(1u << bits), mul2 = flow_opaque(mul * mul), mul4 = flow_opaque(mul2 * mul2);
		uint32_t pk[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			// [v0 | v1 << bits] and [v2 | v3 << bits] in the two 16-bit lanes, then the two lanes joined
			const uint32_t c = __byte_perm(v[j], 0u, 0x4341) * mul + __byte_perm(v[j], 0u, 0x4240);
			pk[j] = __byte_perm(c, 0u, 0x4432) * mul2 + __byte_perm(c, 0u, 0x4410);
		}
		// a group of 8 values = 8 * bits bits: pk[odd] << 4 * bits on top of pk[even]
		const unsigned long long g0 = (unsigned long long)pk[1] * mul4 + pk[0];
		const unsigned long long g1 = (unsigned long long)pk[3] * mul4 + pk[2];
		// group 1 follows group 0 at byte `bits`: g1 << 8 * (bits & 3), placed at word bits >> 2 (0 or 1)
		const uint32_t mb = flow_opaque(1u << ((bits & 3u) * 8u));
		const unsigned long long t = (unsigned long long)(uint32_t)g1 * mb;
		const unsigned long long u = (unsigned long long)(uint32_t)(g1 >> 32) * mb + (uint32_t)(t >> 32);
		const uint32_t t0 = (uint32_t)t, t1 = (uint32_t)u, t2 = (uint32_t)(u >> 32);
		const bool k0 = bits < 4u;
		w[0] = (uint32_t)g0 | (k0 ? t0 : 0u);
		w[1] = (uint32_t)(g0 >> 32) | (k0 ? t1 : t0);
		w[2] = k0 ? t2 : t1;
		w[3] = k0 ? 0u : t2;
	}

	// RLE row payload [mask:2][bytes of src whose mask bit is clear] (:258-293).  nz[j]: 0x80 in every byte of the
	// row that does NOT repeat its predecessor; lut: 256 x u32, byte-compaction selectors (set bits of the index, ascending).
	__device__ __forceinline__ void flow_rle_row(const uint32_t (&src)[4], const uint32_t (&nz)[4], const uint32_t* lut, uint32_t (&w)[4])
	{
		const uint32_t kA = flags_to_mask4(nz[0]) | (flags_to_mask4(nz[1]) << 4);
		const uint32_t kB = flags_to_mask4(nz[2]) | (flags_to_mask4(nz[3]) << 4);
		const uint32_t sA = lut[kA], sB = lut[kB];
		const uint32_t cA = (uint32_t)__popc(kA);
		uint32_t al = __byte_perm(src[0], src[1], sA & 0xFFFFu), ah = __byte_perm(src[0], src[1], sA >> 16);
		const uint32_t bl = __byte_perm(src[2], src[3], sB & 0xFFFFu), bh = __byte_perm(src[2], src[3], sB >> 16);
		// clear A's bytes past cA (the selectors' unused nibbles pick byte 0)
		const uint32_t mb = cA * 8u;
		al = mb >= 32u ? al : (al & ((1u << mb) - 1u));
		ah = mb >= 64u ? ah : (mb <= 32u ? 0u : (ah & ((1u << (mb - 32u)) - 1u)));
		const uint32_t mask = ~(kA | (kB << 8)) & 0xFFFFu; // bit set = the byte repeats
		w[0] = mask | (al << 16);
		w[1] = __funnelshift_l(al, ah, 16);
		w[2] = ah >> 16;
		w[3] = 0u;
		flow_or_shifted(w, bl, bh, 2u + cA);
	}

	// ------------------------------------------------------------------------------------------
	// One byte plane of one block per half-warp: analysis (find_pack_bits_params, block_compress.h:385-535) and
	// emission (encode16x16_generic, :739-806) in one pass.  a[0..3]: the lane's row (16 plane bytes); pvw: byte 3 =
	// the plane byte before the row (0 for row 0, :399).  o: shared-window address of the plane's first byte.
	// Returns the plane's size; kind through kind_out (both uniform over the half-warp).
	// ------------------------------------------------------------------------------------------
	__device__ __forceinline__ uint32_t flow_plane(uint32_t o, const uint32_t (&a)[4], uint32_t pvw, bool half_same, bool emit, int r, uint32_t hsh, const uint32_t* lut,
						      const uint16_t* hlut, uint32_t& kind_out)
	{
		// ---- row statistics (:399-474)
		uint32_t d[4];
		row_deltas(a, pvw, d);
		uint32_t mnv, mxv, mnd, mxd;
		minmax16_hi(a, mnv, mxv);
		minmax16_hi(d, mnd, mxd);
		const uint32_t MN = __vmins2(prmt_sx(mnv, mnd, 0xD591), prmt_sx(mnv, mnd, 0xF7B3)); // [values | deltas] as s16x2
		const uint32_t MX = __vmaxs2(prmt_sx(mxv, mxd, 0xD591), prmt_sx(mxv, mxd, 0xF7B3));
		const uint32_t R = __vsub2(MX, MN);
		// bit widths, type, size and header nibble of the row (:334-352, :420-435, :499-502) from the decision table
		const uint32_t hent = hlut[min(R & 0xFFFFu, FLOW_HLUT_RB - 1u) + min(R >> 16, FLOW_HLUT_RD - 1u) * FLOW_HLUT_RB];
		const bool plain = (hent >> 14) != 0u;
		const uint32_t minv = (plain ? MN : (MN >> 16)) & 0xFFu;
		uint32_t sz = (hent >> 9) & 31u;
		uint32_t h = hent & 15u;
		uint32_t pay = (hent >> 4) & 31u;
		// RLE on the values and on the deltas (:439-474).  Neither can win over a packed row of size <= 3: rs = 2 + (bytes
		// that differ from their predecessor) < 3 means a constant row that continues the previous byte, whose size is 1,
		// and the same holds for the deltas -- planes of constant rows and single steps (the upper bytes of a counter)
		// skip the counting for the whole warp.
		uint32_t nzd[4] = { 0u, 0u, 0u, 0u }, nzx[4] = { 0u, 0u, 0u, 0u };
		bool use_rle = false, use_drle = false;
		if (__any_sync(FULL, sz > 3u)) {
			uint32_t x[4];
			delta_repeats(d, x);
			uint32_t nr = 0, nd = 0;
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				nzd[j] = nonzero_bytes(d[j]);
				nzx[j] = nonzero_bytes(x[j]);
				nr = sad4_acc(nzd[j], 0u, nr);
				nd = sad4_acc(nzx[j], 0u, nd);
			}
			const uint32_t rs = (nr >> 7) + 2u, ds = (nd >> 7) + 2u;
			use_rle = rs < sz;
			sz = min(sz, rs);
			use_drle = ds < sz;
			if (use_drle) {
				h = 6u;
				pay = ds;
			}
			else if (use_rle) {
				h = 7u;
				pay = rs;
			}
		}
		const bool needmin = !(use_rle || use_drle || h == 15u);

		// ---- plane size and kind (:476-490, :1200-1204): one scan carries the payload bytes (bits 0..9) and the
		// stored mins (bits 10..14) of the rows before mine
		uint32_t incl = pay | (needmin ? 1024u : 0u);
#pragma unroll
		for (int dlt = 1; dlt < 16; dlt <<= 1)
			incl = scan_up16_step(incl, dlt, r);
		const uint32_t tot = __shfl_sync(FULL, incl, 15, 16);
		const uint32_t paysum = tot & 1023u, nmins = tot >> 10;
		uint32_t hmp = __shfl_up_sync(FULL, h | (minv << 8), 1, 16); // the previous row's header nibble and min
		if (r == 0)
			hmp = 0u;
		const uint32_t mmb = (__ballot_sync(FULL, minv == (hmp >> 8)) >> hsh) & 0xFFFFu;
		uint32_t total = 8u + paysum + nmins, kind = KIND_NORMAL, minbytes = nmins;
		const uint32_t mcnt = 16u - (uint32_t)__popc(mmb);
		if (mcnt + 2u < nmins) { // mins-RLE (:480-490): 16 - nomin == nmins
			kind = KIND_NORMAL_RLE;
			total -= nmins - (mcnt + 2u);
			minbytes = mcnt + 2u;
		}
		if (total > 256u) { // raw plane (:1200-1204)
			kind = KIND_RAW;
			total = 256u;
		}
		if (half_same) { // :396-418
			kind = KIND_SAME;
			total = 1u;
		}
		kind_out = kind;

		// ---- emission (:739-806)
		uint32_t w[4] = { a[0], a[1], a[2], a[3] };
		uint32_t rowp = o + 16u * (uint32_t)r, n = 0u;
		if (emit) {
			if (kind == KIND_SAME) {
				if (r == 0)
					sts_u8(o, a[0]);
			}
			else {
				n = 16u;
				if (kind != KIND_RAW) {
					if (r & 1) // row headers: two nibbles per byte, row 2i in the low nibble
						sts_u8(o + ((uint32_t)r >> 1), (hmp & 0xFu) | (h << 4));
					if (kind == KIND_NORMAL_RLE) { // mins as [mask:2][values that differ from the previous one]
						if (r == 0) {
							sts_u8(o + 8u, mmb);
							sts_u8(o + 9u, mmb >> 8);
						}
						if (!((mmb >> r) & 1u))
							sts_u8(o + 10u + (uint32_t)__popc(~mmb & ((1u << r) - 1u)), minv);
					}
					else if (needmin)
						sts_u8(o + 7u + (incl >> 10), minv);
					rowp = o + 8u + minbytes + ((incl & 1023u) - pay);
					n = pay;
					if (use_drle)
						flow_rle_row(d, nzx, lut, w);
					else if (use_rle)
						flow_rle_row(a, nzd, lut, w);
					else if (h != 15u && (h & 7u) != 0u) {
						// (value - min) or (delta - min): with both sides biased to unsigned order no byte borrows
						const uint32_t mb = splat(minv) ^ 0x80808080u;
						uint32_t v[4];
#pragma unroll
						for (int j = 0; j < 4; ++j)
							v[j] = (((h & 8u) ? d[j] : a[j]) ^ 0x80808080u) - mb;
						flow_pack_row(v, h & 7u, w);
					}
				}
			}
		}
		flow_store_row(rowp, w, n); // all 32 lanes, n = 0: nothing to store
		return total;
	}

	// ------------------------------------------------------------------------------------------
	// One full 256-element block per half-warp (lane = half * 16 + row).  e: the lane's row, 16 elements = 16 * T
	// contiguous bytes (loaded by the caller, one block ahead); blk: the half's block (global, for the LZ matcher);
	// active: the half has a block this round; sm + out: where its bytes go (MAXB bytes are free there), out32 the
	// same place as a shared-window address.  Returns the encoded size (uniform over the half-warp; 0 when !active).
	// ------------------------------------------------------------------------------------------
	template<int T>
	__device__ __forceinline__ void flow_load_row(const uint8_t* __restrict__ blk, int r, uint32_t (&e)[4 * T])
	{
		const uint4* src = reinterpret_cast<const uint4*>(blk + (size_t)r * 16 * T);
		if ((reinterpret_cast<uintptr_t>(src) & 31u) == 0) {
#pragma unroll
			for (int i = 0; i < T; i += 2) {
				uint4 u, v;
				ld_global_256(src + i, u, v);
				e[4 * i + 0] = u.x, e[4 * i + 1] = u.y, e[4 * i + 2] = u.z, e[4 * i + 3] = u.w;
				e[4 * i + 4] = v.x, e[4 * i + 5] = v.y, e[4 * i + 6] = v.z, e[4 * i + 7] = v.w;
			}
		}
		else {
#pragma unroll
			for (int i = 0; i < T; ++i) {
				const uint4 v = src[i];
				e[4 * i + 0] = v.x, e[4 * i + 1] = v.y, e[4 * i + 2] = v.z, e[4 * i + 3] = v.w;
			}
		}
	}

	template<int T, class Next>
	__device__ __forceinline__ uint32_t flow_encode_block(const uint32_t (&e)[4 * T], const uint8_t* __restrict__ blk, bool active, uint8_t* sm, uint32_t out, uint32_t out32,
							     uint32_t* lz_scratch, const uint32_t* lut, const uint16_t* hlut, int lane, Next&& row_consumed)
	{
		constexpr uint32_t HS = (T + 1) / 2;
		constexpr int NE = 4 * T;
		const int hb = lane >> 4, r = lane & 15;
		const uint32_t hsh = 16u * (uint32_t)hb;

		// ---- all-same planes (:396-418), on the untransposed words: OR of (element ^ first element of the block)
		constexpr int NG = T == 8 ? 2 : 1; // groups of (up to) four planes
		uint32_t X[NG], XA[NG], F[NG], PV[NG];
		if constexpr (T == 2) {
			const uint32_t f = __shfl_sync(FULL, e[0], 0, 16);
			const uint32_t f2 = __byte_perm(f, f, 0x1010);
			uint32_t xx = e[0] ^ f2;
#pragma unroll
			for (int i = 1; i < NE; ++i)
				xx |= e[i] ^ f2;
			X[0] = (xx | (xx >> 16)) & 0xFFFFu;
			F[0] = f;
			uint32_t pv = __shfl_up_sync(FULL, e[NE - 1] >> 16, 1, 16); // the last element of the previous row
			PV[0] = r == 0 ? 0u : pv;
		}
		else {
#pragma unroll
			for (int g = 0; g < NG; ++g) {
				const uint32_t f = __shfl_sync(FULL, e[g], 0, 16);
				uint32_t xx = e[g] ^ f;
#pragma unroll
				for (int i = 1; i < 16; ++i)
					xx |= e[NG * i + g] ^ f;
				X[g] = xx;
				F[g] = f;
				uint32_t pv = __shfl_up_sync(FULL, e[NG * 15 + g], 1, 16);
				PV[g] = r == 0 ? 0u : pv;
			}
		}
#pragma unroll
		for (int g = 0; g < NG; ++g)
			XA[g] = __reduce_or_sync(FULL, X[g]);
#ifndef STENOS_EMU
#pragma unroll
		for (int g = 0; g < NG; ++g)
			asm("" : "+r"(X[g])); // keeps X in its register (the compiler recomputed the 16-LOP3 chain for every plane instead)
#endif
		row_consumed(); // every word of the row has been read (XA depends on all of them): the input buffer may be refilled

		// both blocks of the warp are one value each (runs): [kinds = 0][T plane bytes], nothing to analyse
		bool all_same = XA[0] == 0u;
		if constexpr (NG == 2)
			all_same = all_same && XA[1] == 0u;
		if (all_same) {
			if (active && r == 0) {
#pragma unroll
				for (uint32_t i = 0; i < HS; ++i)
					sts_u8(out32 + i, 0u);
#pragma unroll
				for (uint32_t p = 0; p < (uint32_t)T; ++p)
					sts_u8(out32 + HS + p, F[p >> 2] >> (8u * (p & 3u)));
			}
			return active ? HS + (uint32_t)T : 0u;
		}

		uint32_t pos = HS, kinds = 0;
#pragma unroll
		for (int g = 0; g < NG; ++g) {
			constexpr int NP = T == 2 ? 2 : 4;
#pragma unroll 1
			for (int pl = 0; pl < NP; ++pl) {
				const uint32_t sh8 = 8u * (uint32_t)pl;
				const int p = 4 * g + pl;
				if (((XA[g] >> sh8) & 0xFFu) == 0u) {
					// the plane is one value in both blocks of the warp: kind 0, one byte
					if (active && r == 0)
						sts_u8(out32 + pos, F[g] >> sh8);
					pos += 1u;
					continue;
				}
				uint32_t a[4];
				if constexpr (T == 2) {
					const uint32_t sel = 0x6420u + 0x1111u * (uint32_t)pl;
#pragma unroll
					for (int j = 0; j < 4; ++j)
						a[j] = __byte_perm(e[2 * j], e[2 * j + 1], sel);
				}
				else {
					const uint32_t sel = 0x0040u + 0x0011u * (uint32_t)pl; // [x_p, y_p, ., .]
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						const uint32_t t0 = __byte_perm(e[NG * (4 * j) + g], e[NG * (4 * j + 1) + g], sel);
						const uint32_t t1 = __byte_perm(e[NG * (4 * j + 2) + g], e[NG * (4 * j + 3) + g], sel);
						a[j] = __byte_perm(t0, t1, 0x5410);
					}
				}
				const uint32_t pvw = PV[g] << (24u - sh8); // byte 3 = the plane byte before my row
				const uint32_t sm_ = __ballot_sync(FULL, ((X[g] >> sh8) & 0xFFu) == 0u);
				const bool half_same = ((sm_ >> hsh) & 0xFFFFu) == 0xFFFFu;
				uint32_t kind;
				const uint32_t psz = flow_plane(out32 + pos, a, pvw, half_same, active, r, hsh, lut, hlut, kind);
				kinds |= kind << (4 * p);
				pos += psz;
			}
		}
		if (active && r == 0) {
#pragma unroll
			for (uint32_t i = 0; i < HS; ++i)
				sts_u8(out32 + i, kinds >> (8 * i));
		}
		uint32_t size = active ? pos : 0u;

		// ---- LZ attempt (block_compress.h:1210-1223) on blocks whose plane coding ratio is < 3: the whole warp
		// works on one block at a time; a successful stream (never longer than the plane coding) overwrites it
		if constexpr ((T % 4) == 0) {
			const uint32_t full = pos - HS;
			const uint32_t want = __ballot_sync(FULL, active && full * 3u > (uint32_t)T * 256u);
			if (want) {
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					if (!((want >> (16 * hh)) & 1u))
						continue;
					const uint8_t* gs = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)blk, 16 * hh));
					const uint32_t oo = __shfl_sync(FULL, out, 16 * hh);
					const uint32_t fmax = __shfl_sync(FULL, full, 16 * hh);
					uint32_t w[2 * T];
					load_lane_words<T>(gs, lane, w);
					__syncwarp();
					const uint32_t lr = lz_encode_block<T>(gs, w, sm + oo + 1u, fmax, lz_scratch, lane);
					if (lr) {
						if (lane == 0)
							sm[oo] = (uint8_t)MARK_LZ;
						if (hb == hh)
							size = lr + 1u;
					}
					__syncwarp();
				}
			}
		}
		return size;
	}

	// ------------------------------------------------------------------------------------------
	// copies by one half-warp (r = lane & 15); both halves of a warp call them together with their own arguments
	// ------------------------------------------------------------------------------------------

	// n bytes shared (offset s_off, any alignment) -> global (any alignment), 16-byte stores on the destination's alignment
	__device__ __forceinline__ void half_copy_from_smem(uint8_t* __restrict__ dst, const uint8_t* sm, uint32_t s_off, uint32_t n, int r)
	{
		if (n <= 16u) { // the pieces of constant regions are a few bytes
			if ((uint32_t)r < n)
				dst[r] = sm[s_off + r];
			return;
		}
		const uint32_t head = min((uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u), n);
		if ((uint32_t)r < head)
			dst[r] = sm[s_off + r];
		const uint32_t body = (n - head) >> 4;
		const uint32_t s = s_off + head;
		const uint32_t sh = (s & 3u) * 8u;
		const uint32_t* sw = reinterpret_cast<const uint32_t*>(sm + (s & ~3u));
		uint4* d = reinterpret_cast<uint4*>(dst + head);
		for (uint32_t j = r; j < body; j += 16) {
			const uint32_t a = sw[4 * j], b = sw[4 * j + 1], c = sw[4 * j + 2], e = sw[4 * j + 3], f = sw[4 * j + 4];
			d[j] = make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh), __funnelshift_r(c, e, sh), __funnelshift_r(e, f, sh));
		}
		const uint32_t done = head + (body << 4);
		if (done + (uint32_t)r < n)
			dst[done + r] = sm[s_off + done + r];
	}

	// n bytes global (16-byte aligned) -> global (any alignment)
	__device__ __forceinline__ void half_copy_global(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, int r)
	{
		const uint32_t head = min((uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u), n);
		if ((uint32_t)r < head)
			dst[r] = src[r];
		const uint32_t body = (n - head) >> 4;
		uint4* d = reinterpret_cast<uint4*>(dst + head);
		if (head == 0) {
			const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll 4
			for (uint32_t j = r; j < body; j += 16)
				d[j] = s4[j];
		}
		else {
			const uint32_t sh = (head & 3u) * 8u;
			const uint32_t* sw = reinterpret_cast<const uint32_t*>(src + (head & ~3u));
#pragma unroll 2
			for (uint32_t j = r; j < body; j += 16) {
				const uint32_t a = sw[4 * j], b = sw[4 * j + 1], c = sw[4 * j + 2], e = sw[4 * j + 3];
				const uint32_t f = sh ? sw[4 * j + 4] : 0u;
				d[j] = make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh), __funnelshift_r(c, e, sh), __funnelshift_r(e, f, sh));
			}
		}
		const uint32_t done = head + (body << 4);
		if (done + (uint32_t)r < n)
			dst[done + r] = src[done + r];
	}

#ifndef STENOS_EMU
	// test support (stenos_b200_test_occupy): a CTA that holds its SM's shared memory for a while
	__global__ void occupy_kernel(unsigned long long ns)
	{
		extern __shared__ uint8_t occupy_smem[];
		occupy_smem[threadIdx.x] = 1;
		unsigned long long t0, t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
		do {
			__nanosleep(2000);
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		} while (t1 - t0 < ns);
	}
#endif

	__device__ __forceinline__ uint32_t ld_vol_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
	__device__ __forceinline__ void st_vol_u32(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

	// ------------------------------------------------------------------------------------------
	// the kernel: NT / 32 encoder warps + 1 placer warp
	// ------------------------------------------------------------------------------------------
	template<int T, int NT, int KB>
	__global__ void __launch_bounds__(NT + 32, 1) encode_flow_kernel(EncodeParams P)
	{
		using L = FlowLayout<T, NT, KB>;
		STENOS_DYN_SMEM(uint8_t, smem);
		unsigned long long* sbq = reinterpret_cast<unsigned long long*>(smem + L::CTL_OFF); // (q + 1) << 32 | superblock of the CTA's q-th turn
		uint32_t* lut = reinterpret_cast<uint32_t*>(smem + L::LUT_OFF);
		uint16_t* hlut = reinterpret_cast<uint16_t*>(smem + L::HLUT_OFF);
		const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
		const int hb = lane >> 4, r = lane & 15;
		const uint64_t first_off = P.header_len ? (uint64_t)P.header_len : P.base_offset;
		auto slot_of = [&](uint32_t q) { return reinterpret_cast<FlowSlot*>(smem + L::SLOT_OFF + (q % FLOW_NS) * L::SLOT_BYTES); };
		auto sizes_of = [&](FlowSlot* s) { return reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s) + 64); };
		auto offs_of = [&](FlowSlot* s) { return reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s) + 64 + L::NPC * 4u); };
		// waits decided by lane 0 and broadcast: lanes polling a word on their own may see different values, and what
		// follows a wait are warp collectives
		auto wait_u32 = [&](const uint32_t* p, uint32_t want) {
			while (__shfl_sync(FULL, ld_vol_u32(p), 0) != want)
				FLOW_LONG_WAIT();
		};

		for (uint32_t i = tid; i < L::LUT_OFF / 4u; i += NT + 32)
			reinterpret_cast<uint32_t*>(smem)[i] = 0u;
		for (uint32_t m = tid; m < 256u; m += NT + 32) {
			// byte-compaction selectors: nibble k = position of the k-th set bit of m
			uint32_t sel = 0, k = 0;
			for (uint32_t b = 0; b < 8u; ++b)
				if ((m >> b) & 1u)
					sel |= b << (4u * k++);
			lut[m] = sel;
		}
		for (uint32_t i = tid; i < FLOW_HLUT_N; i += NT + 32)
			hlut[i] = (uint16_t)flow_hlut_entry(i % FLOW_HLUT_RB, i / FLOW_HLUT_RB);
		__syncthreads();
		if (tid == 0) {
			for (uint32_t i = 0; i < (uint32_t)FLOW_NS; ++i)
				slot_of(i)->owner = i + 1u;
			const uint32_t s0 = atomicAdd(P.ticket + 1, 1u); // [1]: [0] belongs to encode_frame_kernel
			sbq[0] = (1ull << 32) | s0;
		}
		__syncthreads();

		if (warp == (int)L::NW) {
			// ================================ the placer ================================
			for (uint32_t q = 0;; ++q) {
				unsigned long long tk;
				while (((tk = __shfl_sync(FULL, *reinterpret_cast<volatile unsigned long long*>(&sbq[q % FLOW_SBQ]), 0)) >> 32) != q + 1u)
					STENOS_SPIN_WAIT();
				const uint32_t s = (uint32_t)tk;
				if (s >= P.n_stream)
					break;
				FlowSlot* slot = slot_of(q);
				wait_u32(&slot->agg, q + 1u); // all pieces are there: sized, scanned and announced by the last encoder warp
				__threadfence_block();
				const uint8_t* in = P.src + (uint64_t)s * P.sb_bytes;
				const uint32_t in_bytes = (uint32_t)min((uint64_t)P.sb_bytes, P.bytes - (uint64_t)s * P.sb_bytes);
				const uint32_t nfull = in_bytes / L::BLOCK;
				const uint32_t rem = in_bytes - nfull * L::BLOCK;
				const uint32_t csize = ld_vol_u32(&slot->csize), tail_off = ld_vol_u32(&slot->tail_off), tail_sz = ld_vol_u32(&slot->tail_sz);
				const bool copy = csize > in_bytes; // stenos.cpp:609-610
				const uint32_t len = copy ? in_bytes : csize;
				const uint32_t out_size = 4u + len;

				// frame offset: decoupled look-back across superblocks (other CTAs), 32 predecessors per probe
				uint64_t fexcl = 0;
				{
					long long top = (long long)s - 1;
					while (top >= 0) {
						const long long idx = top - lane;
						const unsigned long long v = idx >= 0 ? ld_volatile_u64(&P.state[idx]) : LB_INCLUSIVE;
						const uint32_t inc = __ballot_sync(FULL, (v & LB_INCLUSIVE) != 0ull);
						const uint32_t inv = __ballot_sync(FULL, (v >> 62) == 0ull);
						const int fi = inc ? (__ffs((int)inc) - 1) : 32;
						const uint32_t upto = fi >= 31 ? 0xFFFFFFFFu : ((2u << fi) - 1u);
						if (inv & upto) {
							STENOS_SPIN_WAIT();
							continue;
						}
						unsigned long long val = ((upto >> lane) & 1u) ? (v & LB_VALUE) : 0ull;
#pragma unroll
						for (int dd = 16; dd >= 1; dd >>= 1)
							val += __shfl_xor_sync(FULL, val, dd);
						fexcl += val;
						if (fi < 32)
							break;
						top -= 32;
					}
				}
				const uint64_t base = first_off + fexcl;
				uint32_t mode = copy ? 1u : 0u;
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
					mode = 2u;
				}
				else {
					uint8_t* o = P.dst + base;
					if (lane == 0) {
						o[0] = (uint8_t)(copy ? CODE_COPY : CODE_BLOCK);
						o[1] = (uint8_t)len;
						o[2] = (uint8_t)(len >> 8);
						o[3] = (uint8_t)(len >> 16);
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
					if (rem) {
						if (copy)
							warp_copy_global(o + 4u + (size_t)nfull * L::BLOCK, in + (size_t)nfull * L::BLOCK, rem, lane);
						else
							warp_copy_bytes(o + 4u + tail_off, smem + L::TAIL_OFF, tail_sz, lane);
					}
				}
				__syncwarp();
				if (lane == 0) {
					*reinterpret_cast<volatile unsigned long long*>(&slot->base) = base;
					st_vol_u32(&slot->mode, mode);
					st_vol_u32(&slot->sb, s);
					__threadfence_block();
					st_vol_u32(&slot->ready, q + 1u);
				}
				__syncwarp();
			}
			return;
		}

		// ================================ the encoder warps ================================
		// A TASK is a pair of adjacent pieces of one superblock (one piece per half-warp).  Warps take tasks from a counter
		// in shared memory, in order: t -> the CTA's turn q = t / NW, pair j = t % NW.  (Static pieces per warp were measured
		// first: a warp that runs a little slower -- more warps on its scheduler, a spinning neighbour -- fell turns behind,
		// the fast ones ran ahead until their rings were full, the superblock tickets they had taken sat unencoded and
		// every placement behind them waited: 40 % of the warps' time went into waiting for ring space.)
		const uint32_t half = 2u * (uint32_t)warp + (uint32_t)hb;
		uint32_t* lz_scratch = reinterpret_cast<uint32_t*>(smem + L::LZ_OFF + L::LZ_STRIDE * (L::HAS_LZ ? warp : 0));
		const uint32_t stage = L::STAGE_OFF + half * L::REG; // my half-warp's staging ring
		uint32_t* pinfo = reinterpret_cast<uint32_t*>(smem + L::PINFO_OFF) + half * (FLOW_PQ * 4u);
		uint32_t* pmeta = reinterpret_cast<uint32_t*>(smem + L::PMETA_OFF) + (uint32_t)warp * FLOW_PQ;
		uint32_t* task_ctr = reinterpret_cast<uint32_t*>(smem + L::TASK_OFF);
		constexpr uint32_t K = KB; // blocks per piece
		uint32_t pq_head = 0, pq_tail = 0;     // my warp's pieces waiting for their placement: entries [pq_head, pq_tail) mod FLOW_PQ
		uint32_t pq_ring = 0;                  // entries [pq_head, pq_ring) were moved to their spill slots, [pq_ring, pq_tail) sit in the rings
		uint8_t* const spill = P.spill + ((size_t)blockIdx.x * L::NH + half) * (size_t)FLOW_PQ * L::SPILL_SLOT; // my half's spill slots
		uint32_t pos = 0, vcur = 0, vtail = 0; // my half's ring: write position, and monotonic counters of bytes claimed / released
		// bytes of the ring that are used (tests shrink it to the legal minimum so that pieces spill to HBM all the time)
		uint32_t REG = P.ring_cap ? min(L::REG, max(L::REG_MIN, P.ring_cap)) : L::REG;
#ifndef STENOS_EMU
		asm("" : "+r"(REG)); // one register, not three instructions at every use
#endif
		const uint32_t stage32 = smem_addr32(smem) + stage;
		const uint32_t in32 = smem_addr32(smem) + L::IN_OFF + (uint32_t)warp * L::IN_STRIDE; // my warp's input rows (FLOW_STAGING 2)
		// The row of the block my half encodes NEXT is already on its way from HBM while the current block is encoded
		// (with the loads issued at the top of a block, 12 % of the warps' time was the wait for them).
		uint32_t e[4 * T];
#pragma unroll
		for (int i = 0; i < 4 * T; ++i)
			e[i] = 0u;
		bool have = false;

		// A warp takes at most TASK_LIMIT tasks of one turn (its share + 1).  Together with "never wait in the middle of a
		// task" this is what keeps every wait of this kernel safe: a warp that holds a task only ever waits for a turn OLDER
		// than the task's, which completes without it (see the waits below).
#ifndef FLOW_TASK_SLACK
#define FLOW_TASK_SLACK 1
#endif
		constexpr uint32_t TASK_LIMIT = (L::NTASK + L::NW - 1u) / L::NW + FLOW_TASK_SLACK;
		static_assert(TASK_LIMIT * L::NW >= L::NTASK + L::NW && (uint32_t)FLOW_PQ > TASK_LIMIT, "task limit / waiting list too small");
		constexpr uint32_t NO_TASK = 0xFFFFFFFFu;
		uint32_t my_turn = 0xFFFFFFFFu, my_cnt = 0;
		auto try_take = [&]() -> uint32_t {
			const uint32_t peek = __shfl_sync(FULL, ld_vol_u32(task_ctr), 0) / L::NTASK;
			if (peek == my_turn && my_cnt >= TASK_LIMIT)
				return NO_TASK; // the rest of this turn is for the other warps
			uint32_t t = 0;
			if (lane == 0)
				t = atomicAdd(task_ctr, 1u);
			t = __shfl_sync(FULL, t, 0);
			const uint32_t q = t / L::NTASK;
			my_cnt = q == my_turn ? my_cnt + 1u : 1u;
			my_turn = q;
			return t;
		};
		// is the oldest waiting piece placed?  (decided by lane 0 for the warp)
		auto head_ready = [&]() {
			const uint32_t qo = ld_vol_u32(pmeta + pq_head % FLOW_PQ) / L::NTASK;
			return __shfl_sync(FULL, ld_vol_u32(&slot_of(qo)->ready), 0) == qo + 1u;
		};
		// copies my half's oldest waiting piece to the frame and releases its ring space (blocks until it is placed)
		auto drain_one = [&]() {
			const uint32_t to_ = ld_vol_u32(pmeta + pq_head % FLOW_PQ);
			const uint32_t qo = to_ / L::NTASK, jo = to_ % L::NTASK;
			FlowSlot* sl = slot_of(qo);
			wait_u32(&sl->ready, qo + 1u);
			__threadfence_block();
			const uint32_t mode = ld_vol_u32(&sl->mode);
			const unsigned long long base = *reinterpret_cast<volatile unsigned long long*>(&sl->base);
			const uint32_t pi = 2u * jo + (uint32_t)hb;
			const uint32_t* pf = pinfo + (pq_head % FLOW_PQ) * 4u;
			const uint32_t posA = ld_vol_u32(pf + 0), lenA = ld_vol_u32(pf + 1), lenB = ld_vol_u32(pf + 2), vend = ld_vol_u32(pf + 3);
			const bool spilled = pq_head != pq_ring; // uniform over the warp: both halves' pieces of a task move together
			if (mode == 0u) {
				uint8_t* to = P.dst + base + 4u + ld_vol_u32(offs_of(sl) + pi);
				if (spilled)
					half_copy_global(to, spill + (size_t)(pq_head % FLOW_PQ) * L::SPILL_SLOT, lenA + lenB, r);
				else {
					half_copy_from_smem(to, smem, stage + posA, lenA, r);
					if (lenB)
						half_copy_from_smem(to + lenA, smem, stage, lenB, r);
				}
			}
			else if (mode == 1u) {
				// COPY superblock (stenos.cpp:609-610): my blocks of the input instead
				const uint32_t so = ld_vol_u32(&sl->sb);
				const uint32_t in_bytes = (uint32_t)min((uint64_t)P.sb_bytes, P.bytes - (uint64_t)so * P.sb_bytes);
				const uint32_t nfull = in_bytes / L::BLOCK;
				const uint32_t b0 = pi * K;
				const uint32_t cnt = b0 < nfull ? min(K, nfull - b0) : 0u;
				if (cnt)
					half_copy_global(P.dst + base + 4u + (uint64_t)b0 * L::BLOCK, P.src + (uint64_t)so * P.sb_bytes + (uint64_t)b0 * L::BLOCK, cnt * L::BLOCK, r);
			}
			if (!spilled) {
				vtail = vend;
				pq_ring = pq_head + 1u;
			}
			__syncwarp();
			if (lane == 0) {
				const uint32_t old = atomicAdd(&sl->consumed, 1u);
				if (old == L::NTASK - 1u) {
					// every task of this superblock has left: the slot goes to turn qo + FLOW_NS
					sl->arrived = 0u;
					sl->consumed = 0u;
					__threadfence_block();
					st_vol_u32(&sl->owner, qo + (uint32_t)FLOW_NS + 1u);
				}
			}
			__syncwarp();
			++pq_head;
		};

		// moves the oldest piece still in the rings to its spill slot in HBM and releases its ring space
		auto spill_one = [&]() {
			const uint32_t* pf = pinfo + (pq_ring % FLOW_PQ) * 4u;
			const uint32_t posA = ld_vol_u32(pf + 0), lenA = ld_vol_u32(pf + 1), lenB = ld_vol_u32(pf + 2), vend = ld_vol_u32(pf + 3);
			uint8_t* to = spill + (size_t)(pq_ring % FLOW_PQ) * L::SPILL_SLOT;
			half_copy_from_smem(to, smem, stage + posA, lenA, r);
			if (lenB)
				half_copy_from_smem(to + lenA, smem, stage, lenB, r);
			vtail = vend;
			__syncwarp();
			++pq_ring;
		};
		// room for one more block in both rings: the oldest pieces move to their spill slots -- never a wait: the superblock
		// they belong to may be waiting for the task I am working on.  (Placed pieces leave at the top of every task, so a
		// ring that fills up in the middle of one holds pieces that are not placed yet.)
		// my half's piece in the making: where it starts in the ring, its bytes before / after the ring's wrap, and the
		// bytes of it that already moved to its spill slot (pieces are not bounded by the ring: K worst-case blocks may be
		// several times its size, typical data needs a fraction of it)
		uint32_t posA = 0, lenA = 0, lenB = 0, sp_len = 0;
		bool wrapped = false;
		// appends what the ring holds of the piece in the making to its spill slot (the slot of the entry it will get)
		auto flush_current = [&]() {
			uint8_t* to = spill + (size_t)(pq_tail % FLOW_PQ) * L::SPILL_SLOT + sp_len;
			half_copy_from_smem(to, smem, stage + posA, lenA, r);
			if (lenB)
				half_copy_from_smem(to + lenA, smem, stage, lenB, r);
			sp_len += lenA + lenB;
			lenA = 0;
			lenB = 0;
			wrapped = false;
			posA = pos;
			vtail = vcur; // nothing older is left in the ring (see make_room)
			__syncwarp();
		};
		auto make_room = [&](bool need) {
			while (__any_sync(FULL, need && vcur + L::ROOM - vtail > REG)) {
				if (pq_ring != pq_tail)
					spill_one();
				else
					flush_current(); // the piece in the making is what fills the rings
			}
		};

		uint32_t t = try_take();
		bool finishing = false;
		for (;;) {
			// ---- everything that may wait happens here, at the top of a task, and pieces leave for the frame here only
			uint32_t q = 0, j = 0, s = 0;
			FlowSlot* slot = nullptr;
			bool must_drain = false;
			for (;;) {
				if (pq_head != pq_tail && (must_drain || head_ready())) {
					drain_one(); // (waits for the placement when it must)
					must_drain = false;
					continue;
				}
				if (finishing) {
					if (pq_head == pq_tail)
						break;
					must_drain = true;
					continue;
				}
				if (t == NO_TASK) {
					// I hold no task: waiting is safe
					t = try_take();
					if (t == NO_TASK)
						FLOW_LONG_WAIT();
					continue;
				}
				q = t / L::NTASK, j = t % L::NTASK;
				const unsigned long long tk = __shfl_sync(FULL, *reinterpret_cast<volatile unsigned long long*>(&sbq[q % FLOW_SBQ]), 0);
				if ((uint32_t)(tk >> 32) != q + 1u) {
					STENOS_SPIN_HINT(); // the CTA's q-th superblock is being asked for
					continue;
				}
				s = (uint32_t)tk;
				if (s >= P.n_stream) {
					// Superblocks are handed out in increasing order: this CTA has no further work.  Warps that already hold a
					// task of the next turn must see that too (nobody will ask for that turn's superblock any more).
					if (lane == 0)
						*reinterpret_cast<volatile unsigned long long*>(&sbq[(q + 1u) % FLOW_SBQ]) = ((unsigned long long)(q + 2u) << 32) | 0xFFFFFFFFull;
					finishing = true;
					continue;
				}
				if (pq_tail - pq_head >= (uint32_t)FLOW_PQ) {
					must_drain = true; // no entry left for this piece: the oldest one belongs to an older turn (FLOW_PQ > TASK_LIMIT), safe to wait for
					continue;
				}
				slot = slot_of(q);
				if (__shfl_sync(FULL, ld_vol_u32(&slot->owner), 0) == q + 1u)
					break;
				// the slot is still held by turn q - FLOW_NS: somebody (maybe me) has not copied that superblock out yet
				FLOW_LONG_WAIT();
			}
			if (finishing)
				break;
			if (j == L::NTASK / 2u && lane == 0) {
				// The CTA's next superblock is asked for half a turn ahead: early enough that its number is there when the first
				// task of the next turn starts, late enough that the ticket does not sit unencoded for long (every superblock
				// handed out and not yet encoded holds up the placement of all later ones).
				const uint32_t nx = atomicAdd(P.ticket + 1, 1u);
				*reinterpret_cast<volatile unsigned long long*>(&sbq[(q + 1u) % FLOW_SBQ]) = ((unsigned long long)(q + 2u) << 32) | nx;
			}
			const uint8_t* in = P.src + (uint64_t)s * P.sb_bytes;
			const uint32_t in_bytes = (uint32_t)min((uint64_t)P.sb_bytes, P.bytes - (uint64_t)s * P.sb_bytes);
			const uint32_t nfull = in_bytes / L::BLOCK;
			const uint32_t pi = 2u * j + (uint32_t)hb; // my half's piece
			const uint32_t b0 = pi * K;
			const uint32_t cnt = b0 < nfull ? min(K, nfull - b0) : 0u;
			const uint8_t* blk = in + (uint64_t)b0 * L::BLOCK;

			posA = pos, lenA = 0, lenB = 0, sp_len = 0;
			wrapped = false;
			uint32_t tnext = NO_TASK; // my warp's next task, taken during the last block of this one
			const uint32_t kmax = max(1u, __reduce_max_sync(FULL, cnt));
			for (uint32_t it = 0; it < kmax; ++it) {
				const bool active = it < cnt;
				if (active && pos + L::ROOM > REG) {
					// a block never wraps: the piece continues at the start of the ring
					vcur += REG - pos;
					pos = 0;
					wrapped = true;
				}
				make_room(active);
				if (it + 1u == kmax)
					tnext = try_take();
				const uint8_t* myblk = blk + (size_t)it * L::BLOCK;
				// my half's next block: the next one of the piece, or the first one of my piece of the next task
				const uint8_t* nb = nullptr;
				if (it + 1u < cnt)
					nb = myblk + L::BLOCK;
				else if (it + 1u == kmax && tnext != NO_TASK) {
					const uint32_t q2 = tnext / L::NTASK;
					const unsigned long long t2 = *reinterpret_cast<volatile unsigned long long*>(&sbq[q2 % FLOW_SBQ]);
					const uint32_t s2 = (uint32_t)t2;
					if ((uint32_t)(t2 >> 32) == q2 + 1u && s2 < P.n_stream) {
						const uint32_t nfull2 = (uint32_t)min((uint64_t)P.sb_bytes, P.bytes - (uint64_t)s2 * P.sb_bytes) / L::BLOCK;
						const uint32_t b2 = (2u * (tnext % L::NTASK) + (uint32_t)hb) * K;
						if (b2 < nfull2)
							nb = P.src + (uint64_t)s2 * P.sb_bytes + (uint64_t)b2 * L::BLOCK;
					}
				}
				const bool upd = it + 1u < cnt || it + 1u == kmax; // (an idle round of a half with fewer blocks keeps what was fetched)
#if FLOW_STAGING == 2
				if (active && !have)
					flow_fetch_row<T>(in32, myblk + (size_t)r * 16 * T, lane);
				cp_async_commit(); // (wait_group only waits for committed copies)
				cp_async_wait_all();
				flow_read_row<T>(in32, lane, e);
				const uint32_t sz = flow_encode_block<T>(e, myblk, active, smem, stage + pos, stage32 + pos, lz_scratch, lut, hlut, lane, [&]() {
					if (upd) {
						have = nb != nullptr;
						if (have)
							flow_fetch_row<T>(in32, nb + (size_t)r * 16 * T, lane);
					}
					cp_async_commit();
				});
#elif FLOW_STAGING == 1
				if (active && !have)
					flow_load_row<T>(myblk, r, e);
				uint32_t cur[4 * T];
#pragma unroll
				for (int i = 0; i < 4 * T; ++i)
					cur[i] = e[i];
				if (upd) {
					have = nb != nullptr;
					if (have)
						flow_load_row<T>(nb, r, e);
				}
				const uint32_t sz = flow_encode_block<T>(cur, myblk, active, smem, stage + pos, stage32 + pos, lz_scratch, lut, hlut, lane, []() {});
#else
				if (nb != nullptr && ((uint32_t)r * 16u * T) % 128u == 0u)
					prefetch_l2(nb + (size_t)r * 16 * T);
				if (active)
					flow_load_row<T>(myblk, r, e);
				const uint32_t sz = flow_encode_block<T>(e, myblk, active, smem, stage + pos, stage32 + pos, lz_scratch, lut, hlut, lane, []() {});
#endif
				pos += sz;
				vcur += sz;
				if (wrapped)
					lenB += sz;
				else
					lenA += sz;
			}

			// ---- publish my piece; the placer puts the superblock in the frame once all pieces are there
			const bool in_spill = __any_sync(FULL, sp_len != 0u); // (both halves' pieces of a task move together)
			if (in_spill) {
				flush_current(); // the rest of it: the whole piece sits in its spill slot
				lenA = sp_len;
			}
			if (r == 0) {
				uint32_t* pf = pinfo + (pq_tail % FLOW_PQ) * 4u;
				st_vol_u32(pf + 0, posA);
				st_vol_u32(pf + 1, lenA);
				st_vol_u32(pf + 2, lenB);
				st_vol_u32(pf + 3, vcur);
				st_vol_u32(sizes_of(slot) + pi, lenA + lenB);
			}
			if (lane == 0)
				st_vol_u32(pmeta + pq_tail % FLOW_PQ, t);
			++pq_tail;
			if (in_spill)
				pq_ring = pq_tail; // [pq_head, pq_ring) are the spilled entries: this one joins them (pq_ring was pq_tail already)
			__syncwarp();
			__threadfence_block();
			uint32_t na = 0;
			if (lane == 0)
				na = atomicAdd(&slot->arrived, 1u);
			na = __shfl_sync(FULL, na, 0);
			if (na == L::NTASK - 1u) {
				// Last task of the superblock: exclusive scan of the piece sizes, and the superblock's size goes out as the
				// AGGREGATE word of the look-back right away -- successors must never wait for this CTA's placer, which may
				// itself be waiting for a predecessor (measured: with the placer publishing it, placers spun ~100 % of the time).
				__threadfence_block();
				uint32_t* sizes = sizes_of(slot);
				uint32_t* offs = offs_of(slot);
				constexpr uint32_t PER = (L::NPC + 31u) / 32u;
				uint32_t mine[PER], sum = 0;
#pragma unroll
				for (uint32_t k = 0; k < PER; ++k) {
					const uint32_t i = (uint32_t)lane * PER + k;
					mine[k] = i < L::NPC ? ld_vol_u32(sizes + i) : 0u;
					sum += mine[k];
				}
				uint32_t incl = sum;
#pragma unroll
				for (int dlt = 1; dlt < 32; dlt <<= 1) {
					const uint32_t tt = __shfl_up_sync(FULL, incl, dlt);
					if (lane >= dlt)
						incl += tt;
				}
				uint32_t run = incl - sum;
#pragma unroll
				for (uint32_t k = 0; k < PER; ++k) {
					const uint32_t i = (uint32_t)lane * PER + k;
					if (i < L::NPC)
						st_vol_u32(offs + i, run);
					run += mine[k];
				}
				uint32_t csize = __shfl_sync(FULL, incl, 31);
				// the frame's partial tail block (at most one superblock of the launch has one)
				const uint32_t rem = in_bytes - nfull * L::BLOCK;
				const uint32_t tail_off = csize;
				uint32_t tail_sz = 0;
				if (rem) {
					bool e2 = false;
					tail_sz = encode_partial_block<T, false>(in + (size_t)nfull * L::BLOCK, rem, smem + L::TAIL_OFF, lane, 0xFFFFFFFFu, e2);
					__syncwarp();
					csize += tail_sz;
				}
				if (lane == 0) {
					const uint32_t out_size = 4u + (csize > in_bytes ? in_bytes : csize);
					st_volatile_u64(&P.state[s], LB_AGGREGATE | (unsigned long long)out_size);
					st_vol_u32(&slot->csize, csize);
					st_vol_u32(&slot->tail_off, tail_off);
					st_vol_u32(&slot->tail_sz, tail_sz);
					__threadfence_block();
					st_vol_u32(&slot->agg, q + 1u);
				}
				__syncwarp();
			}
			t = tnext;
		}
#if FLOW_STAGING == 2
		cp_async_wait_all();
#endif
	}
}
