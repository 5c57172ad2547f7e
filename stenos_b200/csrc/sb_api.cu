// sb_api.cu -- host side of the C ABI declared in include/stenos_b200.h.
//
// Mirrors the reference's public API for the level-1 path (stenos/internal/stenos.cpp) on top of
// the kernels of sb_kernels.cuh / sb_filters.cuh.  Frame logic that is not arithmetic on the data
// (header layout, superblock sizing, error codes) follows stenos.cpp line by line (cited inline);
// all data-path work runs on the GPU.  The only host-side data work is the Zstd coding of a final
// superblock shorter than 128 bytes (stenos.cpp:435-437), which the reference delegates to libzstd
// as well -- here through dlopen("libzstd.so.1").  Frames written at level >= 2 are decoded by the hybrid path at the
// end of this file: host Zstd, device filters and block decoder (decompress_hybrid).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <thread>
#include <vector>
#include <dlfcn.h>

#include "sb_kernels.cuh"
#include "sb_stream.cuh"
#include "sb_flow.cuh"
#include "sb_decode_rows.cuh"
#include "sb_filters.cuh"
#include "../../include/stenos_b200.h"


namespace
{
	std::atomic<unsigned long long> g_launches{ 0 };

	inline bool is_err(size_t r) { return r >= STENOS_LAST_ERROR_CODE; }

	// ---- libzstd, only for the < 128 byte tail superblock ---------------------------------------
	typedef size_t (*zstd_compress_fn)(void*, size_t, const void*, size_t, int);
	typedef size_t (*zstd_decompress_fn)(void*, size_t, const void*, size_t);
	typedef unsigned (*zstd_iserror_fn)(size_t);
	zstd_compress_fn p_zstd_compress = nullptr;
	zstd_decompress_fn p_zstd_decompress = nullptr;
	zstd_iserror_fn p_zstd_iserror = nullptr;
	typedef int (*zstd_maxclevel_fn)(void);
	zstd_maxclevel_fn p_zstd_maxclevel = nullptr;
	bool load_zstd()
	{
		static std::atomic<int> state{ 0 }; // 0 untried, 1 ok, 2 missing
		int s = state.load();
		if (s == 0) {
			void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
			if (h) {
				p_zstd_compress = (zstd_compress_fn)dlsym(h, "ZSTD_compress");
				p_zstd_decompress = (zstd_decompress_fn)dlsym(h, "ZSTD_decompress");
				p_zstd_iserror = (zstd_iserror_fn)dlsym(h, "ZSTD_isError");
				p_zstd_maxclevel = (zstd_maxclevel_fn)dlsym(h, "ZSTD_maxCLevel");
			}
			s = (p_zstd_compress && p_zstd_decompress && p_zstd_iserror) ? 1 : 2;
			state.store(s);
		}
		return s == 1;
	}

	// ---- small RAII-free device buffer that only grows -----------------------------------------
	struct DevBuf
	{
		uint8_t* p = nullptr;
		size_t cap = 0;
		bool reserve(size_t n)
		{
			if (n <= cap)
				return true;
			if (p)
				cudaFree(p);
			p = nullptr;
			cap = 0;
			size_t want = n + (n >> 3) + 256;
			if (cudaMalloc((void**)&p, want) != cudaSuccess) {
				cudaGetLastError();
				p = nullptr;
				if (cudaMalloc((void**)&p, n + 64) != cudaSuccess) {
					cudaGetLastError();
					p = nullptr;
					return false;
				}
				want = n + 64;
			}
			cap = want;
			return true;
		}
		void release()
		{
			if (p)
				cudaFree(p);
			p = nullptr;
			cap = 0;
		}
	};

	bool is_device_ptr(const void* p)
	{
		if (!p)
			return false;
		cudaPointerAttributes a;
		if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
			cudaGetLastError();
			return false;
		}
		return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
	}

	// Pinned (cudaHostAlloc / cudaHostRegister) host memory is addressable from kernels: returns its device alias, or
	// nullptr for pageable host memory and device pointers.
	uint8_t* pinned_alias(const void* p)
	{
		if (!p)
			return nullptr;
		cudaPointerAttributes a;
		if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
			cudaGetLastError();
			return nullptr;
		}
		return a.type == cudaMemoryTypeHost ? static_cast<uint8_t*>(a.devicePointer) : nullptr;
	}
	// experiments: STENOS_B200_ZERO_COPY bit 0: kernels read pinned host input in place, bit 1: write pinned host output in place
	int zero_copy_mode()
	{
		const char* e = getenv("STENOS_B200_ZERO_COPY");
		return e ? atoi(e) : 0;
	}

	// super_block_size(), stenos.cpp:71-76
	size_t default_superblock(size_t T)
	{
		const size_t bs = T * 256;
		return bs > STENOS_BLOCK_SIZE ? bs : (STENOS_BLOCK_SIZE / bs) * bs;
	}
}

struct stenos_context_s
{
	// parameters (stenos.cpp:94-98)
	int level = 1;
	int threads = 1;
	uint64_t max_ns = 0;
	size_t custom_shift = STENOS_NO_BLOCK_SHIFT;
	// device additions
	int device = -1;
	cudaStream_t user_stream = nullptr;
	bool has_user_stream = false;
	cudaStream_t own_stream = nullptr;
	bool own_stream_made = false;
	// prepared state
	size_t superblock = 0;
	int shift = 0;
	// scratch
	DevBuf in, out, ctl, idx, scan, dtk, blk, spill, hyb, hoffs;
	bool serial_index = false; // tests: force the serial header walk
	bool legacy_encoder = false; // tests: every superblock through encode_frame_kernel
	bool legacy_decoder = false;
	bool env_read = false;
	bool index_ran = false;    // a parallel frame index was enqueued (its verdict is in scan.p)
	unsigned long long* host_result = nullptr; // pinned, 4 words
	// host <-> device pipeline of stenos_compress_generic on host buffers: copy streams, per-chunk events and results
	static constexpr int PIPE_MAX = 64;
	cudaStream_t pipe_in = nullptr, pipe_out = nullptr;
	cudaEvent_t pipe_ev[2 * PIPE_MAX] = {};
	unsigned long long* pipe_host = nullptr; // pinned, 2 words per chunk
	static constexpr size_t SMALL_BYTES = 65536; // host <-> host calls up to this size return [result][bytes] in one copy
	uint8_t* small_host = nullptr;               // pinned, 16 + SMALL_BYTES + 64
	bool small_init()
	{
		if (small_host)
			return true;
		if (cudaMallocHost((void**)&small_host, 16 + SMALL_BYTES + 64) != cudaSuccess) {
			cudaGetLastError();
			small_host = nullptr;
			return false;
		}
		return true;
	}
	unsigned long long* pipe_offs = nullptr; // pinned: superblock offsets of a host frame (pipelined decompress)
	size_t pipe_offs_cap = 0;
	bool pipe_offs_reserve(size_t n)
	{
		if (n <= pipe_offs_cap)
			return true;
		if (pipe_offs)
			cudaFreeHost(pipe_offs);
		pipe_offs = nullptr;
		pipe_offs_cap = 0;
		const size_t want = n + (n >> 2) + 64;
		if (cudaMallocHost((void**)&pipe_offs, want * 8) != cudaSuccess) {
			cudaGetLastError();
			pipe_offs = nullptr;
			return false;
		}
		pipe_offs_cap = want;
		return true;
	}
	DevBuf pipe_res;                         // 2 words per chunk
	bool pipe_ready = false;
	bool pipe_init()
	{
		if (pipe_ready)
			return true;
		if (cudaStreamCreateWithFlags(&pipe_in, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&pipe_out, cudaStreamNonBlocking) != cudaSuccess ||
		    cudaMallocHost((void**)&pipe_host, PIPE_MAX * 16) != cudaSuccess || !pipe_res.reserve(PIPE_MAX * 16)) {
			cudaGetLastError();
			return false;
		}
		for (int i = 0; i < 2 * PIPE_MAX; ++i)
			if (cudaEventCreateWithFlags(&pipe_ev[i], cudaEventDisableTiming) != cudaSuccess) {
				cudaGetLastError();
				return false;
			}
		pipe_ready = true;
		return true;
	}
	int sm_count = 0;

	cudaStream_t stream()
	{
		if (has_user_stream)
			return user_stream;
		if (!own_stream_made) {
			if (cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking) != cudaSuccess)
				own_stream = nullptr;
			own_stream_made = true;
		}
		return own_stream;
	}
	bool activate()
	{
		if (!env_read) {
			// test hook: STENOS_B200_LEGACY_ENCODER=1 sends every superblock through encode_frame_kernel
			const char* e = getenv("STENOS_B200_LEGACY_ENCODER");
			legacy_encoder = e && e[0] == '1';
			e = getenv("STENOS_B200_LEGACY_DECODER"); // likewise: one warp per superblock, lane per half row
			legacy_decoder = e && e[0] == '1';
			env_read = true;
		}
		if (device >= 0 && cudaSetDevice(device) != cudaSuccess) {
			cudaGetLastError();
			return false;
		}
		if (!sm_count) {
			int dev = 0;
			cudaGetDevice(&dev);
			cudaDeviceProp prop;
			if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
				cudaGetLastError();
				return false;
			}
			sm_count = prop.multiProcessorCount;
		}
		if (!host_result) {
			if (cudaMallocHost((void**)&host_result, 64) != cudaSuccess) {
				cudaGetLastError();
				host_result = nullptr;
				return false;
			}
		}
		return true;
	}
	// prepare(), stenos.cpp:115-185, without the time-limited branch
	size_t prepare(size_t T, size_t bytes)
	{
		if (T == 0 || T >= STENOS_MAX_BYTESOFTYPE)
			return STENOS_ERROR_INVALID_BYTESOFTYPE;
		const size_t block = T * 256;
		size_t sb;
		shift = 0;
		if (custom_shift != STENOS_NO_BLOCK_SHIFT) {
			sb = block << custom_shift;
			shift = 255;
		}
		else {
			sb = default_superblock(T);
			if (bytes > sb) {
				shift = level ? (level - 1) / 2 : 0;
				sb <<= (size_t)shift;
			}
		}
		if (sb < block || sb >= STENOS_MAX_BLOCK_BYTES)
			return STENOS_ERROR_INVALID_PARAMETER;
		superblock = sb;
		return 0;
	}
	void free_all()
	{
		in.release();
		out.release();
		ctl.release();
		idx.release();
		scan.release();
		dtk.release();
		blk.release();
		spill.release();
		hyb.release();
		hoffs.release();
		if (host_result)
			cudaFreeHost(host_result);
		host_result = nullptr;
		if (pipe_host)
			cudaFreeHost(pipe_host);
		pipe_host = nullptr;
		if (small_host)
			cudaFreeHost(small_host);
		small_host = nullptr;
		if (pipe_offs)
			cudaFreeHost(pipe_offs);
		pipe_offs = nullptr;
		pipe_offs_cap = 0;
		for (int i = 0; i < 2 * PIPE_MAX; ++i)
			if (pipe_ev[i]) {
				cudaEventDestroy(pipe_ev[i]);
				pipe_ev[i] = nullptr;
			}
		if (pipe_in)
			cudaStreamDestroy(pipe_in);
		if (pipe_out)
			cudaStreamDestroy(pipe_out);
		pipe_in = pipe_out = nullptr;
		pipe_res.release();
		pipe_ready = false;
		if (own_stream_made && own_stream)
			cudaStreamDestroy(own_stream);
		own_stream = nullptr;
		own_stream_made = false;
	}
};

namespace
{
	using namespace sb;

	bool supported_T(size_t T) { return T == 2 || T == 4 || T == 8; } // the fast kernels, the device-resident API, filters, buckets
	// other element sizes (first slice): 3, 5, 6 and 7 bytes -- no LZ at these sizes (block_compress.h:1210: T % 4 == 0
	// only) -- through the generic kernels (encode_frame_kernel / decode_frame_kernel: one warp per block, lane per half row)
	bool generic_T(size_t T) { return T == 3 || T == 5 || T == 6 || T == 7; }
	bool device_T(size_t T) { return supported_T(T) || generic_T(T); }

#ifndef ENCODE_THREADS_2
#define ENCODE_THREADS_2 1024
#endif
#ifndef ENCODE_THREADS_4
#define ENCODE_THREADS_4 512
#endif
#ifndef ENCODE_THREADS_8
#define ENCODE_THREADS_8 256
#endif
#ifndef STREAM_THREADS_2
#define STREAM_THREADS_2 1024
#endif
#ifndef STREAM_THREADS_4
#define STREAM_THREADS_4 640
#endif
#ifndef STREAM_THREADS_8
#define STREAM_THREADS_8 384
#endif

	// ---- launchers ------------------------------------------------------------------------------
	template<int T, int NT>
	size_t launch_encode_T(stenos_context* ctx, const EncodeParams& P)
	{
		const uint32_t smem = EncodeLayout<T>::smem_bytes(NT / 32);
		if (cudaFuncSetAttribute((const void*)encode_frame_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
			cudaGetLastError();
			return STENOS_ERROR_ALLOC;
		}
		// persistent CTAs: one per SM (shared-memory slots allow exactly one), superblocks by ticket
		const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(P.n_sb - P.first_sb, ctx->sm_count));
		auto kern = encode_frame_kernel<T, NT>;
		STENOS_LAUNCH(kern, dim3(grid), dim3(NT), smem, ctx->stream(), P);
		++g_launches;
		return cudaGetLastError() == cudaSuccess ? 0 : STENOS_ERROR_UNDEFINED;
	}
	// bucket mode of encode_frame_kernel (EncodeParams::bucket_stride): buckets of one 256-element block run as small
	// CTAs (one warp, slots for one block: 8 CTAs per SM and more), larger buckets with the frame layout
	template<int T>
	size_t launch_buckets_T(stenos_context* ctx, const EncodeParams& P)
	{
		const bool one_block = P.sb_bytes <= (uint32_t)T * 256u;
		if (one_block) {
			// two buckets per warp: the lane-per-row pair encoder, the room-exact one only where the room could matter
			const uint32_t smem = BUCKET_WARPS * (2u * EncodeLayout<T, 1>::STRIDE + EncodeLayout<T, 1>::LZ_STRIDE);
			int per_sm = 4;
#ifndef STENOS_EMU
			if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_bucket_pairs_kernel<T>, BUCKET_WARPS * 32, smem) != cudaSuccess || per_sm < 1) {
				cudaGetLastError();
				per_sm = 4;
			}
#endif
			const long long need = ((long long)(P.n_sb + 1) / 2 + BUCKET_WARPS - 1) / BUCKET_WARPS;
			const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(need, (long long)ctx->sm_count * per_sm));
			STENOS_LAUNCH(encode_bucket_pairs_kernel<T>, dim3(grid), dim3(BUCKET_WARPS * 32), smem, ctx->stream(), P);
		}
		else {
			constexpr int NT = 256;
			const uint32_t smem = EncodeLayout<T>::smem_bytes(NT / 32);
			if (cudaFuncSetAttribute((const void*)encode_frame_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
				cudaGetLastError();
				return STENOS_ERROR_ALLOC;
			}
			const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(P.n_sb, ctx->sm_count));
			auto kern = encode_frame_kernel<T, NT>;
			STENOS_LAUNCH(kern, dim3(grid), dim3(NT), smem, ctx->stream(), P);
		}
		++g_launches;
		return cudaGetLastError() == cudaSuccess ? 0 : STENOS_ERROR_UNDEFINED;
	}
#ifndef FLOW_THREADS_2
#define FLOW_THREADS_2 608
#endif
#ifndef FLOW_THREADS_4
#define FLOW_THREADS_4 576
#endif
#ifndef FLOW_THREADS_8
#define FLOW_THREADS_8 352
#endif
	// blocks per piece (a task = two pieces)
#ifndef FLOW_KB_2
#define FLOW_KB_2 8
#endif
#ifndef FLOW_KB_4
#define FLOW_KB_4 4
#endif
#ifndef FLOW_KB_8
#define FLOW_KB_8 4
#endif
	// encode_flow_kernel (sb_flow.cuh): the default fast path
	template<int T, int NT, int KB>
	size_t launch_flow_T(stenos_context* ctx, const EncodeParams& P)
	{
		const uint32_t smem = FlowLayout<T, NT, KB>::smem_bytes();
		if (cudaFuncSetAttribute((const void*)encode_flow_kernel<T, NT, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
			cudaGetLastError();
			return STENOS_ERROR_ALLOC;
		}
		// persistent CTAs, one per SM (the staging rings take the SM's shared memory); superblocks by ticket
		const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(P.n_stream, ctx->sm_count));
		if (!ctx->spill.reserve(FlowLayout<T, NT, KB>::spill_bytes(grid)))
			return STENOS_ERROR_ALLOC;
		EncodeParams Q = P;
		Q.spill = ctx->spill.p;
		{
			const char* e = getenv("STENOS_B200_FLOW_RING"); // tests: staging ring bytes (clamped to the legal minimum): forces the spill path
			Q.ring_cap = e ? (uint32_t)strtoul(e, nullptr, 10) : 0u;
		}
		auto kern = encode_flow_kernel<T, NT, KB>;
#ifdef STENOS_EMU
		if (getenv("STENOS_EMU_TRACE"))
			fprintf(stderr, "encode_flow_kernel<%d,%d> grid %u n_stream %u n_sb %u\n", T, NT, grid, P.n_stream, P.n_sb);
#endif
		STENOS_LAUNCH(kern, dim3(grid), dim3(NT + 32), smem, ctx->stream(), Q); // + the placer warp
		++g_launches;
		return cudaGetLastError() == cudaSuccess ? 0 : STENOS_ERROR_UNDEFINED;
	}
	bool use_stream_v1()
	{
		static const bool v = [] { const char* e = getenv("STENOS_B200_ENCODER"); return e && e[0] == '1'; }(); // experiments: the round-1 pipeline
		return v;
	}
	template<int T, int NT>
	size_t launch_stream_T(stenos_context* ctx, const EncodeParams& P)
	{
		const uint32_t smem = StreamLayout<T, NT>::smem_bytes();
		if (cudaFuncSetAttribute((const void*)encode_stream_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
			cudaGetLastError();
			return STENOS_ERROR_ALLOC;
		}
		// persistent CTAs, one per SM (the shared-memory ring takes the SM); CTA c owns superblocks c, c + grid, ...
		const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(P.n_stream, ctx->sm_count));
		auto kern = encode_stream_kernel<T, NT>;
		STENOS_LAUNCH(kern, dim3(grid), dim3(NT), smem, ctx->stream(), P);
		++g_launches;
		return cudaGetLastError() == cudaSuccess ? 0 : STENOS_ERROR_UNDEFINED;
	}
	// Superblocks [0, n_stream) go through the barrier-free pipeline (encode_stream_kernel); the ones whose
	// dst room could change an encoder decision -- and level 0 -- through
	// encode_frame_kernel, which continues the same look-back chain.
	size_t launch_encode(stenos_context* ctx, size_t T, EncodeParams P)
	{
		uint32_t n_stream = 0;
		if (P.level != 0 && !ctx->legacy_encoder && supported_T(T)) {
			const uint64_t block = T * 256, hs = (T + 1) / 2;
			const uint64_t need = (P.sb_bytes / block) * (block + hs) + 8 * T + 32 + 8 * T + block; // worst stream + slack (+ a partial block's worst case)
			const uint64_t first_off = P.header_len ? (uint64_t)P.header_len : P.base_offset;
			if (P.dst_size >= first_off + 4 + need) {
				const uint64_t k = (P.dst_size - first_off - 4 - need) / (4ull + P.sb_bytes) + 1; // superblocks 0..k-1 have room even after k-1 COPY superblocks
				n_stream = (uint32_t)std::min<uint64_t>(k, P.n_sb);
			}
		}
		P.n_stream = n_stream;
		P.first_sb = n_stream;
		size_t r = 0;
		if (n_stream) {
			const bool v1 = use_stream_v1();
			switch (T) {
				case 2: r = v1 ? launch_stream_T<2, STREAM_THREADS_2>(ctx, P) : launch_flow_T<2, FLOW_THREADS_2, FLOW_KB_2>(ctx, P); break;
				case 4: r = v1 ? launch_stream_T<4, STREAM_THREADS_4>(ctx, P) : launch_flow_T<4, FLOW_THREADS_4, FLOW_KB_4>(ctx, P); break;
				case 8: r = v1 ? launch_stream_T<8, STREAM_THREADS_8>(ctx, P) : launch_flow_T<8, FLOW_THREADS_8, FLOW_KB_8>(ctx, P); break;
				default: return STENOS_ERROR_INVALID_PARAMETER;
			}
			if (is_err(r) || n_stream == P.n_sb)
				return r;
		}
		switch (T) {
			case 2: return launch_encode_T<2, ENCODE_THREADS_2>(ctx, P);
			case 4: return launch_encode_T<4, ENCODE_THREADS_4>(ctx, P);
			case 8: return launch_encode_T<8, ENCODE_THREADS_8>(ctx, P);
			case 3: return launch_encode_T<3, 512>(ctx, P);
			case 5: return launch_encode_T<5, 512>(ctx, P);
			case 6: return launch_encode_T<6, 512>(ctx, P);
			case 7: return launch_encode_T<7, 512>(ctx, P);
		}
		return STENOS_ERROR_INVALID_PARAMETER;
	}
	template<int T>
	size_t launch_decode_T(stenos_context* ctx, const DecodeParams& P)
	{
		if (ctx->legacy_decoder) {
			const unsigned grid = (P.n_sb + DECODE_WARPS - 1) / DECODE_WARPS;
			STENOS_LAUNCH(decode_frame_kernel<T>, dim3(grid), dim3(DECODE_WARPS * 32), DECODE_WARPS * 512, ctx->stream(), P);
		}
		else {
			if (!ctx->dtk.reserve(16))
				return STENOS_ERROR_ALLOC;
			cudaMemsetAsync(ctx->dtk.p, 0, 16, ctx->stream());
			DecodeParams Q = P;
			Q.ticket = reinterpret_cast<unsigned long long*>(ctx->dtk.p);
			int per_sm = 1;
			{
				// experiment knob: fewer resident CTAs per SM than the occupancy allows
				static const int cap = [] { const char* e = getenv("STENOS_B200_DECODE_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
#ifndef STENOS_EMU
				if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_pairs_kernel<T>, DECODE2_WARPS * 32, DECODE2_WARPS * (512 + 128)) != cudaSuccess || per_sm < 1) {
					cudaGetLastError();
					per_sm = 1;
				}
#endif
				if (cap > 0)
					per_sm = std::min(per_sm, cap);
			}
			// two superblocks per warp (one per half-warp, one lane per row); persistent warps, work by ticket
			const unsigned warps = (P.n_sb + 1) / 2;
			const unsigned need = (warps + DECODE2_WARPS - 1) / DECODE2_WARPS;
			const unsigned grid = std::min<unsigned>(need, (unsigned)(ctx->sm_count * per_sm));
			STENOS_LAUNCH(decode_pairs_kernel<T>, dim3(grid), dim3(DECODE2_WARPS * 32), DECODE2_WARPS * (512 + 128), ctx->stream(), Q);
		}
		++g_launches;
		return cudaGetLastError() == cudaSuccess ? 0 : STENOS_ERROR_UNDEFINED;
	}
	template<int T>
	size_t launch_decode_generic_T(stenos_context* ctx, const DecodeParams& P)
	{
		const unsigned grid = (P.n_sb + DECODE_WARPS - 1) / DECODE_WARPS;
		STENOS_LAUNCH(decode_frame_kernel<T>, dim3(grid), dim3(DECODE_WARPS * 32), DECODE_WARPS * 512, ctx->stream(), P);
		++g_launches;
		return cudaGetLastError() == cudaSuccess ? 0 : STENOS_ERROR_UNDEFINED;
	}
	size_t launch_decode(stenos_context* ctx, size_t T, const DecodeParams& P)
	{
		if (!P.n_sb)
			return 0;
		switch (T) {
			case 2: return launch_decode_T<2>(ctx, P);
			case 4: return launch_decode_T<4>(ctx, P);
			case 8: return launch_decode_T<8>(ctx, P);
			case 3: return launch_decode_generic_T<3>(ctx, P);
			case 5: return launch_decode_generic_T<5>(ctx, P);
			case 6: return launch_decode_generic_T<6>(ctx, P);
			case 7: return launch_decode_generic_T<7>(ctx, P);
		}
		return STENOS_ERROR_INVALID_PARAMETER;
	}
	template<int T>
	size_t launch_gather_T(stenos_context* ctx, const GatherParams& P)
	{
		if (ctx->legacy_decoder) {
			const unsigned grid = (P.n + DECODE_WARPS - 1) / DECODE_WARPS;
			STENOS_LAUNCH(gather_decode_kernel<T>, dim3(grid), dim3(DECODE_WARPS * 32), DECODE_WARPS * 512, ctx->stream(), P);
		}
		else {
			// persistent warps, bucket pairs round robin (the kernel pipelines its random accesses across iterations)
			const unsigned warps = (P.n + 1) / 2;
			const unsigned need = (warps + DECODE2_WARPS - 1) / DECODE2_WARPS;
			int per_sm = 1;
#ifndef STENOS_EMU
			if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_pairs_kernel<T>, DECODE2_WARPS * 32, DECODE2_WARPS * (512 + 128)) != cudaSuccess || per_sm < 1) {
				cudaGetLastError();
				per_sm = 1;
			}
#endif
			const unsigned grid = std::min<unsigned>(need, (unsigned)(ctx->sm_count * per_sm));
			STENOS_LAUNCH(gather_pairs_kernel<T>, dim3(grid), dim3(DECODE2_WARPS * 32), DECODE2_WARPS * (512 + 128), ctx->stream(), P);
		}
		++g_launches;
		return cudaGetLastError() == cudaSuccess ? 0 : STENOS_ERROR_UNDEFINED;
	}

	// control block in device memory: [result:2 x u64][ticket u32 + pad][state: n_sb x u64]
	struct Control
	{
		unsigned long long* result;
		uint32_t* ticket;
		unsigned long long* state;
	};
	bool make_control(stenos_context* ctx, size_t n_sb, Control& c, bool zero_result)
	{
		const size_t bytes = 32 + n_sb * 8;
		if (!ctx->ctl.reserve(bytes))
			return false;
		c.result = reinterpret_cast<unsigned long long*>(ctx->ctl.p);
		c.ticket = reinterpret_cast<uint32_t*>(ctx->ctl.p + 16);
		c.state = reinterpret_cast<unsigned long long*>(ctx->ctl.p + 32);
		if (zero_result)
			cudaMemsetAsync(ctx->ctl.p, 0, bytes, ctx->stream());
		else
			cudaMemsetAsync(ctx->ctl.p + 16, 0, bytes - 16, ctx->stream());
		return true;
	}


	// Superblock index of a device resident frame: parallel re-synchronising scan for large frames,
	// the serial walk for small ones.  d_offs: [n_sb + 1]; d_result[1] receives error bits.
	size_t enqueue_index(stenos_context* ctx, const uint8_t* d_src, size_t size, size_t first, size_t n_sb, size_t sb, unsigned long long* d_offs,
			     unsigned long long* d_result)
	{
		cudaStream_t st = ctx->stream();
		// segments of about half a superblock: every warp re-synchronises within one superblock and walks a few
		// hops; at most 65536 segments so the single-CTA merge stays short
		size_t seg = std::min<size_t>(std::max<size_t>(sb / 2, 4096), 65536);
		if (const char* e = getenv("STENOS_B200_INDEX_SEG")) // experiments
			seg = std::max<size_t>((size_t)strtoull(e, nullptr, 10), 4096);
		const size_t span = size > first ? size - first : 0;
		if ((span + seg - 1) / seg > 65536)
			seg = (span + 65535) / 65536;
		seg = (seg + 15) & ~(size_t)15;
		const size_t n_seg = (span + seg - 1) / seg;
		if (n_sb < 64 || n_seg < 8 || ctx->serial_index) {
			IndexParams I;
			I.src = d_src;
			I.src_size = size;
			I.first = first;
			I.n_sb = (uint32_t)n_sb;
			I.sb_offsets = d_offs;
			I.result = d_result;
			STENOS_LAUNCH(frame_index_kernel, dim3(1), dim3(32), 0, st, I);
			++g_launches;
			return 0;
		}
		const size_t bytes = 16 + n_seg * (8 + 8 + 4 + 4);
		if (!ctx->scan.reserve(bytes))
			return STENOS_ERROR_ALLOC;
		FastIndexParams F;
		F.src = d_src;
		F.src_size = size;
		F.first = first;
		F.n_sb = (uint32_t)n_sb;
		F.max_csize = (uint32_t)std::min<size_t>(sb, 0xFFFFFFu);
		F.seg_bytes = (uint32_t)seg;
		F.n_seg = (uint32_t)n_seg;
		F.ok = reinterpret_cast<uint32_t*>(ctx->scan.p);
		F.seg_start = reinterpret_cast<unsigned long long*>(ctx->scan.p + 16);
		F.seg_end = F.seg_start + n_seg;
		F.seg_count = reinterpret_cast<uint32_t*>(F.seg_end + n_seg);
		F.seg_base = F.seg_count + n_seg;
		F.sb_offsets = d_offs;
		F.result = d_result;
		ctx->index_ran = true;
		STENOS_LAUNCH(index_scan_kernel, dim3((unsigned)((n_seg + INDEX_WARPS - 1) / INDEX_WARPS)), dim3(INDEX_WARPS * 32), 0, st, F);
		STENOS_LAUNCH(index_merge_kernel, dim3(1), dim3(1024), 256, st, F);
		STENOS_LAUNCH(index_fill_kernel, dim3((unsigned)((n_seg + 127) / 128)), dim3(128), 0, st, F);
		g_launches += 3;
		return cudaGetLastError() == cudaSuccess ? 0 : STENOS_ERROR_UNDEFINED;
	}

	size_t map_device_error(unsigned long long bits)
	{
		if (bits & DEV_ERR_DST_OVERFLOW)
			return STENOS_ERROR_DST_OVERFLOW;
		if (bits & DEV_ERR_SRC_OVERFLOW)
			return STENOS_ERROR_SRC_OVERFLOW;
		if (bits & DEV_ERR_INVALID_INPUT)
			return STENOS_ERROR_INVALID_INPUT;
		return 0;
	}

	// Enqueues the encoder for `bytes` (whole superblocks, each >= 128 bytes or level 0) of a frame
	// or segment.  d_result: device [2] words, zeroed here.
	size_t enqueue_encode(stenos_context* ctx, const uint8_t* d_src, size_t T, size_t bytes, uint8_t* d_dst, size_t dst_size, size_t sb, uint32_t header_len,
			      uint32_t shift_byte, uint64_t frame_bytes, int level, unsigned long long* d_result, unsigned long long* d_sb_offsets)
	{
		const size_t n_sb = (bytes + sb - 1) / sb;
		if (!n_sb)
			return 0;
		if (n_sb > 0xFFFFFFF0ull)
			return STENOS_ERROR_INVALID_PARAMETER;
		Control c;
		if (!make_control(ctx, n_sb, c, d_result == nullptr))
			return STENOS_ERROR_ALLOC;
		if (d_result)
			cudaMemsetAsync(d_result, 0, 16, ctx->stream());
		EncodeParams P;
		P.src = d_src;
		P.bytes = bytes;
		P.dst = d_dst;
		P.dst_size = dst_size;
		P.sb_bytes = (uint32_t)sb;
		P.n_sb = (uint32_t)n_sb;
		P.header_len = header_len;
		P.shift_byte = shift_byte;
		P.frame_bytes = frame_bytes;
		P.level = level;
		P.state = c.state;
		P.ticket = c.ticket;
		P.result = d_result ? d_result : c.result;
		P.sb_offsets = d_sb_offsets;
		P.base_offset = 0;
		P.spill = nullptr;
		P.ring_cap = 0;
		return launch_encode(ctx, T, P);
	}

	bool write_frame_header(uint8_t* h, int shift, uint64_t bytes, size_t sb, bool custom)
	{
		// stenos.cpp:862-874
		h[0] = (uint8_t)shift;
		for (int i = 0; i < 7; ++i)
			h[1 + i] = (uint8_t)(bytes >> (8 * i));
		if (custom)
			for (int i = 0; i < 4; ++i)
				h[8 + i] = (uint8_t)(sb >> (8 * i));
		return true;
	}

	// The final superblock of a level-1 frame when it is shorter than 128 bytes (stenos.cpp:435-437, :658-674): Zstd level 1
	// through the host's libzstd, stored raw if that does not shrink it.  tail, out: host memory; room: bytes left in dst.
	// Returns the bytes written ([code][len:3][payload]) or an error code.
	size_t encode_host_tail(const uint8_t* tail, size_t last_bytes, uint8_t* out, size_t room)
	{
		uint8_t enc[4 + 256];
		if (room < 4)
			return STENOS_ERROR_DST_OVERFLOW;
		if (!load_zstd())
			return STENOS_ERROR_ZSTD_INTERNAL;
		size_t len = 0;
		// zstd_compress_with_context(dst + 4, dst_size - 4, src, bytes, 0) -> zstd level 1 (zstd_wrapper.h:49-56)
		const size_t zr = p_zstd_compress(enc + 4, std::min<size_t>(room - 4, 256), tail, last_bytes, 1);
		if (!p_zstd_iserror(zr) && zr <= last_bytes) {
			len = zr;
			enc[0] = (uint8_t)CODE_ZSTD;
		}
		else { // stenos.cpp:668-669 -> MEMCPY
			if (room < last_bytes + 4)
				return STENOS_ERROR_DST_OVERFLOW;
			enc[0] = (uint8_t)CODE_COPY;
			memcpy(enc + 4, tail, last_bytes);
			len = last_bytes;
		}
		enc[1] = (uint8_t)len;
		enc[2] = (uint8_t)(len >> 8);
		enc[3] = (uint8_t)(len >> 16);
		memcpy(out, enc, len + 4);
		return len + 4;
	}

	// Common body of stenos_compress_generic and stenos_private_compress_block.
	// frame = true : [frame header] + superblocks;   frame = false: one bare superblock
	size_t compress_impl(stenos_context* ctx, const void* src_, size_t T, size_t bytes, void* dst_, size_t dst_size, bool frame, size_t sb)
	{
		if (!ctx->activate())
			return STENOS_ERROR_ALLOC;
		if (ctx->max_ns != 0)
			return STENOS_ERROR_INVALID_PARAMETER; // time limited compression is wall-clock driven: out of scope
		const int level = ctx->level;
		if (level > 1 || !device_T(T))
			return STENOS_ERROR_INVALID_PARAMETER; // no CPU fallback: levels >= 2 need Zstd
		if (sb > STENOS_BLOCK_SIZE)
			return STENOS_ERROR_INVALID_PARAMETER; // shared-memory slots hold at most 128 KiB superblocks
		const uint8_t* src = static_cast<const uint8_t*>(src_);
		uint8_t* dst = static_cast<uint8_t*>(dst_);
		const bool custom = ctx->custom_shift != STENOS_NO_BLOCK_SHIFT;
		const uint32_t header_len = frame ? (custom ? 12u : 8u) : 0u;
		cudaStream_t st = ctx->stream();

		if (frame) {
			// stenos.cpp:862-878
			if (dst_size < 8 || (custom && dst_size < 12))
				return STENOS_ERROR_DST_OVERFLOW;
			if (bytes == 0) {
				uint8_t h[12];
				write_frame_header(h, ctx->shift, 0, sb, custom);
				if (is_device_ptr(dst)) {
					cudaMemcpyAsync(dst, h, header_len, cudaMemcpyHostToDevice, st);
					cudaStreamSynchronize(st);
				}
				else
					memcpy(dst, h, header_len);
				return header_len;
			}
		}
		else if (dst_size < 4)
			return STENOS_ERROR_DST_OVERFLOW; // stenos.cpp:427-429

		const bool src_dev = is_device_ptr(src), dst_dev = is_device_ptr(dst);
		// the tail superblock shorter than 128 bytes goes through libzstd on the host (stenos.cpp:435-437)
		const size_t n_sb = bytes ? (bytes + sb - 1) / sb : (frame ? 0 : 1);
		const size_t last_bytes = bytes - (n_sb ? (n_sb - 1) * sb : 0);
		const bool host_tail = level >= 1 && n_sb && last_bytes < 128 && bytes != 0;
		const size_t dev_bytes = host_tail ? bytes - last_bytes : bytes;

		// ---- host -> host through the device, pipelined: the input travels in chunks of whole superblocks; chunk i is
		// encoded (as an independent segment, into its own worst-case region) while chunk i + 1 is still on the bus, and
		// its stream leaves for the caller's buffer while later chunks are encoded.  Taken only when the caller's dst
		// leaves every superblock ample room (then the reference's room checks are inert, SURVEY.md appendix C2, and
		// the concatenated segments ARE the frame); otherwise the single launch below keeps the exact room arithmetic.
		// (STENOS_B200_PIPELINE_CHUNK = chunk bytes, tests: small frames take the path too; 0 = off)
		const char* pipe_env = getenv("STENOS_B200_PIPELINE_CHUNK");
		const size_t pipe_chunk = pipe_env ? (size_t)strtoull(pipe_env, nullptr, 10) : ((size_t)32 << 20);
		if (frame && !src_dev && !dst_dev && level >= 1 && pipe_chunk && dev_bytes >= 2 * pipe_chunk) {
			const uint64_t block = T * 256, hs = (T + 1) / 2;
			const uint64_t need = (sb / block) * (block + hs) + 8 * T + 32 + 8 * T + block;
			const size_t dev_sb = (dev_bytes + sb - 1) / sb;
			if (dst_size >= (uint64_t)header_len + 4 + need + (uint64_t)(dev_sb - 1) * (4 + sb) && ctx->pipe_init()) {
				size_t chunk = std::max<size_t>(pipe_chunk, (dev_bytes + stenos_context_s::PIPE_MAX - 1) / stenos_context_s::PIPE_MAX);
				chunk = (chunk + sb - 1) / sb * sb;
				const size_t n_chunk = (dev_bytes + chunk - 1) / chunk;
				const size_t region = 16 + (chunk / sb) * (4 + sb) + need + 64; // worst case of a chunk + the slack that keeps all of it on the fast path
				if (n_chunk >= 2 && n_chunk <= (size_t)stenos_context_s::PIPE_MAX && ctx->in.reserve(dev_bytes + 16) && ctx->out.reserve(n_chunk * region + 16)) {
					unsigned long long* d_res = reinterpret_cast<unsigned long long*>(ctx->pipe_res.p);
					cudaEventRecord(ctx->pipe_ev[0], st); // work already queued on the context's stream comes first
					cudaStreamWaitEvent(ctx->pipe_in, ctx->pipe_ev[0], 0);
					for (size_t i = 0; i < n_chunk; ++i) {
						const size_t off = i * chunk, len = std::min(chunk, dev_bytes - off);
						cudaMemcpyAsync(ctx->in.p + off, src + off, len, cudaMemcpyHostToDevice, ctx->pipe_in);
						cudaEventRecord(ctx->pipe_ev[2 * i + 1], ctx->pipe_in);
					}
					size_t err = 0;
					for (size_t i = 0; i < n_chunk && !err; ++i) {
						const size_t off = i * chunk, len = std::min(chunk, dev_bytes - off);
						cudaStreamWaitEvent(st, ctx->pipe_ev[2 * i + 1], 0);
						const size_t r = enqueue_encode(ctx, ctx->in.p + off, T, len, ctx->out.p + i * region, region, sb, i == 0 ? header_len : 0u, (uint32_t)ctx->shift, bytes, level,
										d_res + 2 * i, nullptr);
						if (is_err(r))
							err = r;
						cudaMemcpyAsync(ctx->pipe_host + 2 * i, d_res + 2 * i, 16, cudaMemcpyDeviceToHost, st);
						cudaEventRecord(ctx->pipe_ev[2 * i], st);
					}
					size_t total = 0;
					for (size_t i = 0; i < n_chunk && !err; ++i) {
						if (cudaEventSynchronize(ctx->pipe_ev[2 * i]) != cudaSuccess) {
							cudaGetLastError();
							err = STENOS_ERROR_UNDEFINED;
							break;
						}
						err = map_device_error(ctx->pipe_host[2 * i + 1]);
						const size_t len = (size_t)ctx->pipe_host[2 * i];
						if (!err && total + len > dst_size)
							err = STENOS_ERROR_DST_OVERFLOW;
						if (err)
							break;
						cudaMemcpyAsync(dst + total, ctx->out.p + i * region, len, cudaMemcpyDeviceToHost, ctx->pipe_out);
						total += len;
					}
					cudaStreamSynchronize(ctx->pipe_in);
					cudaStreamSynchronize(st);
					if (cudaStreamSynchronize(ctx->pipe_out) != cudaSuccess) {
						cudaGetLastError();
						return STENOS_ERROR_UNDEFINED;
					}
					if (err)
						return err;
					if (host_tail) {
						const size_t tr = encode_host_tail(src + dev_bytes, last_bytes, dst + total, dst_size - total);
						if (is_err(tr))
							return tr;
						total += tr;
					}
					return total;
				}
			}
		}

		// ---- stage input
		const uint8_t* d_src = src;
		const int zc = (src_dev && dst_dev) ? 0 : zero_copy_mode();
		bool src_direct = src_dev;
		if (!src_dev && (zc & 1)) {
			if (const uint8_t* a = pinned_alias(src)) {
				d_src = a;
				src_direct = true;
			}
		}
		if (dev_bytes) {
			if (!src_direct || ((uintptr_t)d_src & 15u)) {
				if (!ctx->in.reserve(dev_bytes + 16))
					return STENOS_ERROR_ALLOC;
				cudaMemcpyAsync(ctx->in.p, src, dev_bytes, src_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
				d_src = ctx->in.p;
			}
		}
		// ---- output
		uint8_t* d_dst = dst;
		size_t d_cap = dst_size;
		bool dst_direct = dst_dev;
		if (!dst_dev && (zc & 2)) {
			if (uint8_t* a = pinned_alias(dst)) {
				d_dst = a;
				dst_direct = true;
			}
		}
		// Small host destinations (a cvector bucket is 1 KiB): [result words][stream] sit next to each other in the device
		// staging buffer and come back in ONE copy and ONE stream synchronisation instead of two of each.
		bool small_out = false;
		if (!dst_direct) {
			// The staging buffer never needs more than the worst case of the frame (every superblock
			// stored as COPY); the kernel still receives the caller's dst_size because the reference's
			// room arithmetic depends on it.
			const size_t alloc = std::min(dst_size, (size_t)header_len + n_sb * 4 + bytes + 16);
			if (!ctx->out.reserve(alloc + 32))
				return STENOS_ERROR_ALLOC;
			d_dst = ctx->out.p + 16;
			small_out = !host_tail && dev_bytes != 0 && alloc <= stenos_context_s::SMALL_BYTES && ctx->small_init();
		}

		size_t total = header_len;
		if (dev_bytes || (bytes == 0 && !frame)) {
			size_t r;
			if (bytes == 0) {
				// empty bare superblock: [6][0:3] (stenos.cpp:431-433)
				const uint8_t h[4] = { (uint8_t)CODE_COPY, 0, 0, 0 };
				cudaMemcpyAsync(d_dst, h, 4, cudaMemcpyDefault, st);
				cudaStreamSynchronize(st);
				total = 4;
				r = 0;
			}
			else {
				if (small_out) {
					const size_t alloc = std::min(dst_size, (size_t)header_len + n_sb * 4 + bytes + 16);
					r = enqueue_encode(ctx, d_src, T, dev_bytes, d_dst, d_cap, sb, header_len, (uint32_t)ctx->shift, bytes, level, reinterpret_cast<unsigned long long*>(ctx->out.p),
							   nullptr);
					if (is_err(r))
						return r;
					cudaMemcpyAsync(ctx->small_host, ctx->out.p, 16 + alloc, cudaMemcpyDeviceToHost, st);
					if (cudaStreamSynchronize(st) != cudaSuccess) {
						cudaGetLastError();
						return STENOS_ERROR_UNDEFINED;
					}
					const unsigned long long* hr = reinterpret_cast<const unsigned long long*>(ctx->small_host);
					const size_t e = map_device_error(hr[1]);
					if (e)
						return e;
					total = (size_t)hr[0];
					if (total > dst_size)
						return STENOS_ERROR_DST_OVERFLOW;
					memcpy(dst, ctx->small_host + 16, total);
					return total;
				}
				r = enqueue_encode(ctx, d_src, T, dev_bytes, d_dst, d_cap, sb, header_len, (uint32_t)ctx->shift, bytes, level, nullptr, nullptr);
				if (is_err(r))
					return r;
				cudaMemcpyAsync(ctx->host_result, ctx->ctl.p, 16, cudaMemcpyDeviceToHost, st);
				if (cudaStreamSynchronize(st) != cudaSuccess) {
					cudaGetLastError();
					return STENOS_ERROR_UNDEFINED;
				}
				const size_t e = map_device_error(ctx->host_result[1]);
				if (e)
					return e;
				total = (size_t)ctx->host_result[0];
			}
		}
		else if (frame) {
			// only a tiny tail: the header is written from the host
			uint8_t h[12];
			write_frame_header(h, ctx->shift, bytes, sb, custom);
			cudaMemcpyAsync(d_dst, h, header_len, cudaMemcpyDefault, st);
		}

		if (host_tail) {
			uint8_t tail[128], enc[4 + 256];
			if (src_dev) {
				cudaMemcpyAsync(tail, src + dev_bytes, last_bytes, cudaMemcpyDeviceToHost, st);
				cudaStreamSynchronize(st);
			}
			else
				memcpy(tail, src + dev_bytes, last_bytes);
			if (total + 4 > dst_size)
				return STENOS_ERROR_DST_OVERFLOW;
			size_t len = 0;
			bool zok = false;
			if (!load_zstd())
				return STENOS_ERROR_ZSTD_INTERNAL;
			// zstd_compress_with_context(dst + 4, dst_size - 4, src, bytes, 0) -> zstd level 1 (zstd_wrapper.h:49-56)
			const size_t room = std::min<size_t>(dst_size - total - 4, 256);
			const size_t zr = p_zstd_compress(enc + 4, room, tail, last_bytes, 1);
			if (!p_zstd_iserror(zr) && zr <= last_bytes) {
				zok = true;
				len = zr;
				enc[0] = (uint8_t)CODE_ZSTD;
			}
			if (!zok) { // stenos.cpp:668-669 -> MEMCPY
				if (dst_size - total < last_bytes + 4)
					return STENOS_ERROR_DST_OVERFLOW;
				enc[0] = (uint8_t)CODE_COPY;
				memcpy(enc + 4, tail, last_bytes);
				len = last_bytes;
			}
			enc[1] = (uint8_t)len;
			enc[2] = (uint8_t)(len >> 8);
			enc[3] = (uint8_t)(len >> 16);
			cudaMemcpyAsync(d_dst + total, enc, len + 4, cudaMemcpyDefault, st);
			cudaStreamSynchronize(st);
			total += len + 4;
		}
		if (!dst_direct) {
			cudaMemcpyAsync(dst, d_dst, total, cudaMemcpyDeviceToHost, st);
			if (cudaStreamSynchronize(st) != cudaSuccess) {
				cudaGetLastError();
				return STENOS_ERROR_UNDEFINED;
			}
		}
		return total;
	}

	// host walk over the superblock headers (stenos.cpp:1124-1143); returns 0 or an error code
	size_t host_frame_index(const uint8_t* src, size_t size, size_t first, size_t n_sb, unsigned long long* offs)
	{
		size_t at = first;
		for (size_t i = 0; i < n_sb; ++i) {
			if (at + 4 > size)
				return STENOS_ERROR_SRC_OVERFLOW;
			offs[i] = at;
			const size_t csize = (size_t)src[at + 1] | ((size_t)src[at + 2] << 8) | ((size_t)src[at + 3] << 16);
			if (at + 4 + csize > size)
				return STENOS_ERROR_INVALID_INPUT;
			at += 4 + csize;
		}
		offs[n_sb] = at;
		return 0;
	}

	size_t decompress_hybrid(stenos_context* ctx, const uint8_t* src, bool src_dev, size_t T, size_t size, uint8_t* dst, size_t dst_size, size_t prior_error);

	size_t decompress_level1(stenos_context* ctx, const void* src_, size_t T, size_t size, void* dst_, size_t dst_size, bool frame, size_t bare_sb)
	{
		if (T == 0 || T >= STENOS_MAX_BYTESOFTYPE)
			return STENOS_ERROR_INVALID_BYTESOFTYPE; // stenos.cpp:1067-1068
		if (!ctx->activate())
			return STENOS_ERROR_ALLOC;
		const uint8_t* src = static_cast<const uint8_t*>(src_);
		uint8_t* dst = static_cast<uint8_t*>(dst_);
		const bool src_dev = is_device_ptr(src), dst_dev = is_device_ptr(dst);
		cudaStream_t st = ctx->stream();

		uint64_t total;
		size_t sb, first;
		if (frame) {
			// stenos.cpp:1078-1107
			if (size < 8)
				return STENOS_ERROR_SRC_OVERFLOW;
			uint8_t h[12] = { 0 };
			const size_t hl = std::min<size_t>(size, 12);
			if (src_dev) {
				cudaMemcpyAsync(h, src, hl, cudaMemcpyDeviceToHost, st);
				cudaStreamSynchronize(st);
			}
			else
				memcpy(h, src, hl);
			const unsigned shift = h[0];
			if (shift > 4 && shift != 255)
				return STENOS_ERROR_INVALID_INPUT;
			total = 0;
			for (int i = 0; i < 7; ++i)
				total |= (uint64_t)h[1 + i] << (8 * i);
			if (total > dst_size)
				return STENOS_ERROR_DST_OVERFLOW;
			if (total == 0)
				return 0;
			first = 8;
			if (shift == 255) {
				if (size < 12)
					return STENOS_ERROR_SRC_OVERFLOW;
				sb = (size_t)h[8] | ((size_t)h[9] << 8) | ((size_t)h[10] << 16) | ((size_t)h[11] << 24);
				first = 12;
				if (sb == 0)
					return STENOS_ERROR_INVALID_INPUT;
			}
			else
				sb = default_superblock(T) << shift;
		}
		else {
			// stenos_private_decompress_block, stenos.cpp:780-804: one superblock, dsize = dst_size
			if (size < 4)
				return STENOS_ERROR_SRC_OVERFLOW;
			total = dst_size;
			sb = std::max<size_t>(bare_sb, 1);
			if (total > sb)
				sb = total;
			first = 0;
			if (total == 0)
				return 0;
		}
		if (!device_T(T))
			return STENOS_ERROR_INVALID_PARAMETER;
		const size_t n_sb = (size_t)((total + sb - 1) / sb);
		if (n_sb > 0xFFFFFFF0ull)
			return STENOS_ERROR_INVALID_PARAMETER;

		// ---- stage the compressed bytes and build the superblock index
		const uint8_t* d_src = src;
		if (!ctx->idx.reserve((n_sb + 1) * 8))
			return STENOS_ERROR_ALLOC;
		unsigned long long* d_offs = reinterpret_cast<unsigned long long*>(ctx->idx.p);
		if (!ctx->ctl.reserve(64))
			return STENOS_ERROR_ALLOC;
		unsigned long long* d_result = reinterpret_cast<unsigned long long*>(ctx->ctl.p);
		cudaMemsetAsync(d_result, 0, 16, st);
		unsigned last_code = 0;
		size_t last_at = 0, last_csize = 0;

		// ---- host -> host through the device, pipelined (stenos.cpp:1052-1208 on host buffers): the frame is cut in
		// chunks of whole superblocks (32 MiB of output); the host walks the headers of chunk i + 1 (stenos.cpp:1124-1143)
		// while the compressed bytes of chunk i are on the bus, the decoder runs on chunk i as soon as they have landed
		// and its output leaves for the caller's buffer on a third stream while later chunks are decoded.
		{
			const char* pipe_env = getenv("STENOS_B200_PIPELINE_CHUNK");
			const size_t pipe_chunk = pipe_env ? (size_t)strtoull(pipe_env, nullptr, 10) : ((size_t)32 << 20);
			const size_t last_dsize0 = (size_t)(total - (uint64_t)(n_sb - 1) * sb);
			if (frame && !src_dev && !dst_dev && pipe_chunk && total >= 2 * (uint64_t)pipe_chunk && last_dsize0 >= 128 && ctx->pipe_init()) {
				size_t chunk_sb = std::max<size_t>(1, std::max<size_t>(pipe_chunk, ((size_t)total + stenos_context_s::PIPE_MAX - 1) / stenos_context_s::PIPE_MAX) / sb);
				const size_t n_chunk = (n_sb + chunk_sb - 1) / chunk_sb;
				if (n_chunk >= 2 && n_chunk <= (size_t)stenos_context_s::PIPE_MAX && ctx->pipe_offs_reserve(n_sb + 1) && ctx->in.reserve(size + 32) &&
				    ctx->out.reserve((size_t)total + 32)) {
					unsigned long long* offs = ctx->pipe_offs;
					cudaEventRecord(ctx->pipe_ev[0], st); // work already queued on the context's stream (the memset above) comes first
					cudaStreamWaitEvent(ctx->pipe_in, ctx->pipe_ev[0], 0);
					size_t at = first, err = 0;
					for (size_t c = 0; c < n_chunk && !err; ++c) {
						const size_t lo = c * chunk_sb, hi = std::min(n_sb, lo + chunk_sb);
						const size_t at0 = at;
						for (size_t i = lo; i < hi; ++i) { // host_frame_index for this chunk
							if (at + 4 > size) {
								err = STENOS_ERROR_SRC_OVERFLOW;
								break;
							}
							offs[i] = at;
							const size_t csize = (size_t)src[at + 1] | ((size_t)src[at + 2] << 8) | ((size_t)src[at + 3] << 16);
							if (at + 4 + csize > size) {
								err = STENOS_ERROR_INVALID_INPUT;
								break;
							}
							at += 4 + csize;
						}
						if (err)
							break;
						offs[hi] = at;
						cudaMemcpyAsync(ctx->in.p + at0, src + at0, at - at0, cudaMemcpyHostToDevice, ctx->pipe_in);
						cudaMemcpyAsync(d_offs + lo, offs + lo, (hi - lo + 1) * 8, cudaMemcpyHostToDevice, ctx->pipe_in);
						cudaEventRecord(ctx->pipe_ev[2 * c + 1], ctx->pipe_in);
						cudaStreamWaitEvent(st, ctx->pipe_ev[2 * c + 1], 0);
						DecodeParams P;
						P.ticket = nullptr;
						P.src = ctx->in.p;
						P.src_size = at; // everything up to the end of this chunk is on the device
						P.dst = ctx->out.p;
						P.total = total;
						P.sb_bytes = (uint32_t)sb;
						P.n_sb = (uint32_t)(hi - lo);
						P.first_sb = (uint32_t)lo;
						P.sb_offsets = d_offs;
						P.result = d_result;
						P.skip_zstd_tail = 0;
						P.dst_origin = 0;
						const size_t lr = launch_decode(ctx, T, P);
						if (is_err(lr)) {
							err = lr;
							break;
						}
						cudaEventRecord(ctx->pipe_ev[2 * c], st);
						cudaStreamWaitEvent(ctx->pipe_out, ctx->pipe_ev[2 * c], 0);
						const size_t o0 = lo * sb, o1 = std::min<uint64_t>((uint64_t)hi * sb, total);
						cudaMemcpyAsync(dst + o0, ctx->out.p + o0, o1 - o0, cudaMemcpyDeviceToHost, ctx->pipe_out);
					}
					cudaMemcpyAsync(ctx->host_result, d_result, 16, cudaMemcpyDeviceToHost, st);
					cudaStreamSynchronize(ctx->pipe_in);
					const bool ok = cudaStreamSynchronize(st) == cudaSuccess;
					if (cudaStreamSynchronize(ctx->pipe_out) != cudaSuccess || !ok) {
						cudaGetLastError();
						return STENOS_ERROR_UNDEFINED;
					}
					if (err)
						return err;
					const size_t e = map_device_error(ctx->host_result[1]);
					if (e)
						return e;
					return (size_t)total;
				}
			}
		}
		// ---- small host -> host calls (a cvector bucket is 1 KiB): pinned offsets, [result words][bytes] back in one copy,
		// one stream synchronisation per call instead of three
		{
			const size_t last_dsize0 = (size_t)(total - (uint64_t)(n_sb - 1) * sb);
			if (!src_dev && !dst_dev && total <= stenos_context_s::SMALL_BYTES && last_dsize0 >= 128 && ctx->small_init() && ctx->pipe_offs_reserve(n_sb + 1) &&
			    ctx->in.reserve(size + 32) && ctx->out.reserve((size_t)total + 64)) {
				const size_t e0 = host_frame_index(src, size, first, n_sb, ctx->pipe_offs);
				if (e0)
					return e0;
				const size_t used = (size_t)ctx->pipe_offs[n_sb];
				unsigned long long* d_res2 = reinterpret_cast<unsigned long long*>(ctx->out.p);
				cudaMemsetAsync(d_res2, 0, 16, st);
				cudaMemcpyAsync(ctx->in.p, src, used, cudaMemcpyHostToDevice, st);
				cudaMemcpyAsync(d_offs, ctx->pipe_offs, (n_sb + 1) * 8, cudaMemcpyHostToDevice, st);
				DecodeParams P;
				P.ticket = nullptr;
				P.src = ctx->in.p;
				P.src_size = used;
				P.dst = ctx->out.p + 16;
				P.total = total;
				P.sb_bytes = (uint32_t)sb;
				P.n_sb = (uint32_t)n_sb;
				P.first_sb = 0;
				P.sb_offsets = d_offs;
				P.result = d_res2;
				P.skip_zstd_tail = 0;
				P.dst_origin = 0;
				const size_t lr = launch_decode(ctx, T, P);
				if (is_err(lr))
					return lr;
				cudaMemcpyAsync(ctx->small_host, ctx->out.p, 16 + (size_t)total, cudaMemcpyDeviceToHost, st);
				if (cudaStreamSynchronize(st) != cudaSuccess) {
					cudaGetLastError();
					return STENOS_ERROR_UNDEFINED;
				}
				const size_t e = map_device_error(reinterpret_cast<const unsigned long long*>(ctx->small_host)[1]);
				if (e)
					return frame ? e : STENOS_ERROR_INVALID_INPUT;
				memcpy(dst, ctx->small_host + 16, (size_t)total);
				return (size_t)total;
			}
		}
		if (!src_dev) {
			unsigned long long* offs = (unsigned long long*)malloc((n_sb + 1) * 8);
			if (!offs)
				return STENOS_ERROR_ALLOC;
			const size_t e = host_frame_index(src, size, first, n_sb, offs);
			if (e) {
				free(offs);
				return e;
			}
			last_at = (size_t)offs[n_sb - 1];
			last_code = src[last_at];
			last_csize = (size_t)(offs[n_sb] - offs[n_sb - 1]) - 4;
			const size_t used = (size_t)offs[n_sb];
			if (!ctx->in.reserve(used + 32)) {
				free(offs);
				return STENOS_ERROR_ALLOC;
			}
			cudaMemcpyAsync(ctx->in.p, src, used, cudaMemcpyHostToDevice, st);
			cudaMemcpyAsync(d_offs, offs, (n_sb + 1) * 8, cudaMemcpyHostToDevice, st);
			cudaStreamSynchronize(st); // offs is pageable: the copy above is staged before returning, but keep it simple
			free(offs);
			d_src = ctx->in.p;
			size = used;
		}
		else {
			const size_t ir = enqueue_index(ctx, src, size, first, n_sb, sb, d_offs, d_result);
			if (is_err(ir))
				return ir;
		}

		// ---- output
		uint8_t* d_dst = dst;
		if (!dst_dev || ((uintptr_t)dst & 15u)) {
			if (!ctx->out.reserve((size_t)total + 32))
				return STENOS_ERROR_ALLOC;
			d_dst = ctx->out.p;
		}
		DecodeParams P;
		P.ticket = nullptr;
		P.src = d_src;
		P.src_size = size;
		P.dst = d_dst;
		P.total = total;
		P.sb_bytes = (uint32_t)sb;
		P.n_sb = (uint32_t)n_sb;
		P.first_sb = 0;
		P.sb_offsets = d_offs;
		P.result = d_result;
		P.skip_zstd_tail = 1;
		P.dst_origin = 0;
		const size_t lr = launch_decode(ctx, T, P);
		if (is_err(lr))
			return lr;
		cudaMemcpyAsync(ctx->host_result, d_result, 16, cudaMemcpyDeviceToHost, st);
		// the last superblock may be a Zstd coded tail (< 128 bytes): needs its header on the host
		const size_t last_dsize = (size_t)(total - (uint64_t)(n_sb - 1) * sb);
		uint8_t tail_hdr[4 + 256];
		if (src_dev && last_dsize < 128) {
			cudaMemcpyAsync(ctx->host_result + 2, d_offs + (n_sb - 1), 16, cudaMemcpyDeviceToHost, st);
			cudaStreamSynchronize(st);
			last_at = (size_t)ctx->host_result[2];
			if (last_at + 4 <= size) {
				cudaMemcpyAsync(tail_hdr, src + last_at, std::min<size_t>(size - last_at, sizeof(tail_hdr)), cudaMemcpyDeviceToHost, st);
				cudaStreamSynchronize(st);
				last_code = tail_hdr[0];
				last_csize = (size_t)tail_hdr[1] | ((size_t)tail_hdr[2] << 8) | ((size_t)tail_hdr[3] << 16);
			}
		}
		if (cudaStreamSynchronize(st) != cudaSuccess) {
			cudaGetLastError();
			return STENOS_ERROR_UNDEFINED;
		}
		const size_t e = map_device_error(ctx->host_result[1]);
		if (e)
			return frame ? e : STENOS_ERROR_INVALID_INPUT;
		if (last_dsize < 128 && last_code == (unsigned)CODE_ZSTD) {
			// stenos.cpp:694-699
			if (!load_zstd())
				return STENOS_ERROR_ZSTD_INTERNAL;
			if (last_csize > 256 || last_at + 4 + last_csize > size)
				return STENOS_ERROR_INVALID_INPUT;
			uint8_t dec[128];
			const uint8_t* payload;
			if (src_dev)
				payload = tail_hdr + 4;
			else
				payload = src + last_at + 4;
			const size_t zr = p_zstd_decompress(dec, last_dsize, payload, last_csize);
			if (p_zstd_iserror(zr) || zr != last_dsize)
				return STENOS_ERROR_INVALID_INPUT;
			cudaMemcpyAsync(d_dst + (total - last_dsize), dec, last_dsize, cudaMemcpyHostToDevice, st);
			cudaStreamSynchronize(st);
		}
		if (d_dst != dst) {
			cudaMemcpyAsync(dst, d_dst, (size_t)total, dst_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
			if (cudaStreamSynchronize(st) != cudaSuccess) {
				cudaGetLastError();
				return STENOS_ERROR_UNDEFINED;
			}
		}
		return (size_t)total;
	}

	// stenos_decompress_generic / stenos_private_decompress_block.  Frames written by the reference at level >= 2 hold
	// superblocks that went through Zstd (codes 2..5, stenos.cpp:34-39): those take the hybrid path below (host Zstd,
	// device filters / block decoder).  The first superblock's code is looked at up front; a frame that starts with
	// level-1 superblocks and has Zstd ones further on comes back from the device decoder as invalid input and is
	// then looked at again by the hybrid path (which returns that error if it finds no Zstd superblock either).
	size_t decompress_impl(stenos_context* ctx, const void* src_, size_t T, size_t size, void* dst_, size_t dst_size, bool frame, size_t bare_sb)
	{
		if (!frame || !supported_T(T) || size < 13) // (hybrid frames: T in {2,4,8}; the generic sizes decode level 0 / 1 frames)
			return decompress_level1(ctx, src_, T, size, dst_, dst_size, frame, bare_sb);
		if (!ctx->activate())
			return STENOS_ERROR_ALLOC;
		const uint8_t* src = static_cast<const uint8_t*>(src_);
		const bool src_dev = is_device_ptr(src);
		uint8_t h[16] = { 0 };
		const size_t hl = std::min<size_t>(size, 16);
		if (src_dev) {
			cudaMemcpyAsync(h, src, hl, cudaMemcpyDeviceToHost, ctx->stream());
			cudaStreamSynchronize(ctx->stream());
		}
		else
			memcpy(h, src, hl);
		const size_t first = h[0] == 255 ? 12 : 8;
		const unsigned code0 = first < hl ? h[first] : 0u;
		uint64_t total = 0;
		for (int i = 0; i < 7; ++i)
			total |= (uint64_t)h[1 + i] << (8 * i);
		// (a code-2 superblock of a level-1 frame is its tail of < 128 bytes: decompress_level1 handles it)
		if (code0 == 3u || code0 == 4u || code0 == 5u || (code0 == (unsigned)CODE_ZSTD && total >= 128))
			return decompress_hybrid(ctx, src, src_dev, T, size, static_cast<uint8_t*>(dst_), dst_size, 0);
		const size_t r = decompress_level1(ctx, src_, T, size, dst_, dst_size, frame, bare_sb);
		if (r == STENOS_ERROR_INVALID_INPUT)
			return decompress_hybrid(ctx, src, src_dev, T, size, static_cast<uint8_t*>(dst_), dst_size, r);
		return r;
	}

	template<class F>
	size_t guarded(F&& f) noexcept
	{
		try {
			return f();
		}
		catch (...) {
			return STENOS_ERROR_ALLOC; // stenos.cpp:190-199
		}
	}
}

extern "C" {

stenos_context* stenos_make_context(void)
{
	void* m = malloc(sizeof(stenos_context_s));
	if (!m)
		return nullptr;
	return new (m) stenos_context_s();
}
void stenos_destroy_context(stenos_context* ctx)
{
	if (ctx) {
		if (ctx->device >= 0)
			cudaSetDevice(ctx->device);
		ctx->free_all();
		ctx->~stenos_context_s();
		free(ctx);
	}
}
void stenos_reset_context(stenos_context* ctx)
{
	if (ctx) {
		ctx->level = 1;
		ctx->threads = 1;
		ctx->max_ns = 0;
	}
}
size_t stenos_set_level(stenos_context* ctx, int level)
{
	ctx->level = level > 9 ? 9 : (level < 0 ? 0 : level);
	return 0;
}
size_t stenos_set_threads(stenos_context* ctx, int threads)
{
	ctx->threads = threads < 1 ? 1 : threads;
	return 0;
}
size_t stenos_set_max_nanoseconds(stenos_context* ctx, uint64_t nanoseconds)
{
	ctx->max_ns = nanoseconds;
	return 0;
}
size_t stenos_set_block_size(stenos_context* ctx, size_t blocksize_shift)
{
	if (blocksize_shift >= 16 && blocksize_shift != STENOS_NO_BLOCK_SHIFT)
		return STENOS_ERROR_INVALID_PARAMETER;
	ctx->custom_shift = blocksize_shift;
	return 0;
}
size_t stenos_memory_footprint(stenos_context* ctx)
{
	return sizeof(stenos_context_s) + ctx->in.cap + ctx->out.cap + ctx->ctl.cap + ctx->idx.cap + (ctx->host_result ? 64 : 0);
}
int stenos_has_error(size_t r)
{
	return r >= STENOS_LAST_ERROR_CODE;
}
size_t stenos_bound(size_t bytes)
{
	const size_t min_superblock = 65792;
	const size_t n = bytes / min_superblock + (bytes % min_superblock ? 1 : 0);
	return 12 + (n == 0 ? 1 : n) * 4 + bytes;
}

size_t stenos_compress_generic(stenos_context* ctx, const void* src, size_t bytesoftype, size_t bytes, void* dst, size_t dst_size)
{
	return guarded([&]() -> size_t {
		const size_t prep = ctx->prepare(bytesoftype, bytes);
		if (is_err(prep))
			return prep;
		return compress_impl(ctx, src, bytesoftype, bytes, dst, dst_size, true, ctx->superblock);
	});
}
size_t stenos_decompress_generic(stenos_context* ctx, const void* src, size_t bytesoftype, size_t bytes, void* dst, size_t dst_size)
{
	return guarded([&]() -> size_t { return decompress_impl(ctx, src, bytesoftype, bytes, dst, dst_size, true, 0); });
}
// stenos_compress / stenos_decompress (stenos.cpp:1210-1226) build a context per call in the reference, where that is a
// few pointer writes.  Here a context owns device scratch, pinned words, streams and events (cudaMalloc, cudaMallocHost and
// stream creation cost more than a small call's kernels), so every calling thread keeps one and reuses it; the thread's
// exit releases it.
namespace
{
	struct ThreadContext
	{
		stenos_context_s ctx;
		int dev = -1; // the device its scratch lives on
		~ThreadContext() { ctx.free_all(); }
	};
	stenos_context_s* thread_context()
	{
		static thread_local ThreadContext tc;
		stenos_context_s* c = &tc.ctx;
		int cur = 0;
		if (cudaGetDevice(&cur) != cudaSuccess)
			cudaGetLastError();
		if (cur != tc.dev) { // the thread switched devices since its last call: the scratch does not follow
			c->free_all();
			c->sm_count = 0;
			tc.dev = cur;
		}
		// back to the defaults of a fresh context (stenos.cpp:94-106); the scratch stays
		c->level = 1;
		c->threads = 1;
		c->max_ns = 0;
		c->custom_shift = STENOS_NO_BLOCK_SHIFT;
		c->device = -1;
		return c;
	}
}
size_t stenos_compress(const void* src, size_t bytesoftype, size_t bytes, void* dst, size_t dst_size, int level)
{
	stenos_context_s* ctx = thread_context();
	ctx->level = level > 9 ? 9 : (level < 0 ? 0 : level);
	return stenos_compress_generic(ctx, src, bytesoftype, bytes, dst, dst_size);
}
size_t stenos_decompress(const void* src, size_t bytesoftype, size_t bytes, void* dst, size_t dst_size)
{
	return stenos_decompress_generic(thread_context(), src, bytesoftype, bytes, dst, dst_size);
}
size_t stenos_get_info(const void* src_, size_t bytesoftype, size_t bytes, stenos_info* info)
{
	// stenos.cpp:1019-1050
	const uint8_t* src = static_cast<const uint8_t*>(src_);
	if (bytes < 8)
		return STENOS_ERROR_SRC_OVERFLOW;
	const unsigned shift = src[0];
	if (shift > 4 && shift != 255)
		return STENOS_ERROR_INVALID_INPUT;
	uint64_t total = 0;
	for (int i = 0; i < 7; ++i)
		total |= (uint64_t)src[1 + i] << (8 * i);
	info->decompressed_size = (size_t)total;
	if (shift == 255) {
		if (bytes < 12)
			return STENOS_ERROR_SRC_OVERFLOW;
		info->superblock_size = (size_t)src[8] | ((size_t)src[9] << 8) | ((size_t)src[10] << 16) | ((size_t)src[11] << 24);
		return 12;
	}
	info->superblock_size = default_superblock(bytesoftype) << shift;
	return 8;
}

struct stenos_timer_s
{
	std::chrono::steady_clock::time_point t0;
};
stenos_timer* stenos_make_timer(void)
{
	stenos_timer* t = (stenos_timer*)malloc(sizeof(stenos_timer_s));
	if (!t)
		return nullptr;
	new (t) stenos_timer_s();
	t->t0 = std::chrono::steady_clock::now();
	return t;
}
void stenos_destroy_timer(stenos_timer* timer)
{
	if (timer)
		free(timer);
}
void stenos_tick(stenos_timer* timer)
{
	timer->t0 = std::chrono::steady_clock::now();
}
uint64_t stenos_tock(stenos_timer* timer)
{
	return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - timer->t0).count();
}

size_t stenos_private_compress_block(stenos_context* ctx, const void* src, size_t bytesoftype, size_t super_block_size, size_t bytes, void* dst, size_t dst_size)
{
	return guarded([&]() -> size_t {
		if (bytesoftype == 0 || bytesoftype >= STENOS_MAX_BYTESOFTYPE)
			return STENOS_ERROR_INVALID_BYTESOFTYPE;
		if (bytes > super_block_size)
			return STENOS_ERROR_INVALID_PARAMETER;
		ctx->superblock = super_block_size;
		return compress_impl(ctx, src, bytesoftype, bytes, dst, dst_size, false, std::max<size_t>(super_block_size, 1));
	});
}
size_t stenos_private_decompress_block(stenos_context* ctx, const void* src, size_t bytesoftype, size_t super_block_size, size_t bytes, void* dst, size_t dst_size)
{
	return guarded([&]() -> size_t {
		ctx->superblock = super_block_size;
		return decompress_impl(ctx, src, bytesoftype, bytes, dst, dst_size, false, super_block_size);
	});
}
size_t stenos_private_block_size(const void* src_, size_t src_size)
{
	if (src_size < 4)
		return STENOS_ERROR_SRC_OVERFLOW;
	const uint8_t* s = static_cast<const uint8_t*>(src_);
	return ((size_t)s[1] | ((size_t)s[2] << 8) | ((size_t)s[3] << 16)) + 4;
}
size_t stenos_private_block_csize(const void* src_)
{
	if (!src_)
		return 0;
	const uint8_t* s = static_cast<const uint8_t*>(src_);
	return ((size_t)s[1] | ((size_t)s[2] << 8) | ((size_t)s[3] << 16)) + 4;
}
size_t stenos_private_create_compression_header(size_t decompressed_size, size_t super_block_size, void* dst_, size_t dst_size)
{
	if (dst_size < 12)
		return STENOS_ERROR_DST_OVERFLOW;
	write_frame_header(static_cast<uint8_t*>(dst_), 255, decompressed_size, super_block_size, true);
	return 12;
}

// ---- device additions -----------------------------------------------------------------------

size_t stenos_set_device(stenos_context* ctx, int device)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return STENOS_ERROR_INVALID_PARAMETER;
	}
	if (device >= n)
		return STENOS_ERROR_INVALID_PARAMETER;
	if (device != ctx->device) {
		ctx->free_all();
		ctx->sm_count = 0;
	}
	ctx->device = device < 0 ? -1 : device;
	return 0;
}
size_t stenos_set_stream(stenos_context* ctx, void* cuda_stream)
{
	ctx->user_stream = (cudaStream_t)cuda_stream;
	ctx->has_user_stream = true;
	return 0;
}

static size_t check_async_args(stenos_context* ctx, size_t T, const void* aligned_ptr)
{
	if (T == 0 || T >= STENOS_MAX_BYTESOFTYPE)
		return STENOS_ERROR_INVALID_BYTESOFTYPE;
	if (!supported_T(T) || ctx->level > 1 || ctx->max_ns)
		return STENOS_ERROR_INVALID_PARAMETER;
	if ((uintptr_t)aligned_ptr & 15u)
		return STENOS_ERROR_INVALID_PARAMETER;
	return ctx->activate() ? 0 : STENOS_ERROR_ALLOC;
}

size_t stenos_b200_superblock_size(stenos_context* ctx, size_t bytesoftype, size_t bytes)
{
	const size_t r = ctx->prepare(bytesoftype, bytes);
	return is_err(r) ? r : ctx->superblock;
}

size_t stenos_b200_compress_async(stenos_context* ctx, const void* d_src, size_t T, size_t bytes, void* d_dst, size_t dst_size, unsigned long long* d_result,
				  unsigned long long* d_sb_offsets)
{
	return guarded([&]() -> size_t {
		size_t r = check_async_args(ctx, T, d_src);
		if (r)
			return r;
		r = ctx->prepare(T, bytes);
		if (is_err(r))
			return r;
		const size_t sb = ctx->superblock;
		if (sb > STENOS_BLOCK_SIZE || !d_result || bytes == 0)
			return STENOS_ERROR_INVALID_PARAMETER;
		const size_t last = bytes - ((bytes - 1) / sb) * sb;
		if (last < 128 && ctx->level >= 1)
			return STENOS_ERROR_INVALID_PARAMETER;
		const bool custom = ctx->custom_shift != STENOS_NO_BLOCK_SHIFT;
		if (dst_size < (custom ? 12u : 8u))
			return STENOS_ERROR_DST_OVERFLOW;
		return enqueue_encode(ctx, (const uint8_t*)d_src, T, bytes, (uint8_t*)d_dst, dst_size, sb, custom ? 12u : 8u, (uint32_t)ctx->shift, bytes, ctx->level, d_result,
				      d_sb_offsets);
	});
}

size_t stenos_b200_compress_segment_async(stenos_context* ctx, const void* d_src, size_t T, size_t seg_bytes, void* d_dst, size_t dst_size, unsigned long long* d_result,
					  unsigned long long* d_sb_offsets)
{
	return guarded([&]() -> size_t {
		size_t r = check_async_args(ctx, T, d_src);
		if (r)
			return r;
		if (!ctx->superblock || !d_result || seg_bytes == 0)
			return STENOS_ERROR_INVALID_PARAMETER; // call stenos_b200_superblock_size() with the FRAME size first
		const size_t sb = ctx->superblock;
		const size_t last = seg_bytes - ((seg_bytes - 1) / sb) * sb;
		if (sb > STENOS_BLOCK_SIZE || (last < 128 && ctx->level >= 1))
			return STENOS_ERROR_INVALID_PARAMETER;
		return enqueue_encode(ctx, (const uint8_t*)d_src, T, seg_bytes, (uint8_t*)d_dst, dst_size, sb, 0, 0, 0, ctx->level, d_result, d_sb_offsets);
	});
}

int stenos_b200_index_accepted(stenos_context* ctx)
{
	if (!ctx || !ctx->index_ran || !ctx->scan.p)
		return -1;
	if (ctx->device >= 0)
		cudaSetDevice(ctx->device);
	uint32_t v = 0;
	cudaMemcpyAsync(&v, ctx->scan.p, 4, cudaMemcpyDeviceToHost, ctx->stream());
	cudaStreamSynchronize(ctx->stream());
	return v ? 1 : 0;
}

size_t stenos_b200_frame_index_async(stenos_context* ctx, const void* d_frame, size_t frame_bytes, size_t T, unsigned long long* d_sb_offsets, size_t capacity,
				     unsigned long long* d_result)
{
	return guarded([&]() -> size_t {
		if (!ctx->activate())
			return STENOS_ERROR_ALLOC;
		cudaStream_t st = ctx->stream();
		uint8_t h[12] = { 0 };
		if (frame_bytes < 8)
			return STENOS_ERROR_SRC_OVERFLOW;
		cudaMemcpyAsync(h, d_frame, std::min<size_t>(frame_bytes, 12), cudaMemcpyDeviceToHost, st);
		cudaStreamSynchronize(st);
		stenos_info info;
		const size_t first = stenos_get_info(h, T, std::min<size_t>(frame_bytes, 12), &info);
		if (is_err(first))
			return first;
		if (!info.superblock_size)
			return STENOS_ERROR_INVALID_INPUT;
		const size_t n_sb = (info.decompressed_size + info.superblock_size - 1) / info.superblock_size;
		if (n_sb + 1 > capacity)
			return STENOS_ERROR_DST_OVERFLOW;
		cudaMemsetAsync(d_result, 0, 16, st);
		const size_t ir = enqueue_index(ctx, (const uint8_t*)d_frame, frame_bytes, first, n_sb, info.superblock_size, d_sb_offsets, d_result);
		return is_err(ir) ? ir : n_sb;
	});
}

size_t stenos_b200_decompress_range_async(stenos_context* ctx, const void* d_frame, size_t frame_bytes, size_t T, size_t decompressed_bytes, size_t first_sb, size_t n_sb,
					  const unsigned long long* d_sb_offsets, void* d_dst, unsigned long long* d_result)
{
	return guarded([&]() -> size_t {
		size_t r = check_async_args(ctx, T, d_dst);
		if (r)
			return r;
		if (!ctx->superblock || !d_sb_offsets || !d_result)
			return STENOS_ERROR_INVALID_PARAMETER;
		cudaMemsetAsync(d_result, 0, 16, ctx->stream());
		DecodeParams P;
		P.ticket = nullptr;
		P.src = (const uint8_t*)d_frame;
		P.src_size = frame_bytes;
		P.dst = (uint8_t*)d_dst;
		P.total = decompressed_bytes;
		P.sb_bytes = (uint32_t)ctx->superblock;
		P.n_sb = (uint32_t)n_sb;
		P.first_sb = (uint32_t)first_sb;
		P.sb_offsets = d_sb_offsets;
		P.result = d_result;
		P.skip_zstd_tail = 0;
		P.dst_origin = (uint64_t)first_sb * ctx->superblock;
		return launch_decode(ctx, T, P);
	});
}

size_t stenos_b200_decompress_async(stenos_context* ctx, const void* d_src, size_t T, size_t bytes, void* d_dst, size_t dst_size, size_t decompressed_bytes,
				    unsigned long long* d_result, const unsigned long long* d_sb_offsets)
{
	return guarded([&]() -> size_t {
		size_t r = check_async_args(ctx, T, d_dst);
		if (r)
			return r;
		if (!d_result || decompressed_bytes > dst_size)
			return STENOS_ERROR_INVALID_PARAMETER;
		// superblock size as the compressor chose it for this size (level <= 1 -> shift 0)
		const int saved = ctx->level;
		ctx->level = 1;
		r = ctx->prepare(T, decompressed_bytes);
		ctx->level = saved;
		if (is_err(r))
			return r;
		const size_t sb = ctx->superblock;
		const size_t n_sb = (decompressed_bytes + sb - 1) / sb;
		cudaStream_t st = ctx->stream();
		cudaMemsetAsync(d_result, 0, 16, st);
		const unsigned long long* offs = d_sb_offsets;
		if (!offs) {
			if (!ctx->idx.reserve((n_sb + 1) * 8))
				return STENOS_ERROR_ALLOC;
			unsigned long long* built = reinterpret_cast<unsigned long long*>(ctx->idx.p);
			const size_t ir = enqueue_index(ctx, (const uint8_t*)d_src, bytes, ctx->custom_shift != STENOS_NO_BLOCK_SHIFT ? 12 : 8, n_sb, sb, built, d_result);
			if (is_err(ir))
				return ir;
			offs = built;
		}
		DecodeParams P;
		P.ticket = nullptr;
		P.src = (const uint8_t*)d_src;
		P.src_size = bytes;
		P.dst = (uint8_t*)d_dst;
		P.total = decompressed_bytes;
		P.sb_bytes = (uint32_t)sb;
		P.n_sb = (uint32_t)n_sb;
		P.first_sb = 0;
		P.sb_offsets = offs;
		P.result = d_result;
		P.skip_zstd_tail = 0;
		P.dst_origin = 0;
		return launch_decode(ctx, T, P);
	});
}

size_t stenos_b200_gather_decode_async(stenos_context* ctx, const void* d_frame, size_t frame_bytes, size_t T, size_t bucket_bytes, size_t decompressed_bytes,
				       const unsigned long long* d_sb_offsets, size_t n_buckets_total, const unsigned int* d_bucket_ids, size_t n, void* d_dst,
				       unsigned long long* d_result)
{
	return guarded([&]() -> size_t {
		size_t r = check_async_args(ctx, T, d_dst);
		if (r)
			return r;
		if (!d_sb_offsets || !d_bucket_ids || !d_result || bucket_bytes == 0 || (bucket_bytes & 15u) || n > 0xFFFFFFF0ull)
			return STENOS_ERROR_INVALID_PARAMETER;
		cudaMemsetAsync(d_result, 0, 16, ctx->stream());
		if (!n)
			return 0;
		GatherParams P;
		P.src = (const uint8_t*)d_frame;
		P.src_size = frame_bytes;
		P.dst = (uint8_t*)d_dst;
		P.total = decompressed_bytes;
		P.bucket_bytes = (uint32_t)bucket_bytes;
		P.n_buckets = (uint32_t)n_buckets_total;
		P.sb_offsets = d_sb_offsets;
		P.ids = d_bucket_ids;
		P.n = (uint32_t)n;
		P.result = d_result;
		switch (T) {
			case 2: return launch_gather_T<2>(ctx, P);
			case 4: return launch_gather_T<4>(ctx, P);
			case 8: return launch_gather_T<8>(ctx, P);
		}
		return STENOS_ERROR_INVALID_PARAMETER;
	});
}

size_t stenos_b200_compress_buckets_async(stenos_context* ctx, const void* d_src, size_t T, size_t bucket_bytes, size_t total_bytes, const unsigned int* d_bucket_ids,
					  size_t n, void* d_slots, size_t slot_stride, unsigned int* d_sizes, unsigned long long* d_result)
{
	return guarded([&]() -> size_t {
		size_t r = check_async_args(ctx, T, const_cast<void*>(d_src));
		if (r)
			return r;
		if (ctx->level > 1 || ctx->max_ns != 0)
			return STENOS_ERROR_INVALID_PARAMETER;
		if (!d_slots || !d_sizes || !d_result || bucket_bytes == 0 || bucket_bytes % (T * 256) != 0 || bucket_bytes > STENOS_BLOCK_SIZE || slot_stride < 8 ||
		    slot_stride > 0xFFFFFFF0ull || n > 0xFFFFFFF0ull)
			return STENOS_ERROR_INVALID_PARAMETER;
		// a last bucket shorter than 128 bytes is Zstd's on the host (stenos.cpp:435-437): use stenos_private_compress_block for it
		const size_t tail = total_bytes % bucket_bytes;
		if (tail != 0 && tail < 128)
			return STENOS_ERROR_INVALID_PARAMETER;
		cudaStream_t st = ctx->stream();
		cudaMemsetAsync(d_result, 0, 16, st);
		if (!n)
			return 0;
		if (!ctx->dtk.reserve(64))
			return STENOS_ERROR_ALLOC;
		cudaMemsetAsync(ctx->dtk.p, 0, 64, st);
		EncodeParams P;
		P.src = (const uint8_t*)d_src;
		P.bytes = total_bytes;
		P.dst = (uint8_t*)d_slots;
		P.dst_size = (uint64_t)n * slot_stride;
		P.sb_bytes = (uint32_t)bucket_bytes;
		P.n_sb = (uint32_t)n;
		P.header_len = 0;
		P.shift_byte = 255;
		P.frame_bytes = total_bytes;
		P.level = ctx->level;
		P.state = nullptr;
		P.ticket = reinterpret_cast<uint32_t*>(ctx->dtk.p);
		P.result = d_result;
		P.sb_offsets = nullptr;
		P.base_offset = 0;
		P.first_sb = 0;
		P.n_stream = 0;
		P.spill = nullptr;
		P.ring_cap = 0;
		P.bucket_stride = (uint32_t)slot_stride;
		P.bucket_ids = d_bucket_ids;
		P.bucket_sizes = d_sizes;
		switch (T) {
			case 2: return launch_buckets_T<2>(ctx, P);
			case 4: return launch_buckets_T<4>(ctx, P);
			case 8: return launch_buckets_T<8>(ctx, P);
		}
		return STENOS_ERROR_INVALID_PARAMETER;
	});
}

} // extern "C"

// ---- filters --------------------------------------------------------------------------------
namespace
{
	enum FilterOp { OP_SHUFFLE, OP_UNSHUFFLE, OP_DELTA, OP_DELTA_INV };

	template<int T>
	void launch_shuffle_T(stenos_context* ctx, const FilterParams& P, bool inverse)
	{
		const uint64_t nchunks = (P.bytes + P.chunk - 1) / P.chunk;
		const uint64_t groups_per_chunk = (P.chunk / T + 15) / 16 + 1; // + the group that copies the leftover bytes
		const uint64_t per_cta = (uint64_t)FILTER_THREADS * (inverse ? 1 : ShuffleGroups<T>::N); // groups a CTA handles
		const uint64_t ctas_per_chunk = (groups_per_chunk + per_cta - 1) / per_cta;
		FilterParams Q = P;
		Q.chunk_in_y = nchunks <= 65535 ? 1u : 0u; // grid.y is limited to 65535; one of the two always fits
		const dim3 grid = Q.chunk_in_y ? dim3((unsigned)ctas_per_chunk, (unsigned)nchunks) : dim3((unsigned)nchunks, (unsigned)ctas_per_chunk);
		if (inverse)
			STENOS_LAUNCH(unshuffle_kernel<T>, grid, dim3(FILTER_THREADS), 0, ctx->stream(), Q);
		else
			STENOS_LAUNCH(shuffle_kernel<T>, grid, dim3(FILTER_THREADS), 0, ctx->stream(), Q);
		++g_launches;
	}

	size_t filter_impl(stenos_context* ctx, FilterOp op, size_t T, size_t bytes, size_t chunk, const void* src_, void* dst_, int with_delta)
	{
		if (!ctx->activate())
			return STENOS_ERROR_ALLOC;
		if ((op == OP_SHUFFLE || op == OP_UNSHUFFLE) && !(T == 1 || supported_T(T)))
			return STENOS_ERROR_INVALID_PARAMETER;
		if (bytes == 0)
			return 0;
		if (chunk == 0 || chunk > bytes)
			chunk = bytes;
		const uint8_t* src = static_cast<const uint8_t*>(src_);
		uint8_t* dst = static_cast<uint8_t*>(dst_);
		const bool src_dev = is_device_ptr(src), dst_dev = is_device_ptr(dst);
		cudaStream_t st = ctx->stream();
		const uint8_t* d_src = src;
		uint8_t* d_dst = dst;
		if (!src_dev) {
			if (!ctx->in.reserve(bytes + 16))
				return STENOS_ERROR_ALLOC;
			cudaMemcpyAsync(ctx->in.p, src, bytes, cudaMemcpyHostToDevice, st);
			d_src = ctx->in.p;
		}
		if (!dst_dev) {
			if (!ctx->out.reserve(bytes + 16))
				return STENOS_ERROR_ALLOC;
			d_dst = ctx->out.p;
		}
		FilterParams P;
		P.src = d_src;
		P.dst = d_dst;
		P.bytes = bytes;
		P.chunk = chunk;
		P.with_delta = with_delta ? 1u : 0u;
		P.chunk_in_y = 0;
		const uint64_t nchunks = (bytes + chunk - 1) / chunk;
		auto run_delta = [&](const FilterParams& Q, bool inverse) {
			if (inverse) {
				STENOS_LAUNCH(delta_inv_kernel, dim3((unsigned)(nchunks * 4)), dim3(DELTA_INV_THREADS), 128, st, Q);
			}
			else {
				const uint64_t groups = nchunks * ((chunk + 15) / 16);
				STENOS_LAUNCH(delta_kernel, dim3((unsigned)((groups + FILTER_THREADS * DELTA_GROUPS - 1) / (FILTER_THREADS * DELTA_GROUPS))), dim3(FILTER_THREADS), 0, st, Q);
			}
			++g_launches;
		};
		auto run_shuffle = [&](const FilterParams& Q, bool inverse) {
			switch (T) {
				case 1: cudaMemcpyAsync(Q.dst, Q.src, bytes, cudaMemcpyDeviceToDevice, st); break; // shuffle.cpp:86-87
				case 2: launch_shuffle_T<2>(ctx, Q, inverse); break;
				case 4: launch_shuffle_T<4>(ctx, Q, inverse); break;
				case 8: launch_shuffle_T<8>(ctx, Q, inverse); break;
			}
		};
		switch (op) {
			case OP_SHUFFLE:
				if (T == 1 && with_delta) {
					run_delta(P, false);
				}
				else
					run_shuffle(P, false);
				break;
			case OP_UNSHUFFLE:
				if (with_delta) {
					// delta_inv, then unshuffle (stenos.cpp:722-724).  Whole chunks that are a multiple of 16 elements go
					// through the fused kernel (2N of traffic); a ragged tail chunk, tiny chunks and unaligned buffers
					// through delta_inv into scratch + unshuffle (4N).
					uint64_t fused_chunks = 0;
					if (supported_T(T) && chunk > 2048 && chunk % (16 * T) == 0 && (((uintptr_t)d_src | (uintptr_t)d_dst) & 15u) == 0)
						fused_chunks = bytes / chunk;
					if (fused_chunks) {
						FilterParams F = P;
						F.bytes = fused_chunks * chunk;
						switch (T) {
							case 2: STENOS_LAUNCH(unshuffle_delta_kernel<2>, dim3((unsigned)fused_chunks), dim3(UnshuffleDeltaThreads<2>::N), 256 + 2 * 16 * UnshuffleDeltaThreads<2>::N, st, F); break;
							case 4: STENOS_LAUNCH(unshuffle_delta_kernel<4>, dim3((unsigned)fused_chunks), dim3(UnshuffleDeltaThreads<4>::N), 256 + 4 * 16 * UnshuffleDeltaThreads<4>::N, st, F); break;
							case 8: STENOS_LAUNCH(unshuffle_delta_kernel<8>, dim3((unsigned)fused_chunks), dim3(UnshuffleDeltaThreads<8>::N), 256 + 8 * 16 * UnshuffleDeltaThreads<8>::N, st, F); break;
						}
						++g_launches;
					}
					const uint64_t done = fused_chunks * chunk;
					if (done < bytes) {
						const uint64_t rest = bytes - done;
						if (!ctx->idx.reserve(rest + 16))
							return STENOS_ERROR_ALLOC;
						FilterParams A = P;
						A.src = P.src + done;
						A.bytes = rest;
						A.dst = ctx->idx.p;
						const uint64_t rchunks = (rest + chunk - 1) / chunk;
						STENOS_LAUNCH(delta_inv_kernel, dim3((unsigned)(rchunks * 4)), dim3(DELTA_INV_THREADS), 128, st, A);
						++g_launches;
						FilterParams B = A;
						B.src = ctx->idx.p;
						B.dst = P.dst + done;
						B.with_delta = 0;
						run_shuffle(B, true);
					}
				}
				else
					run_shuffle(P, true);
				break;
			case OP_DELTA: run_delta(P, false); break;
			case OP_DELTA_INV: run_delta(P, true); break;
		}
		if (!dst_dev) {
			cudaMemcpyAsync(dst, d_dst, bytes, cudaMemcpyDeviceToHost, st);
			if (cudaStreamSynchronize(st) != cudaSuccess) {
				cudaGetLastError();
				return STENOS_ERROR_UNDEFINED;
			}
		}
		return cudaGetLastError() == cudaSuccess ? bytes : STENOS_ERROR_UNDEFINED;
	}

	// ------------------------------------------------------------------------------------------
	// Hybrid decoder of frames written at level >= 2 (decompress_generic_superblock, stenos.cpp:681-753).  Zstd runs on
	// the host (the reference links it too; here libzstd.so.1 through dlopen), everything else on the device:
	//   code 2  Zstd(raw)                         -> host Zstd, stored as a COPY superblock
	//   code 3  Zstd(shuffle(raw))                -> host Zstd, COPY superblock, then unshuffle_kernel over the superblock
	//   code 4  Zstd(delta(shuffle(raw)))         -> host Zstd, COPY superblock, then unshuffle_delta_kernel
	//   code 5  Zstd(block stream)                -> host Zstd, stored as a code-1 superblock for decode_pairs_kernel
	//   code 1 / 6                                -> as they are
	// The "lowered" superblocks sit in fixed slots of 4 + superblock bytes (the device decoder takes any offsets), so the
	// host threads (stenos_set_threads) work on independent superblocks.  One H2D copy, one decode launch, one filter
	// launch per run of code-3 / code-4 superblocks, chunk = the frame's superblock size.
	// ------------------------------------------------------------------------------------------
	size_t decompress_hybrid(stenos_context* ctx, const uint8_t* src, bool src_dev, size_t T, size_t size, uint8_t* dst, size_t dst_size, size_t prior_error)
	{
		cudaStream_t st = ctx->stream();
		const bool dst_dev = is_device_ptr(dst);
		std::unique_ptr<uint8_t[]> host_frame;
		if (src_dev) {
			host_frame.reset(new (std::nothrow) uint8_t[size]);
			if (!host_frame)
				return STENOS_ERROR_ALLOC;
			cudaMemcpyAsync(host_frame.get(), src, size, cudaMemcpyDeviceToHost, st);
			if (cudaStreamSynchronize(st) != cudaSuccess) {
				cudaGetLastError();
				return STENOS_ERROR_UNDEFINED;
			}
			src = host_frame.get();
		}
		// ---- frame header (stenos.cpp:1078-1107)
		if (size < 8)
			return STENOS_ERROR_SRC_OVERFLOW;
		const unsigned shift = src[0];
		if (shift > 4 && shift != 255)
			return STENOS_ERROR_INVALID_INPUT;
		uint64_t total = 0;
		for (int i = 0; i < 7; ++i)
			total |= (uint64_t)src[1 + i] << (8 * i);
		if (total > dst_size)
			return STENOS_ERROR_DST_OVERFLOW;
		if (total == 0)
			return 0;
		size_t first = 8, sb;
		if (shift == 255) {
			if (size < 12)
				return STENOS_ERROR_SRC_OVERFLOW;
			sb = (size_t)src[8] | ((size_t)src[9] << 8) | ((size_t)src[10] << 16) | ((size_t)src[11] << 24);
			first = 12;
			if (sb == 0)
				return STENOS_ERROR_INVALID_INPUT;
		}
		else
			sb = default_superblock(T) << shift;
		const size_t n_sb = (size_t)((total + sb - 1) / sb);
		if (n_sb > 0xFFFFFFF0ull || sb > 0xFFFFFFu)
			return STENOS_ERROR_INVALID_PARAMETER;
		// ---- the walk (stenos.cpp:1124-1143)
		std::vector<unsigned long long> at(n_sb + 1);
		bool any = false;
		{
			size_t a = first;
			for (size_t i = 0; i < n_sb; ++i) {
				if (a + 4 > size)
					return STENOS_ERROR_SRC_OVERFLOW;
				at[i] = a;
				const unsigned code = src[a];
				any = any || (code >= 2u && code <= 5u);
				const size_t csize = (size_t)src[a + 1] | ((size_t)src[a + 2] << 8) | ((size_t)src[a + 3] << 16);
				if (a + 4 + csize > size)
					return STENOS_ERROR_INVALID_INPUT;
				a += 4 + csize;
			}
			at[n_sb] = a;
		}
		if (!any)
			return prior_error ? prior_error : STENOS_ERROR_INVALID_INPUT; // a level-1 frame: decompress_level1's verdict stands
		if (!load_zstd())
			return STENOS_ERROR_ZSTD_INTERNAL;
		// ---- lower every superblock into its slot
		const size_t slot = (4 + sb + 15) & ~(size_t)15;
		const size_t lowered = n_sb * slot;
		std::unique_ptr<uint8_t[]> low(new (std::nothrow) uint8_t[lowered + 32]);
		std::vector<uint8_t> filt(n_sb, 0);
		if (!low)
			return STENOS_ERROR_ALLOC;
		std::atomic<size_t> next(0);
		std::atomic<size_t> failed(0);
		auto work = [&]() {
			for (;;) {
				const size_t i = next.fetch_add(1);
				if (i >= n_sb || failed.load())
					return;
				const size_t dsize = (size_t)std::min<uint64_t>(sb, total - (uint64_t)i * sb);
				const uint8_t* p = src + at[i];
				const unsigned code = p[0];
				const size_t csize = (size_t)(at[i + 1] - at[i]) - 4;
				uint8_t* o = low.get() + i * slot;
				size_t len = 0;
				unsigned out_code = CODE_COPY;
				if (code == (unsigned)CODE_BLOCK || code == (unsigned)CODE_COPY) {
					if (csize > sb || (code == (unsigned)CODE_COPY && csize != dsize)) {
						failed.store(STENOS_ERROR_INVALID_INPUT);
						return;
					}
					memcpy(o + 4, p + 4, csize);
					len = csize;
					out_code = code;
				}
				else if (code >= 2u && code <= 5u) {
					const size_t cap = code == 5u ? sb : dsize; // :734: the block stream may be as long as a superblock
					const size_t zr = p_zstd_decompress(o + 4, cap, p + 4, csize);
					if (p_zstd_iserror(zr) || ((code == 3u || code == 4u) && zr != dsize)) {
						failed.store(STENOS_ERROR_INVALID_INPUT);
						return;
					}
					if (code == 5u) {
						len = zr;
						out_code = CODE_BLOCK;
						if (zr == 0) {
							failed.store(STENOS_ERROR_INVALID_INPUT);
							return;
						}
					}
					else {
						if (zr < dsize)
							memset(o + 4 + zr, 0, dsize - zr); // code 2 (:697-699 only checks for a Zstd error)
						len = dsize;
						filt[i] = code == 3u ? 1 : (code == 4u ? 2 : 0);
					}
				}
				else {
					failed.store(STENOS_ERROR_INVALID_INPUT); // :746-748
					return;
				}
				o[0] = (uint8_t)out_code;
				o[1] = (uint8_t)len;
				o[2] = (uint8_t)(len >> 8);
				o[3] = (uint8_t)(len >> 16);
			}
		};
		{
			const size_t nt = std::min<size_t>(std::max(1, ctx->threads), std::min<size_t>(n_sb, 256));
			std::vector<std::thread> pool;
			for (size_t t = 1; t < nt; ++t)
				pool.emplace_back(work);
			work();
			for (auto& t : pool)
				t.join();
		}
		if (failed.load())
			return failed.load();
		std::vector<unsigned long long> loffs(n_sb + 1);
		for (size_t i = 0; i <= n_sb; ++i)
			loffs[i] = (unsigned long long)(i * slot);
		// ---- device: copy, decode, inverse filters
		if (!ctx->in.reserve(lowered + 32) || !ctx->hoffs.reserve((n_sb + 1) * 8) || !ctx->ctl.reserve(64) || !ctx->out.reserve((size_t)total + 32))
			return STENOS_ERROR_ALLOC;
		unsigned long long* d_offs = reinterpret_cast<unsigned long long*>(ctx->hoffs.p);
		unsigned long long* d_result = reinterpret_cast<unsigned long long*>(ctx->ctl.p);
		cudaMemsetAsync(d_result, 0, 16, st);
		cudaMemcpyAsync(ctx->in.p, low.get(), lowered, cudaMemcpyHostToDevice, st);
		cudaMemcpyAsync(d_offs, loffs.data(), (n_sb + 1) * 8, cudaMemcpyHostToDevice, st);
		cudaStreamSynchronize(st); // pageable sources
		DecodeParams P;
		P.ticket = nullptr;
		P.src = ctx->in.p;
		P.src_size = lowered;
		P.dst = ctx->out.p;
		P.total = total;
		P.sb_bytes = (uint32_t)sb;
		P.n_sb = (uint32_t)n_sb;
		P.first_sb = 0;
		P.sb_offsets = d_offs;
		P.result = d_result;
		P.skip_zstd_tail = 0;
		P.dst_origin = 0;
		const size_t lr = launch_decode(ctx, T, P);
		if (is_err(lr))
			return lr;
		bool any_filter = false;
		for (size_t i = 0; i < n_sb; ++i)
			any_filter = any_filter || filt[i] != 0;
		if (any_filter && !ctx->hyb.reserve((size_t)total + 32))
			return STENOS_ERROR_ALLOC;
		for (size_t i = 0; i < n_sb;) {
			size_t j = i + 1;
			while (j < n_sb && filt[j] == filt[i])
				++j;
			if (filt[i]) {
				const size_t off = i * sb;
				const size_t bytes = (size_t)std::min<uint64_t>(total, (uint64_t)j * sb) - off;
				const size_t fr = filter_impl(ctx, OP_UNSHUFFLE, T, bytes, sb, ctx->out.p + off, ctx->hyb.p + off, filt[i] == 2 ? 1 : 0);
				if (is_err(fr))
					return fr;
			}
			i = j;
		}
		cudaMemcpyAsync(ctx->host_result, d_result, 16, cudaMemcpyDeviceToHost, st);
		if (cudaStreamSynchronize(st) != cudaSuccess) {
			cudaGetLastError();
			return STENOS_ERROR_UNDEFINED;
		}
		const size_t e = map_device_error(ctx->host_result[1]);
		if (e)
			return e;
		for (size_t i = 0; i < n_sb;) {
			size_t j = i + 1;
			while (j < n_sb && (filt[j] != 0) == (filt[i] != 0))
				++j;
			const size_t off = i * sb;
			const size_t bytes = (size_t)std::min<uint64_t>(total, (uint64_t)j * sb) - off;
			cudaMemcpyAsync(dst + off, (filt[i] ? ctx->hyb.p : ctx->out.p) + off, bytes, dst_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
			i = j;
		}
		if (cudaStreamSynchronize(st) != cudaSuccess) {
			cudaGetLastError();
			return STENOS_ERROR_UNDEFINED;
		}
		return (size_t)total;
	}

	// ------------------------------------------------------------------------------------------
	// Hybrid encoder of level >= 2 with a FORCED strategy (first slice): every superblock as
	// TRANSPOSED_ZSTD (code 3, stenos.cpp:617-634) or TRANSPOSED_DELTA_ZSTD (code 4, :636-656) -- the shuffle (+ byte
	// delta) of the whole input in one device launch, chunk = the level's superblock (prepare(), :159-162), Zstd per
	// superblock on the host threads at the reference's level mapping (:448-456, zstd_wrapper.h:49-56).  What is NOT
	// here is the reference's choice between its strategies (lz4_guess_ratio, :492-558): on inputs where the reference
	// picks the forced strategy for every superblock the frame is identical, byte for byte (tests).
	// ------------------------------------------------------------------------------------------
	size_t compress_strategy_impl(stenos_context* ctx, const uint8_t* src, size_t T, size_t bytes, uint8_t* dst, size_t dst_size, int level, int strategy)
	{
		if (T == 0 || T >= STENOS_MAX_BYTESOFTYPE)
			return STENOS_ERROR_INVALID_BYTESOFTYPE;
		if (!supported_T(T) || level < 2 || level > 9 || (strategy != 3 && strategy != 4 && strategy != 5) || bytes % T != 0 || ctx->custom_shift != STENOS_NO_BLOCK_SHIFT ||
		    ctx->max_ns != 0)
			return STENOS_ERROR_INVALID_PARAMETER;
		if (strategy == 5 && level != 2)
			return STENOS_ERROR_INVALID_PARAMETER; // the device block encoder holds superblocks of 128 KiB: level 2 (levels >= 3 use 256 KiB .. 2 MiB)
		if (!ctx->activate())
			return STENOS_ERROR_ALLOC;
		if (!load_zstd() || !p_zstd_maxclevel)
			return STENOS_ERROR_ZSTD_INTERNAL;
		cudaStream_t st = ctx->stream();
		const bool src_dev = is_device_ptr(src), dst_dev = is_device_ptr(dst);
		// prepare(), stenos.cpp:159-162
		size_t sb = default_superblock(T);
		int shift = 0;
		if (bytes > sb) {
			shift = (level - 1) / 2;
			sb <<= (size_t)shift;
		}
		if (dst_size < 8)
			return STENOS_ERROR_DST_OVERFLOW;
		// level -> Zstd level (stenos.cpp:448-456: level - 1, skipping 4; zstd_wrapper.h:49-56)
		int reduced = level - 1;
		if (reduced >= 4)
			++reduced;
		const int zlevel = reduced < 1 ? 1 : (reduced < 9 ? 2 * reduced - 1 : p_zstd_maxclevel());
		const size_t n_sb = (bytes + sb - 1) / sb;
		std::vector<uint8_t> frame_host;
		uint8_t* out = dst;
		if (dst_dev) {
			frame_host.resize(std::min<size_t>(dst_size, 8 + bytes + 4 * n_sb + 64));
			out = frame_host.data();
		}
		const size_t out_cap = dst_dev ? frame_host.size() : dst_size;
		write_frame_header(out, shift, bytes, sb, false);
		size_t total = 8;
		if (bytes == 0)
			goto done;
		{
			// ---- the device stage, back to the host: the filtered bytes (strategies 3 / 4), or every superblock's block stream
			// (strategy 5, stenos.cpp:560-603: block_compress_generic into a buffer of `bytes` bytes -- the bucket mode of the
			// block encoder with a slot of 4 + bytes per superblock reproduces that room; then Zstd over the stream)
			const size_t stride5 = sb + 4;
			const size_t stage_bytes = strategy == 5 ? n_sb * stride5 : bytes;
			std::unique_ptr<uint8_t[]> filt(new (std::nothrow) uint8_t[stage_bytes]);
			if (!filt || !ctx->hyb.reserve(stage_bytes + 32))
				return STENOS_ERROR_ALLOC;
			if (strategy == 5) {
				const uint8_t* d_in = src;
				if (!src_dev || ((uintptr_t)src & 15u)) {
					if (!ctx->in.reserve(bytes + 32))
						return STENOS_ERROR_ALLOC;
					cudaMemcpyAsync(ctx->in.p, src, bytes, src_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
					d_in = ctx->in.p;
				}
				if (!ctx->hoffs.reserve(n_sb * 4 + 64) || !ctx->ctl.reserve(64))
					return STENOS_ERROR_ALLOC;
				unsigned int* d_sizes = reinterpret_cast<unsigned int*>(ctx->hoffs.p);
				unsigned long long* d_res = reinterpret_cast<unsigned long long*>(ctx->ctl.p);
				const int saved_level = ctx->level;
				ctx->level = 1; // the block level of every level >= 1 (stenos.cpp:430, block_compress.h:1122-1124)
				const size_t n_full = bytes / sb, last = bytes - n_full * sb;
				size_t br = 0;
				if (n_full)
					br = stenos_b200_compress_buckets_async(ctx, d_in, T, sb, n_full * sb, nullptr, n_full, ctx->hyb.p, stride5, d_sizes, d_res);
				if (!is_err(br) && last >= 128) // the partial last superblock: its own room (4 + its bytes)
					br = stenos_b200_compress_buckets_async(ctx, d_in + n_full * sb, T, sb, last, nullptr, 1, ctx->hyb.p + n_full * stride5, last + 4, d_sizes + n_full, d_res);
				ctx->level = saved_level;
				if (is_err(br))
					return br;
			}
			else {
				const size_t fr = filter_impl(ctx, OP_SHUFFLE, T, bytes, sb, src, ctx->hyb.p, strategy == 4 ? 1 : 0);
				if (is_err(fr))
					return fr;
			}
			cudaMemcpyAsync(filt.get(), ctx->hyb.p, stage_bytes, cudaMemcpyDeviceToHost, st);
			if (cudaStreamSynchronize(st) != cudaSuccess) {
				cudaGetLastError();
				return STENOS_ERROR_UNDEFINED;
			}
			// ---- Zstd per superblock
			struct Piece
			{
				std::unique_ptr<uint8_t[]> z; // (not a vector: no zero fill of the worst case)
				size_t len = 0;
				unsigned code = 0;
			};
			std::vector<Piece> pieces(n_sb);
			std::atomic<size_t> next(0), failed(0);
			auto raw_of = [&](size_t i, size_t n, std::vector<uint8_t>& tmp) -> const uint8_t* {
				if (!src_dev)
					return src + i * sb;
				tmp.resize(n);
				cudaMemcpy(tmp.data(), src + i * sb, n, cudaMemcpyDeviceToHost);
				return tmp.data();
			};
			auto work = [&]() {
				for (;;) {
					const size_t i = next.fetch_add(1);
					if (i >= n_sb || failed.load())
						return;
					const size_t n = std::min(sb, bytes - i * sb);
					Piece& pc = pieces[i];
					std::vector<uint8_t> tmp;
					const uint8_t* in = filt.get() + i * sb;
					unsigned code = (unsigned)strategy;
					int zl = zlevel;
					size_t zn = n; // bytes that go through Zstd
					if (n < 128) { // stenos.cpp:435-437: a tiny superblock is Zstd on the raw bytes, zstd_level 0 -> 1
						in = raw_of(i, n, tmp);
						code = (unsigned)CODE_ZSTD;
						zl = 1;
					}
					else if (strategy == 5) {
						const uint8_t* slot = filt.get() + i * stride5;
						const size_t cs = (size_t)slot[1] | ((size_t)slot[2] << 8) | ((size_t)slot[3] << 16);
						if (slot[0] == (uint8_t)CODE_BLOCK && cs <= n) {
							in = slot + 4;
							zn = cs;
						}
						else {
							// the block coding did not shrink the superblock (:564-572): the reference now picks a Zstd strategy by its
							// LZ4 estimates; the forced strategy falls back to the one it takes when they are low, direct Zstd (:658)
							in = raw_of(i, n, tmp);
							code = (unsigned)CODE_ZSTD;
						}
					}
					const size_t cap = zn + zn / 128 + 512;
					pc.z.reset(new (std::nothrow) uint8_t[std::max(cap, n)]);
					if (!pc.z) {
						failed.store(STENOS_ERROR_ALLOC);
						return;
					}
					const size_t zr = p_zstd_compress(pc.z.get(), cap, in, zn, zl);
					if (code == 5u && (p_zstd_iserror(zr) || zr > zn)) { // :586-596 NO_ZSTD: the block stream as it is, code 1
						memcpy(pc.z.get(), in, zn);
						pc.len = zn;
						pc.code = (unsigned)CODE_BLOCK;
					}
					else if (p_zstd_iserror(zr) || zr > n) { // :627-628 -> MEMCPY (:363-374)
						const uint8_t* raw = raw_of(i, n, tmp);
						memcpy(pc.z.get(), raw, n);
						pc.len = n;
						pc.code = (unsigned)CODE_COPY;
					}
					else {
						pc.len = zr;
						pc.code = code;
					}
				}
			};
			{
				const size_t nt = std::min<size_t>(std::max(1, ctx->threads), std::min<size_t>(n_sb, 256));
				std::vector<std::thread> pool;
				for (size_t t = 1; t < nt; ++t)
					pool.emplace_back(work);
				work();
				for (auto& t : pool)
					t.join();
			}
			if (failed.load())
				return failed.load();
			for (size_t i = 0; i < n_sb; ++i) {
				const Piece& pc = pieces[i];
				if (total + 4 + pc.len > out_cap)
					return STENOS_ERROR_DST_OVERFLOW; // :629-630
				out[total] = (uint8_t)pc.code;
				out[total + 1] = (uint8_t)pc.len;
				out[total + 2] = (uint8_t)(pc.len >> 8);
				out[total + 3] = (uint8_t)(pc.len >> 16);
				memcpy(out + total + 4, pc.z.get(), pc.len);
				total += 4 + pc.len;
			}
		}
	done:
		if (dst_dev) {
			cudaMemcpyAsync(dst, out, total, cudaMemcpyHostToDevice, st);
			if (cudaStreamSynchronize(st) != cudaSuccess) {
				cudaGetLastError();
				return STENOS_ERROR_UNDEFINED;
			}
		}
		return total;
	}
}

extern "C" {

size_t stenos_b200_compress_strategy(stenos_context* ctx, const void* src, size_t T, size_t bytes, void* dst, size_t dst_size, int level, int strategy)
{
	return guarded([&]() -> size_t { return compress_strategy_impl(ctx, static_cast<const uint8_t*>(src), T, bytes, static_cast<uint8_t*>(dst), dst_size, level, strategy); });
}

size_t stenos_b200_shuffle(stenos_context* ctx, size_t T, size_t bytes, size_t chunk, const void* src, void* dst, int with_delta)
{
	return guarded([&]() -> size_t { return filter_impl(ctx, OP_SHUFFLE, T, bytes, chunk, src, dst, with_delta); });
}
size_t stenos_b200_unshuffle(stenos_context* ctx, size_t T, size_t bytes, size_t chunk, const void* src, void* dst, int with_delta)
{
	return guarded([&]() -> size_t { return filter_impl(ctx, OP_UNSHUFFLE, T, bytes, chunk, src, dst, with_delta); });
}
size_t stenos_b200_delta(stenos_context* ctx, size_t bytes, size_t chunk, const void* src, void* dst)
{
	return guarded([&]() -> size_t { return filter_impl(ctx, OP_DELTA, 1, bytes, chunk, src, dst, 1); });
}
size_t stenos_b200_delta_inv(stenos_context* ctx, size_t bytes, size_t chunk, const void* src, void* dst)
{
	return guarded([&]() -> size_t { return filter_impl(ctx, OP_DELTA_INV, 1, bytes, chunk, src, dst, 1); });
}

size_t stenos_b200_synchronize(stenos_context* ctx)
{
	if (!ctx->activate())
		return STENOS_ERROR_ALLOC;
	if (cudaStreamSynchronize(ctx->stream()) != cudaSuccess) {
		cudaGetLastError();
		return STENOS_ERROR_UNDEFINED;
	}
	return 0;
}
// Test support: keeps `ctas` SMs busy (one CTA of smem_kb KiB of shared memory each) for about `ns` nanoseconds on the
// given stream -- the co-residency test launches the encoder next to it.
size_t stenos_b200_test_occupy(void* cuda_stream, int ctas, unsigned smem_kb, unsigned long long ns)
{
#ifndef STENOS_EMU
	const int smem = (int)smem_kb * 1024;
	if (cudaFuncSetAttribute((const void*)sb::occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
		cudaGetLastError();
		return STENOS_ERROR_INVALID_PARAMETER;
	}
	sb::occupy_kernel<<<dim3((unsigned)ctas), dim3(32), smem, (cudaStream_t)cuda_stream>>>(ns);
	return cudaGetLastError() == cudaSuccess ? 0 : STENOS_ERROR_UNDEFINED;
#else
	(void)cuda_stream, (void)ctas, (void)smem_kb, (void)ns;
	return 0;
#endif
}
unsigned long long stenos_b200_kernel_launches(void)
{
	return g_launches.load();
}
const char* stenos_b200_build_target(void)
{
#ifdef STENOS_EMU
	return "emu";
#else
	return "sm_100a";
#endif
}

} // extern "C"
