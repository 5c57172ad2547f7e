/*
 * stenos_b200.h -- C ABI of the B200-native (sm_100a) level-1 stenos path.
 *
 * The first part is the reference's own public API, unchanged in names, argument meaning and
 * error behaviour (reference: /root/reference/stenos/stenos.h:115-301), so that code written
 * against stenos.h -- including the header-only stenos::cvector (stenos/cvector.hpp) -- links
 * against libstenos_b200.so instead of libstenos.so.  Each entry cites the reference function
 * it replaces (file:line relative to /root/reference).
 *
 * The second part is what the device adds: a device selector and a stream on the context, and
 * entry points that work on device-resident buffers without host synchronisation.
 *
 * Domain of the device path: compression levels 0 and 1, no time limit; element sizes 2, 4 and 8 bytes (the fast
 * kernels, every entry point) and 3, 5, 6 and 7 bytes (generic kernels, the stenos_compress / stenos_decompress family only).
 * Decompression also reads the frames the reference writes at levels 2..9 (element sizes 2, 4, 8: Zstd on the host,
 * the rest on the device).  Anything else returns STENOS_ERROR_INVALID_PARAMETER -- there is no CPU fallback.
 * src / dst may be host pointers (pageable or pinned) or device pointers; they are detected.
 */
#ifndef STENOS_B200_H
#define STENOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STENOS_B200_EXPORT __attribute__((visibility("default")))

/* stenos/stenos.h:57-84 */
#define STENOS_BLOCK_SIZE (131072)
#define STENOS_MAX_BLOCK_BYTES ((1u << 24u) - 1u)
#define STENOS_MAX_BYTESOFTYPE (STENOS_MAX_BLOCK_BYTES / 256)
#define STENOS_NO_BLOCK_SHIFT ((size_t)-1)
#define STENOS_ERROR_UNDEFINED ((size_t)(-1))
#define STENOS_ERROR_SRC_OVERFLOW ((size_t)(-2))
#define STENOS_ERROR_ALLOC ((size_t)(-3))
#define STENOS_ERROR_INVALID_INPUT ((size_t)(-4))
#define STENOS_ERROR_INVALID_INSTRUCTION_SET ((size_t)(-5))
#define STENOS_ERROR_DST_OVERFLOW ((size_t)(-6))
#define STENOS_ERROR_INVALID_BYTESOFTYPE ((size_t)(-7))
#define STENOS_ERROR_ZSTD_INTERNAL ((size_t)(-8))
#define STENOS_ERROR_INVALID_PARAMETER ((size_t)(-9))
#define STENOS_LAST_ERROR_CODE ((size_t)(-100))

typedef struct stenos_context_s stenos_context; /* stenos/stenos.h:103 */
typedef struct stenos_timer_s stenos_timer;     /* stenos/stenos.h:266 */
typedef struct stenos_info_s                    /* stenos/stenos.h:242-246 */
{
	size_t decompressed_size;
	size_t superblock_size;
} stenos_info;

/* ---- reference API (same symbols as libstenos) --------------------------------------------- */

/* stenos/internal/stenos.cpp:231-243 */
STENOS_B200_EXPORT stenos_context* stenos_make_context(void);
STENOS_B200_EXPORT void stenos_destroy_context(stenos_context* ctx);
/* stenos.cpp:245-252 (level, threads, time limit back to defaults; custom block size kept, as in the reference) */
STENOS_B200_EXPORT void stenos_reset_context(stenos_context* ctx);
/* stenos.cpp:254-262: clamped to 0..9; levels >= 2 are rejected at compress time */
STENOS_B200_EXPORT size_t stenos_set_level(stenos_context* ctx, int level);
/* stenos.cpp:264-268: accepted and ignored (the GPU grid replaces the thread pool; the stream is
 * always the reference's single-thread stream) */
STENOS_B200_EXPORT size_t stenos_set_threads(stenos_context* ctx, int threads);
/* stenos.cpp:270-274: stored; a non zero limit is rejected at compress time (wall-clock driven output) */
STENOS_B200_EXPORT size_t stenos_set_max_nanoseconds(stenos_context* ctx, uint64_t nanoseconds);
/* stenos.cpp:276-286 */
STENOS_B200_EXPORT size_t stenos_set_block_size(stenos_context* ctx, size_t blocksize_shift);
/* stenos.cpp:288-308: host bytes of the context + its device scratch */
STENOS_B200_EXPORT size_t stenos_memory_footprint(stenos_context* ctx);
/* stenos.cpp:310-314 */
STENOS_B200_EXPORT int stenos_has_error(size_t r);
/* stenos.cpp:316-320, stenos.h:37-42 */
STENOS_B200_EXPORT size_t stenos_bound(size_t bytes);
/* stenos.cpp:844-1017 */
STENOS_B200_EXPORT size_t stenos_compress_generic(stenos_context* ctx, const void* src, size_t bytesoftype, size_t bytes, void* dst, size_t dst_size);
/* stenos.cpp:1052-1208.  Divergence: sizes that are an exact multiple of the
 * superblock size decode correctly here; the reference rejects them. */
STENOS_B200_EXPORT size_t stenos_decompress_generic(stenos_context* ctx, const void* src, size_t bytesoftype, size_t bytes, void* dst, size_t dst_size);
/* stenos.cpp:1210-1218 */
STENOS_B200_EXPORT size_t stenos_compress(const void* src, size_t bytesoftype, size_t bytes, void* dst, size_t dst_size, int level);
/* stenos.cpp:1219-1226 */
STENOS_B200_EXPORT size_t stenos_decompress(const void* src, size_t bytesoftype, size_t bytes, void* dst, size_t dst_size);
/* stenos.cpp:1019-1050 (host pointer) */
STENOS_B200_EXPORT size_t stenos_get_info(const void* src, size_t bytesoftype, size_t bytes, stenos_info* info);
/* stenos.cpp:1232-1257 */
STENOS_B200_EXPORT stenos_timer* stenos_make_timer(void);
STENOS_B200_EXPORT void stenos_destroy_timer(stenos_timer* timer);
STENOS_B200_EXPORT void stenos_tick(stenos_timer* timer);
STENOS_B200_EXPORT uint64_t stenos_tock(stenos_timer* timer);
/* private API used by stenos::cvector: stenos.cpp:768-842 */
STENOS_B200_EXPORT size_t stenos_private_compress_block(stenos_context* ctx, const void* src, size_t bytesoftype, size_t super_block_size, size_t bytes, void* dst, size_t dst_size);
STENOS_B200_EXPORT size_t stenos_private_decompress_block(stenos_context* ctx, const void* src, size_t bytesoftype, size_t super_block_size, size_t bytes, void* dst, size_t dst_size);
STENOS_B200_EXPORT size_t stenos_private_block_size(const void* src, size_t src_size);
STENOS_B200_EXPORT size_t stenos_private_block_csize(const void* src);
STENOS_B200_EXPORT size_t stenos_private_create_compression_header(size_t decompressed_size, size_t super_block_size, void* dst, size_t dst_size);

/* ---- device additions ---------------------------------------------------------------------- */

/* Device selector on the context (BASELINE.json north_star).  device < 0: the calling thread's
 * current CUDA device.  Returns 0 or STENOS_ERROR_INVALID_PARAMETER. */
STENOS_B200_EXPORT size_t stenos_set_device(stenos_context* ctx, int device);
/* Run this context's work on an existing CUDA stream (cudaStream_t passed as void*); NULL = the
 * context's own stream. */
STENOS_B200_EXPORT size_t stenos_set_stream(stenos_context* ctx, void* cuda_stream);

/* Device-resident, asynchronous forms.  All pointers are device pointers; nothing is copied to or
 * from the host and the call returns after enqueueing on the context's stream.
 *   d_result[0] = bytes written (compress) / decompressed (decompress)
 *   d_result[1] = 0, or an error bit set: 1 DST_OVERFLOW, 2 SRC_OVERFLOW, 4 INVALID_INPUT
 * Requirements: level 0/1, bytesoftype in {2,4,8}, d_src (compress) and d_dst (decompress) 16-byte
 * aligned, the final superblock >= 128 bytes (a shorter one needs the host's libzstd: use
 * stenos_compress_generic).  Return value: 0 when enqueued, else an error code.
 * d_sb_offsets (optional, may be NULL): compress writes the offset of every superblock header
 * ([count+1] entries); decompress reads it when given and otherwise walks the frame headers on the
 * device first (the walk is serial: stenos.cpp:1124-1143). */
STENOS_B200_EXPORT size_t stenos_b200_compress_async(stenos_context* ctx, const void* d_src, size_t bytesoftype, size_t bytes, void* d_dst, size_t dst_size,
						    unsigned long long* d_result, unsigned long long* d_sb_offsets);
STENOS_B200_EXPORT size_t stenos_b200_decompress_async(stenos_context* ctx, const void* d_src, size_t bytesoftype, size_t bytes, void* d_dst, size_t dst_size,
						      size_t decompressed_bytes, unsigned long long* d_result, const unsigned long long* d_sb_offsets);

/* Multi-GPU building blocks: a SEGMENT is a run of whole superblocks encoded
 * back to back without the frame header.  A frame is [header] + the segments of all ranks in order.
 * seg_bytes must be a multiple of the superblock size except for the last segment of the frame. */
STENOS_B200_EXPORT size_t stenos_b200_compress_segment_async(stenos_context* ctx, const void* d_src, size_t bytesoftype, size_t seg_bytes, void* d_dst,
							    size_t dst_size, unsigned long long* d_result, unsigned long long* d_sb_offsets);
/* Decodes superblocks [first_sb, first_sb + n_sb) of a frame whose header offsets are d_sb_offsets
 * (absolute offsets into d_frame); d_dst receives the bytes from decompressed offset first_sb*superblock. */
STENOS_B200_EXPORT size_t stenos_b200_decompress_range_async(stenos_context* ctx, const void* d_frame, size_t frame_bytes, size_t bytesoftype,
							    size_t decompressed_bytes, size_t first_sb, size_t n_sb, const unsigned long long* d_sb_offsets,
							    void* d_dst, unsigned long long* d_result);
/* superblock size the context would use for (bytesoftype, bytes) -- prepare(), stenos.cpp:115-185 */
STENOS_B200_EXPORT size_t stenos_b200_superblock_size(stenos_context* ctx, size_t bytesoftype, size_t bytes);
/* Superblock header offsets of a device-resident frame: d_sb_offsets[count+1] (replaces the serial
 * walk of stenos.cpp:1124-1143 by a parallel re-synchronising scan that is verified on the device
 * and falls back to the serial walk when the verification fails). */
STENOS_B200_EXPORT size_t stenos_b200_frame_index_async(stenos_context* ctx, const void* d_frame, size_t frame_bytes, size_t bytesoftype,
						       unsigned long long* d_sb_offsets, size_t capacity, unsigned long long* d_result);
/* Diagnostics: waits for the stream; 1 = the last parallel frame index of this context was accepted,
 * 0 = it fell back to the serial walk, -1 = no parallel index was run. */
STENOS_B200_EXPORT int stenos_b200_index_accepted(stenos_context* ctx);

/* Filters (levels >= 2 pre-Zstd stages): stenos::shuffle / unshuffle / delta / delta_inv
 * (stenos/internal/shuffle.h:33,45; delta.h:33,38) applied independently to every `chunk` bytes of
 * the buffer (chunk = the superblock size of the level; 0 = one chunk).  Host or device pointers.
 * with_delta != 0 fuses the byte delta into the shuffle (TRANSPOSED_DELTA path, stenos.cpp:636-656)
 * and, for unshuffle, undoes it first (stenos.cpp:711-725).  Returns bytes or an error code. */
STENOS_B200_EXPORT size_t stenos_b200_shuffle(stenos_context* ctx, size_t bytesoftype, size_t bytes, size_t chunk, const void* src, void* dst, int with_delta);
STENOS_B200_EXPORT size_t stenos_b200_unshuffle(stenos_context* ctx, size_t bytesoftype, size_t bytes, size_t chunk, const void* src, void* dst, int with_delta);
STENOS_B200_EXPORT size_t stenos_b200_delta(stenos_context* ctx, size_t bytes, size_t chunk, const void* src, void* dst);
STENOS_B200_EXPORT size_t stenos_b200_delta_inv(stenos_context* ctx, size_t bytes, size_t chunk, const void* src, void* dst);

/* Hybrid level >= 2, first slice.  DECODING is complete: stenos_decompress[_generic] reads
 * every frame the reference writes at levels 2..9 for bytesoftype in {2,4,8} (superblock codes 2..5 of
 * stenos.cpp:34-39: Zstd on the host threads of stenos_set_threads, inverse filters and block decoder on the device).
 * ENCODING with a forced strategy: every superblock as code 3 (Zstd over the shuffled input, stenos.cpp:617-634),
 * code 4 (Zstd over the shuffled + byte-delta input, :636-656), level in 2..9, or code 5 (Zstd over the superblock's
 * block stream, :560-603; level 2 only: the device block encoder holds 128 KiB superblocks) (superblock size and Zstd level as the
 * reference maps them); the reference's own choice between strategies (lz4_guess_ratio, :492-558) is not made here,
 * so stenos_compress with level >= 2 still returns STENOS_ERROR_INVALID_PARAMETER.  Where the reference picks the
 * forced strategy for every superblock the frames are identical.  Host or device pointers; synchronous. */
STENOS_B200_EXPORT size_t stenos_b200_compress_strategy(stenos_context* ctx, const void* src, size_t bytesoftype, size_t bytes, void* dst, size_t dst_size, int level,
						       int strategy);

/* stenos::cvector random access (BASELINE.json config 5): a serialized cvector is a frame with
 * shift 255 whose superblocks are the buckets (cvector.hpp:3034-3093).  Given the bucket header
 * offsets (stenos_b200_frame_index_async) decode n buckets, chosen by d_bucket_ids, into a dense
 * [n x bucket_bytes] device array. */
STENOS_B200_EXPORT size_t stenos_b200_gather_decode_async(stenos_context* ctx, const void* d_frame, size_t frame_bytes, size_t bytesoftype, size_t bucket_bytes,
							 size_t decompressed_bytes, const unsigned long long* d_sb_offsets, size_t n_buckets_total,
							 const unsigned int* d_bucket_ids, size_t n, void* d_dst, unsigned long long* d_result);

/* stenos::cvector, the write side (cvector.hpp:1394-1420): compresses n buckets of a device
 * resident array in ONE launch.  Bucket i is bytes [id * bucket_bytes, (id + 1) * bucket_bytes) of d_src (id =
 * d_bucket_ids[i], or i when d_bucket_ids is NULL; the array's last bucket may be shorter) and is written as a bare
 * superblock [code][csize:3][payload] at d_slots + i * slot_stride -- byte for byte what
 * stenos_private_compress_block(ctx, bucket, bytesoftype, bucket_bytes, bytes, slot, slot_stride) returns, including the
 * decisions that depend on the room (cvector passes slot_stride = bucket_bytes + 16).
 * d_sizes[i] = 4 + csize (0 and error bit 1 in d_result[1] when the slot is too small).  bucket_bytes: a multiple of
 * 256 elements, at most 128 KiB.  The slots + sizes are what stenos_b200_gather_decode_async reads through an offset
 * index (offset i = i * slot_stride). */
STENOS_B200_EXPORT size_t stenos_b200_compress_buckets_async(stenos_context* ctx, const void* d_src, size_t bytesoftype, size_t bucket_bytes, size_t total_bytes,
							    const unsigned int* d_bucket_ids, size_t n, void* d_slots, size_t slot_stride, unsigned int* d_sizes,
							    unsigned long long* d_result);

/* Waits for the context's stream. */
STENOS_B200_EXPORT size_t stenos_b200_synchronize(stenos_context* ctx);
/* Test support: keeps `ctas` SMs busy (one CTA holding smem_kb KiB of shared memory each) for about `ns` nanoseconds
 * on the given stream, so that a test can launch the encoder next to a kernel that leaves it only a few SMs. */
STENOS_B200_EXPORT size_t stenos_b200_test_occupy(void* cuda_stream, int ctas, unsigned smem_kb, unsigned long long ns);
/* Number of kernels this library has launched in the process (bench.py reports it as gpu_launches). */
STENOS_B200_EXPORT unsigned long long stenos_b200_kernel_launches(void);
/* "sm_100a" for the product build, "emu" for the CPU test build of tests/emu. */
STENOS_B200_EXPORT const char* stenos_b200_build_target(void);

#ifdef __cplusplus
}
#endif
#endif
