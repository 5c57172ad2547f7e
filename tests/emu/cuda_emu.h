// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny single-threaded SIMT emulator that lets the product's CUDA sources
// (stenos_b200/csrc/*.cu, *.cuh) be compiled with g++ and unit-tested in a container
// without a GPU.  Every CUDA thread is a ucontext fiber; warp collectives and __syncthreads()
// are rendez-vous points the scheduler switches on, so warp-synchronous code keeps its
// semantics (all 32 lanes of a warp must reach a collective, as with a full mask on hardware).
//
// This is NOT a CPU fallback: the product library (stenos_b200/libstenos_b200.so) is built by
// nvcc from the same sources and never contains or loads this file.  The emulated build is
// tests/emu/libstenos_b200_emu.so and is loaded by tests/test_emu_*.py only.
#pragma once
#ifndef STENOS_EMU
#error "cuda_emu.h is only for the -DSTENOS_EMU test build"
#endif

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>
#include <ucontext.h>

// ---------------------------------------------------------------------------------------------
// Qualifiers
// ---------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3
{
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu
{
	unsigned x, y, z;
};
struct uint2
{
	unsigned x, y;
};
struct uint4
{
	unsigned x, y, z, w;
};
struct int4
{
	int x, y, z, w;
};
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{ a, b, c, d }; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{ a, b }; }

namespace emu
{
	struct Warp
	{
		uint64_t buf[32];
		uint64_t res[32];
		unsigned arrived = 0;
		unsigned gen = 0;
	};
	struct Cta
	{
		unsigned nthreads = 0;
		unsigned bar_arrived = 0;
		unsigned bar_gen = 0;
		std::vector<Warp> warps;
		uint3_emu block_idx;
	};
	struct Lane
	{
		ucontext_t ctx;
		char* stack = nullptr;
		bool done = false;
		Cta* cta = nullptr;
		unsigned tid = 0;
	};
	struct State
	{
		ucontext_t sched;
		Lane* cur = nullptr;
		uint3_emu grid_dim, block_dim;
		std::function<void()> body;
		uint8_t* dyn_smem = nullptr; // per-CTA dynamic shared memory of the running lane's CTA
	};
	inline State& st()
	{
		static State s;
		return s;
	}
	inline void yield()
	{
		Lane* l = st().cur;
		swapcontext(&l->ctx, &st().sched);
	}
	inline void trampoline()
	{
		st().body();
		st().cur->done = true;
		swapcontext(&st().cur->ctx, &st().sched);
	}

	// Runs `ctas_at_once` CTAs concurrently (fibers, round robin), grid in order.
	inline void run_grid(dim3 grid, dim3 block, size_t dyn_smem_bytes, unsigned ctas_at_once, const std::function<void()>& body)
	{
		State& S = st();
		S.grid_dim = { grid.x, grid.y, grid.z };
		S.block_dim = { block.x, block.y, block.z };
		S.body = body;
		const unsigned nthreads = block.x * block.y * block.z;
		const size_t stack_bytes = 256 * 1024;
		unsigned total = grid.x * grid.y * grid.z;
		if (ctas_at_once == 0)
			ctas_at_once = 1;
		for (unsigned first = 0; first < total; first += ctas_at_once) {
			unsigned n = std::min(ctas_at_once, total - first);
			std::vector<Cta> ctas(n);
			std::vector<Lane> lanes((size_t)n * nthreads);
			std::vector<std::vector<uint8_t>> smem(n);
			for (unsigned c = 0; c < n; ++c) {
				unsigned b = first + c;
				ctas[c].nthreads = nthreads;
				ctas[c].warps.resize((nthreads + 31) / 32);
				ctas[c].block_idx = { b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y) };
				smem[c].assign(dyn_smem_bytes + 64, 0xCD);
				for (unsigned t = 0; t < nthreads; ++t) {
					Lane& l = lanes[(size_t)c * nthreads + t];
					l.cta = &ctas[c];
					l.tid = t;
					l.stack = (char*)malloc(stack_bytes);
					getcontext(&l.ctx);
					l.ctx.uc_stack.ss_sp = l.stack;
					l.ctx.uc_stack.ss_size = stack_bytes;
					l.ctx.uc_link = &S.sched;
					makecontext(&l.ctx, (void (*)())trampoline, 0);
				}
			}
			size_t remaining = lanes.size();
			while (remaining) {
				for (size_t i = 0; i < lanes.size(); ++i) {
					Lane& l = lanes[i];
					if (l.done)
						continue;
					S.cur = &l;
					size_t c = i / nthreads;
					// 16-byte aligned dynamic smem base
					S.dyn_smem = (uint8_t*)(((uintptr_t)smem[c].data() + 15) & ~(uintptr_t)15);
					swapcontext(&S.sched, &l.ctx);
					if (l.done) {
						--remaining;
						free(l.stack);
						l.stack = nullptr;
					}
				}
			}
		}
		S.cur = nullptr;
	}

	inline unsigned lane_id() { return st().cur->tid & 31; }
	inline Warp& warp() { return st().cur->cta->warps[st().cur->tid >> 5]; }

	// deposit v, wait for the 32 lanes, return the snapshot
	inline const uint64_t* exchange(uint64_t v)
	{
		Warp& w = warp();
		unsigned n = std::min(32u, st().cur->cta->nthreads - (st().cur->tid & ~31u));
		w.buf[lane_id()] = v;
		unsigned g = w.gen;
		if (++w.arrived == n) {
			memcpy(w.res, w.buf, sizeof(w.res));
			w.arrived = 0;
			++w.gen;
		}
		else {
			while (w.gen == g)
				yield();
		}
		return w.res;
	}
}

// ---------------------------------------------------------------------------------------------
// Built-in variables
// ---------------------------------------------------------------------------------------------
struct emu_tid_t
{
	struct X
	{
		operator unsigned() const { return emu::st().cur->tid % emu::st().block_dim.x; }
	} x;
	struct Y
	{
		operator unsigned() const { return (emu::st().cur->tid / emu::st().block_dim.x) % emu::st().block_dim.y; }
	} y;
	struct Z
	{
		operator unsigned() const { return emu::st().cur->tid / (emu::st().block_dim.x * emu::st().block_dim.y); }
	} z;
};
struct emu_bid_t
{
	struct X
	{
		operator unsigned() const { return emu::st().cur->cta->block_idx.x; }
	} x;
	struct Y
	{
		operator unsigned() const { return emu::st().cur->cta->block_idx.y; }
	} y;
	struct Z
	{
		operator unsigned() const { return emu::st().cur->cta->block_idx.z; }
	} z;
};
struct emu_bdim_t
{
	struct X
	{
		operator unsigned() const { return emu::st().block_dim.x; }
	} x;
	struct Y
	{
		operator unsigned() const { return emu::st().block_dim.y; }
	} y;
	struct Z
	{
		operator unsigned() const { return emu::st().block_dim.z; }
	} z;
};
struct emu_gdim_t
{
	struct X
	{
		operator unsigned() const { return emu::st().grid_dim.x; }
	} x;
	struct Y
	{
		operator unsigned() const { return emu::st().grid_dim.y; }
	} y;
	struct Z
	{
		operator unsigned() const { return emu::st().grid_dim.z; }
	} z;
};
static emu_tid_t threadIdx;
static emu_bid_t blockIdx;
static emu_bdim_t blockDim;
static emu_gdim_t gridDim;
static const int warpSize = 32;

// dynamic shared memory: `extern __shared__ T name[];` is spelled EMU_DYN_SMEM(T, name) in the sources
#define STENOS_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::st().dyn_smem)

// ---------------------------------------------------------------------------------------------
// Synchronisation and warp collectives (full-mask, convergent use only)
// ---------------------------------------------------------------------------------------------
static inline void __syncthreads()
{
	emu::Cta* c = emu::st().cur->cta;
	unsigned g = c->bar_gen;
	if (++c->bar_arrived == c->nthreads) {
		c->bar_arrived = 0;
		++c->bar_gen;
	}
	else {
		while (c->bar_gen == g)
			emu::yield();
	}
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::exchange(0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { emu::yield(); }
static inline unsigned __activemask() { return 0xffffffffu; }

template<class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t* r = emu::exchange(raw);
	int lane = (int)emu::lane_id();
	int base = lane & ~(width - 1);
	uint64_t o = r[base + (src & (width - 1))];
	T out;
	memcpy(&out, &o, sizeof(T));
	return out;
}
template<class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32)
{
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t* r = emu::exchange(raw);
	int lane = (int)emu::lane_id();
	int base = lane & ~(width - 1);
	int src = lane - (int)delta;
	uint64_t o = (src < base) ? raw : r[src];
	T out;
	memcpy(&out, &o, sizeof(T));
	return out;
}
template<class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32)
{
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t* r = emu::exchange(raw);
	int lane = (int)emu::lane_id();
	int base = lane & ~(width - 1);
	int src = lane + (int)delta;
	uint64_t o = (src >= base + width) ? raw : r[src];
	T out;
	memcpy(&out, &o, sizeof(T));
	return out;
}
template<class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t* r = emu::exchange(raw);
	int lane = (int)emu::lane_id();
	(void)width;
	uint64_t o = r[lane ^ m];
	T out;
	memcpy(&out, &o, sizeof(T));
	return out;
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
	const uint64_t* r = emu::exchange(pred ? 1 : 0);
	unsigned m = 0;
	for (int i = 0; i < 32; ++i)
		m |= (unsigned)(r[i] & 1) << i;
	return m;
}
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
	const uint64_t* r = emu::exchange(v);
	unsigned s = 0;
	for (int i = 0; i < 32; ++i)
		s += (unsigned)r[i];
	return s;
}
static inline int __reduce_add_sync(unsigned m, int v) { return (int)__reduce_add_sync(m, (unsigned)v); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
	const uint64_t* r = emu::exchange(v);
	unsigned s = 0;
	for (int i = 0; i < 32; ++i)
		s = std::max(s, (unsigned)r[i]);
	return s;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v)
{
	const uint64_t* r = emu::exchange(v);
	unsigned s = 0xffffffffu;
	for (int i = 0; i < 32; ++i)
		s = std::min(s, (unsigned)r[i]);
	return s;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
	const uint64_t* r = emu::exchange(v);
	unsigned s = 0;
	for (int i = 0; i < 32; ++i)
		s |= (unsigned)r[i];
	return s;
}
static inline unsigned __match_any_sync(unsigned, unsigned v)
{
	const uint64_t* r = emu::exchange(v);
	unsigned m = 0;
	for (int i = 0; i < 32; ++i)
		if ((unsigned)r[i] == v)
			m |= 1u << i;
	return m;
}

// ---------------------------------------------------------------------------------------------
// Integer intrinsics
// ---------------------------------------------------------------------------------------------
namespace emu
{
	// PTX prmt.b32, default mode (selector bit 3 = replicate the sign of the selected byte)
	static inline unsigned prmt_b32(unsigned a, unsigned b, unsigned s)
	{
		uint64_t t = ((uint64_t)b << 32) | a;
		unsigned r = 0;
		for (int i = 0; i < 4; ++i) {
			unsigned sel = (s >> (4 * i)) & 0xF;
			unsigned byte = (unsigned)(t >> (8 * (sel & 7))) & 0xFF;
			if (sel & 8)
				byte = (byte & 0x80) ? 0xFF : 0x00;
			r |= byte << (8 * i);
		}
		return r;
	}
}
// CUDA's __byte_perm only honours the low 3 bits of every selector nibble
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) { return emu::prmt_b32(a, b, s & 0x7777u); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v)
{
	unsigned r = 0;
	for (int i = 0; i < 32; ++i)
		r |= ((v >> i) & 1u) << (31 - i);
	return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
	uint64_t t = ((uint64_t)hi << 32) | lo;
	return (unsigned)(t >> (sh & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh)
{
	uint64_t t = ((uint64_t)hi << 32) | lo;
	return (unsigned)((t << (sh & 31)) >> 32);
}
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned sh)
{
	uint64_t t = ((uint64_t)hi << 32) | lo;
	sh = sh > 32 ? 32 : sh;
	return (unsigned)(t >> sh);
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int emu_s16(unsigned v) { return (int)(int16_t)(v & 0xFFFF); }
static inline unsigned emu_pack16(int lo, int hi) { return ((unsigned)lo & 0xFFFF) | ((unsigned)hi << 16); }
static inline unsigned __vmins2(unsigned a, unsigned b)
{
	return emu_pack16(std::min(emu_s16(a), emu_s16(b)), std::min(emu_s16(a >> 16), emu_s16(b >> 16)));
}
static inline unsigned __vmaxs2(unsigned a, unsigned b)
{
	return emu_pack16(std::max(emu_s16(a), emu_s16(b)), std::max(emu_s16(a >> 16), emu_s16(b >> 16)));
}
static inline unsigned __vimin3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmins2(__vmins2(a, b), c); }
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline unsigned __vminu2(unsigned a, unsigned b)
{
	return (std::min(a & 0xFFFF, b & 0xFFFF)) | (std::min(a >> 16, b >> 16) << 16);
}
static inline unsigned __vmaxu2(unsigned a, unsigned b)
{
	return (std::max(a & 0xFFFF, b & 0xFFFF)) | (std::max(a >> 16, b >> 16) << 16);
}
static inline unsigned __vadd2(unsigned a, unsigned b) { return ((a + b) & 0xFFFF) | (((a >> 16) + (b >> 16)) << 16); }
static inline unsigned __vsub2(unsigned a, unsigned b) { return ((a - b) & 0xFFFF) | (((a >> 16) - (b >> 16)) << 16); }
static inline unsigned __vadd4(unsigned a, unsigned b)
{
	unsigned r = 0;
	for (int i = 0; i < 4; ++i)
		r |= (((a >> (8 * i)) + (b >> (8 * i))) & 0xFF) << (8 * i);
	return r;
}
static inline unsigned __vsub4(unsigned a, unsigned b)
{
	unsigned r = 0;
	for (int i = 0; i < 4; ++i)
		r |= (((a >> (8 * i)) - (b >> (8 * i))) & 0xFF) << (8 * i);
	return r;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
template<class T>
static inline T __ldg(const T* p)
{
	return *p;
}

// ---------------------------------------------------------------------------------------------
// Atomics (single OS thread: plain operations are atomic with respect to the fibers)
// ---------------------------------------------------------------------------------------------
template<class T>
static inline T atomicAdd(T* p, T v)
{
	T o = *p;
	*p = o + v;
	return o;
}
template<class T>
static inline T atomicOr(T* p, T v)
{
	T o = *p;
	*p = o | v;
	return o;
}
template<class T>
static inline T atomicMax(T* p, T v)
{
	T o = *p;
	*p = o > v ? o : v;
	return o;
}
template<class T>
static inline T atomicMin(T* p, T v)
{
	T o = *p;
	*p = o < v ? o : v;
	return o;
}
template<class T>
static inline T atomicExch(T* p, T v)
{
	T o = *p;
	*p = v;
	return o;
}
template<class T>
static inline T atomicCAS(T* p, T cmp, T v)
{
	T o = *p;
	if (o == cmp)
		*p = v;
	return o;
}

// ---------------------------------------------------------------------------------------------
// A sliver of the runtime API, enough for the host layer (stenos_b200/csrc/api.cu)
// ---------------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum
{
	cudaSuccess = 0,
	cudaErrorMemoryAllocation = 2,
	cudaErrorInvalidValue = 1
};
enum cudaMemcpyKind
{
	cudaMemcpyHostToHost = 0,
	cudaMemcpyHostToDevice = 1,
	cudaMemcpyDeviceToHost = 2,
	cudaMemcpyDeviceToDevice = 3,
	cudaMemcpyDefault = 4
};
enum cudaMemoryType
{
	cudaMemoryTypeUnregistered = 0,
	cudaMemoryTypeHost = 1,
	cudaMemoryTypeDevice = 2,
	cudaMemoryTypeManaged = 3
};
struct cudaPointerAttributes
{
	cudaMemoryType type;
	int device;
	void* devicePointer;
	void* hostPointer;
};
struct cudaDeviceProp
{
	int multiProcessorCount;
	size_t sharedMemPerBlockOptin;
	int major, minor;
};
enum cudaFuncAttribute
{
	cudaFuncAttributeMaxDynamicSharedMemorySize = 8
};
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n)
{
	*p = aligned_alloc(256, (n + 255) & ~(size_t)255);
	if (*p)
		memset(*p, 0xAB, n);
	return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p)
{
	free(p);
	return cudaSuccess;
}
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind)
{
	memmove(d, s, n);
	return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0)
{
	memmove(d, s, n);
	return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0)
{
	memset(d, v, n);
	return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n)
{
	memset(d, v, n);
	return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s)
{
	*s = nullptr;
	return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned)
{
	*s = nullptr;
	return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
// events: everything in the emulator runs synchronously, so they are always complete
typedef void* cudaEvent_t;
#define cudaEventDisableTiming 2
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned)
{
	*e = reinterpret_cast<cudaEvent_t>(1);
	return cudaSuccess;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d)
{
	*d = 0;
	return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceCount(int* n)
{
	*n = 1;
	return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
	p->multiProcessorCount = 2;
	p->sharedMemPerBlockOptin = 227 * 1024;
	p->major = 10;
	p->minor = 0;
	return cudaSuccess;
}
// In the emulated build every pointer is "host"; the API layer is told what to treat as device
// memory through this hook (tests flip it to exercise both code paths).
namespace emu
{
	inline bool& pointers_are_device()
	{
		static bool v = false;
		return v;
	}
}
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p)
{
	a->type = emu::pointers_are_device() ? cudaMemoryTypeDevice : cudaMemoryTypeUnregistered;
	a->device = 0;
	a->devicePointer = (void*)p;
	a->hostPointer = (void*)p;
	return cudaSuccess;
}
template<class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int)
{
	return cudaSuccess;
}
#define cudaStreamNonBlocking 1

namespace emu
{
	inline unsigned& ctas_at_once()
	{
		static unsigned v = 2;
		return v;
	}
}
// Kernel launch: STENOS_LAUNCH(kernel, grid, block, smem, stream, args...)
#define STENOS_LAUNCH(kernel, grid, block, smem, stream, ...) emu::run_grid(grid, block, smem, emu::ctas_at_once(), [&]() { kernel(__VA_ARGS__); })
