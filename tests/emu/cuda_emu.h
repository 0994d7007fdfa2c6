/*
 * cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A tiny functional emulator of the CUDA execution model, just big enough to
 * run rtlsdr_b200/csrc/scan_kernels.cuh on the GPU-less build box: one OS
 * thread per CUDA thread, std::barrier for __syncthreads / warp shuffles,
 * blocks executed one after another.  It checks index math and data flow, not
 * performance and not data races.  Nothing in the shipped library includes it.
 */
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct dim3 {
	unsigned x, y, z;
	dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline int2 make_int2(int x, int y) { return int2{ x, y }; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{ x, y, z, w }; }

namespace cuda_emu {
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline unsigned char *g_smem = nullptr;
inline std::barrier<> *g_block_bar = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> g_warp_bar;
inline std::vector<uint64_t> g_shfl;

inline unsigned char *dyn_smem() { return g_smem; }
inline void copy16(void *d, const void *s)
{
	if (((uintptr_t)d & 15) || ((uintptr_t)s & 15)) {
		fprintf(stderr, "cuda_emu: cp.async 16 with a misaligned address\n");
		abort();
	}
	memcpy(d, s, 16);
}
inline void copy8(void *d, const void *s)
{
	if (((uintptr_t)d & 7) || ((uintptr_t)s & 7)) {
		fprintf(stderr, "cuda_emu: cp.async 8 with a misaligned address\n");
		abort();
	}
	memcpy(d, s, 8);
}
inline unsigned linear_tid() { return t_threadIdx.x + g_blockDim.x * (t_threadIdx.y + g_blockDim.y * t_threadIdx.z); }

template <class T>
inline T shfl_idx(T v, unsigned src_lane)
{
	unsigned tid = linear_tid(), w = tid / 32, lane = tid % 32;
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	g_shfl[w * 32 + lane] = raw;
	g_warp_bar[w]->arrive_and_wait();
	raw = g_shfl[w * 32 + (src_lane & 31)];
	g_warp_bar[w]->arrive_and_wait();
	T out;
	memcpy(&out, &raw, sizeof(T));
	return out;
}


/* ---- mbarrier / bulk copy / named barriers (functional model) ---- */
struct EmuMbar {
	int init = 0, pending = 0;
	long long tx = 0;
	unsigned phase = 0; /* parity of the phase in progress */
};
inline std::mutex g_mbar_mu;
inline std::map<const void *, EmuMbar> g_mbar;
inline std::map<int, std::unique_ptr<std::barrier<>>> g_named_bar;

inline void mbar_settle(EmuMbar &m)
{
	if (m.pending == 0 && m.tx == 0) {
		m.phase ^= 1u;
		m.pending = m.init;
	}
}
inline void mbar_init(const void *bar, int count)
{
	std::lock_guard<std::mutex> lk(g_mbar_mu);
	EmuMbar m;
	m.init = m.pending = count;
	g_mbar[bar] = m;
}
inline void mbar_arrive(const void *bar, unsigned expect_bytes)
{
	std::lock_guard<std::mutex> lk(g_mbar_mu);
	EmuMbar &m = g_mbar.at(bar);
	m.tx += expect_bytes;
	if (--m.pending < 0) {
		fprintf(stderr, "cuda_emu: more arrivals than the mbarrier was initialised for\n");
		abort();
	}
	mbar_settle(m);
}
inline void mbar_wait(const void *bar, unsigned parity)
{
	for (;;) {
		{
			std::lock_guard<std::mutex> lk(g_mbar_mu);
			if (g_mbar.at(bar).phase != (parity & 1u))
				return;
		}
		std::this_thread::yield();
	}
}
inline void bulk_copy(void *d, const void *s, unsigned bytes, const void *bar)
{
	if (((uintptr_t)d & 15) || ((uintptr_t)s & 15) || (bytes & 15)) {
		fprintf(stderr, "cuda_emu: cp.async.bulk with a misaligned address or size\n");
		abort();
	}
	memcpy(d, s, bytes);
	std::lock_guard<std::mutex> lk(g_mbar_mu);
	EmuMbar &m = g_mbar.at(bar);
	m.tx -= bytes;
	mbar_settle(m);
}
inline void named_bar_sync(int id, int count)
{
	std::barrier<> *b;
	{
		std::lock_guard<std::mutex> lk(g_mbar_mu);
		auto &slot = g_named_bar[id];
		if (!slot)
			slot.reset(new std::barrier<>((std::ptrdiff_t)count));
		b = slot.get();
	}
	b->arrive_and_wait();
}
inline void warp_sync() { g_warp_bar[linear_tid() / 32]->arrive_and_wait(); }

/* run `body` for every thread of every block; blocks are sequential */
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body)
{
	const unsigned nthreads = block.x * block.y * block.z;
	const unsigned nblocks = grid.x * grid.y * grid.z;
	std::vector<unsigned char> smem(smem_bytes + 64);
	g_smem = (unsigned char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
	g_blockDim = block;
	g_gridDim = grid;
	std::barrier<> bar((std::ptrdiff_t)nthreads);
	g_block_bar = &bar;
	const unsigned nwarps = (nthreads + 31) / 32;
	g_warp_bar.clear();
	for (unsigned w = 0; w < nwarps; w++) {
		unsigned cnt = std::min(32u, nthreads - w * 32);
		g_warp_bar.emplace_back(new std::barrier<>((std::ptrdiff_t)cnt));
	}
	g_shfl.assign((size_t)nwarps * 32, 0);
	g_mbar.clear();
	g_named_bar.clear();
	std::vector<std::thread> pool;
	pool.reserve(nthreads);
	for (unsigned tid = 0; tid < nthreads; tid++) {
		pool.emplace_back([&, tid]() {
			t_threadIdx = dim3(tid % block.x, (tid / block.x) % block.y, tid / (block.x * block.y));
			for (unsigned b = 0; b < nblocks; b++) {
				t_blockIdx = dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y));
				body();
				bar.arrive_and_wait(); /* next block reuses the shared memory */
			}
		});
	}
	for (auto &th : pool)
		th.join();
	g_block_bar = nullptr;
}
} // namespace cuda_emu

#define threadIdx (::cuda_emu::t_threadIdx)
#define blockIdx (::cuda_emu::t_blockIdx)
#define blockDim (::cuda_emu::g_blockDim)
#define gridDim (::cuda_emu::g_gridDim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

static inline void __syncthreads() { ::cuda_emu::g_block_bar->arrive_and_wait(); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask)
{
	return ::cuda_emu::shfl_idx(v, (::cuda_emu::linear_tid() % 32) ^ (unsigned)lane_mask);
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline unsigned __brev(unsigned v)
{
	unsigned r = 0;
	for (int i = 0; i < 32; i++) {
		r = (r << 1) | (v & 1u);
		v >>= 1;
	}
	return r;
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c)
{
	for (int i = 0; i < 4; i++)
		c += ((a >> (8 * i)) & 0xFFu) * ((b >> (8 * i)) & 0xFFu);
	return c;
}
/* 16-bit x 8-bit dot products, signed x signed: _lo uses bytes 0, 1 of b, _hi bytes 2, 3 */
static inline int emu_dp2a(int a, int b, int c, int byte0)
{
	const int lo = (int16_t)(a & 0xFFFF), hi = (int16_t)((unsigned)a >> 16);
	const int b0 = (int8_t)(((unsigned)b >> (8 * byte0)) & 0xFF), b1 = (int8_t)(((unsigned)b >> (8 * byte0 + 8)) & 0xFF);
	return c + lo * b0 + hi * b1;
}
static inline int __dp2a_lo(int a, int b, int c) { return emu_dp2a(a, b, c, 0); }
static inline int __dp2a_hi(int a, int b, int c) { return emu_dp2a(a, b, c, 2); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }

static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v)
{
	return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
template <class T>
static inline T emu_atomic_max(T *p, T v)
{
	T old = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
	}
	return old;
}
static inline long long atomicMax(long long *p, long long v) { return emu_atomic_max(p, v); }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { return emu_atomic_max(p, v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
	const unsigned long long v = ((unsigned long long)b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; i++)
		r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
	return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <class F>
static inline unsigned emu_vcmp4(unsigned a, unsigned b, F f)
{
	unsigned r = 0;
	for (int i = 0; i < 4; i++)
		if (f((a >> (8 * i)) & 0xFF, (b >> (8 * i)) & 0xFF))
			r |= 0xFFu << (8 * i);
	return r;
}
static inline unsigned __vcmpeq4(unsigned a, unsigned b) { return emu_vcmp4(a, b, [](unsigned x, unsigned y) { return x == y; }); }
static inline unsigned __vcmpltu4(unsigned a, unsigned b) { return emu_vcmp4(a, b, [](unsigned x, unsigned y) { return x < y; }); }
static inline unsigned __vcmpgtu4(unsigned a, unsigned b) { return emu_vcmp4(a, b, [](unsigned x, unsigned y) { return x > y; }); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
