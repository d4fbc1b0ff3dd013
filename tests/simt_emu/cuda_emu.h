// cuda_emu.h -- a tiny single-threaded SIMT emulator.  TEST INFRASTRUCTURE ONLY.
//
// Lets tests/ compile csnappy_b200/csrc/*.cu with g++ (-DCSB_CPU_EMU) and execute the
// kernels' per-lane code as cooperative fibers (ucontext), with warp collectives
// (__ballot_sync / __shfl_sync / __match_any_sync / __syncwarp, full or partial masks)
// implemented as rendezvous points between the fibers of one warp.  It exists so that the
// warp-level logic of the kernels can be regression-tested in the CPU-only container
// against the oracle; it is never linked into libcsnappy_b200.so and nothing under
// csnappy_b200/ references it.  Fibers are scheduled round-robin and switch only at
// collectives, which is one legal interleaving of independent thread scheduling.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <map>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint4 {
	uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct emu_dim3 {
	unsigned x, y, z;
};
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace csb_emu {

struct Rendezvous {
	unsigned arrived = 0, gen = 0;
	uint32_t slot[32];
};

struct Fiber {
	ucontext_t ctx;
	char *stack = nullptr;
	bool done = false;
	unsigned tid = 0;
};

struct Cta {
	std::vector<Fiber> fibers;
	std::map<std::pair<unsigned, unsigned>, Rendezvous> rv;	 // (warp, mask) -> rendezvous
	ucontext_t sched;
	int cur = -1;
	uint8_t *smem = nullptr;
	std::function<void()> body;
};

extern Cta *g_cta;

inline void yield() { swapcontext(&g_cta->fibers[g_cta->cur].ctx, &g_cta->sched); }

inline Rendezvous &rendezvous(unsigned mask)
{
	return g_cta->rv[{threadIdx.x >> 5, mask}];
}

inline void arrive_and_wait(Rendezvous &r, unsigned mask)
{
	const unsigned gen = r.gen;
	if (++r.arrived == (unsigned)__builtin_popcount(mask)) {
		r.arrived = 0;
		r.gen++;
	} else {
		while (r.gen == gen)
			yield();
	}
}

// every lane in mask contributes v; returns a pointer to the 32 slots (valid until the lane's next collective)
inline const uint32_t *exchange(unsigned mask, uint32_t v, uint32_t *copy)
{
	Rendezvous &r = rendezvous(mask);
	r.slot[threadIdx.x & 31] = v;
	arrive_and_wait(r, mask);
	memcpy(copy, r.slot, sizeof(r.slot));
	arrive_and_wait(r, mask);
	return copy;
}

void run_cta(unsigned cta, unsigned grid, unsigned threads, size_t smem_bytes, std::function<void()> body);
uint8_t *smem_base();

}  // namespace csb_emu

static inline void __syncwarp(unsigned mask = 0xffffffffu)
{
	csb_emu::arrive_and_wait(csb_emu::rendezvous(mask), mask);
}
static inline unsigned __ballot_sync(unsigned mask, bool p)
{
	uint32_t s[32];
	csb_emu::exchange(mask, p ? 1u : 0u, s);
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if (((mask >> i) & 1u) && s[i])
			r |= 1u << i;
	return r;
}
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	static_assert(sizeof(T) == 4, "emulator handles 32-bit shuffles");
	uint32_t s[32], raw;
	memcpy(&raw, &v, 4);
	csb_emu::exchange(mask, raw, s);
	const unsigned lane = threadIdx.x & 31u;
	const unsigned from = (lane & ~(unsigned)(width - 1)) + ((unsigned)src & (unsigned)(width - 1));
	T out;
	memcpy(&out, &s[from], 4);
	return out;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned key)
{
	uint32_t s[32];
	csb_emu::exchange(mask, key, s);
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if (((mask >> i) & 1u) && s[i] == key)
			r |= 1u << i;
	return r;
}
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
	sh &= 31u;
	return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
template <typename T>
static inline T __ldg(const T *p) { return *p; }
template <typename T>
static inline T __ldcg(const T *p) { return *p; }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v)
{
	uint32_t old = *p;
	*p = old + v;
	return old;
}
