// device_common.cuh -- helpers shared by the sm_100a kernels.
//
// Execution model used by both codecs: a GROUP of G lanes (G = 8, 16 or 32, a
// power-of-two slice of one warp) cooperates on one block; a CTA holds as many
// groups as shared memory allows; the grid is persistent (a multiple of the SM
// count) and groups claim blocks from a global counter.  All intra-group
// communication is warp-level (ballot / shfl / match / syncwarp with the
// group's lane mask), never __syncthreads, so groups progress independently.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace csb {

constexpr uint32_t kHashMul = 0x1e35a7bdu;  // csnappy_compress.c:230

template <int G>
struct Group {
	unsigned lane;    // 0..G-1 within the group
	unsigned shift;   // first warp lane of the group
	unsigned mask;    // warp lane mask of the group
	__device__ __forceinline__ Group()
	{
		unsigned wl = threadIdx.x & 31u;
		lane = wl & (G - 1);
		shift = wl & ~(unsigned)(G - 1);
		mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << shift);
	}
	__device__ __forceinline__ void sync() const { __syncwarp(mask); }
	__device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(mask, p) >> shift; }
	template <typename T>
	__device__ __forceinline__ T bcast(T v, int src) const { return __shfl_sync(mask, v, src, G); }
	template <typename T>
	__device__ __forceinline__ T up(T v, unsigned delta) const { return __shfl_up_sync(mask, v, delta, G); }
	__device__ __forceinline__ unsigned match(unsigned key) const { return __match_any_sync(mask, key) >> shift; }
	__device__ __forceinline__ unsigned min(unsigned v) const { return __reduce_min_sync(mask, v); }
	__device__ __forceinline__ unsigned max(unsigned v) const { return __reduce_max_sync(mask, v); }
	__device__ __forceinline__ unsigned add(unsigned v) const { return __reduce_add_sync(mask, v); }
};

// ---- shared-memory addresses, mbarrier and bulk async copy (TMA engine, SASS: UBLKCP) ----------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
		     "l"(src), "r"(bytes), "r"(bar)
		     : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
		     : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest 0 bulk groups have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// non-blocking test of the phase with the given parity
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		     : "=r"(ok)
		     : "r"(bar), "r"(parity)
		     : "memory");
	return ok != 0;
}

// little-endian 32-bit load at an arbitrary byte offset of a shared-memory area (any alignment)
__device__ __forceinline__ uint32_t lds32u(const uint8_t *area, uint32_t off)
{
	const uintptr_t a = reinterpret_cast<uintptr_t>(area) + off;
	const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
	return __funnelshift_r(w[0], w[1], ((uint32_t)a & 3u) * 8u);
}

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
		     : "l"(p));
	return r;
}

__device__ __forceinline__ void stg_stream(uint4 *p, const uint4 &v)
{
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
		     "r"(v.z), "r"(v.w)
		     : "memory");
}

// Cooperative copy of n bytes global -> shared (dst 16-byte aligned).  Vector path when src is
// 16-byte aligned, 4-byte path when 4-aligned, byte path otherwise.  Never reads outside [src, src+n).
template <int G>
__device__ __forceinline__ void load_block_to_smem(const Group<G> &g, uint8_t *dst, const uint8_t *src, uint32_t n)
{
	const uintptr_t a = reinterpret_cast<uintptr_t>(src);
	if ((a & 15u) == 0) {
		const uint32_t nv = n >> 4;
		const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
		uint4 *d4 = reinterpret_cast<uint4 *>(dst);
		for (uint32_t i = g.lane; i < nv; i += G)
			d4[i] = ldg_stream(s4 + i);
		for (uint32_t i = (nv << 4) + g.lane; i < n; i += G)
			dst[i] = src[i];
	} else if ((a & 3u) == 0) {
		const uint32_t nw = n >> 2;
		const uint32_t *s1 = reinterpret_cast<const uint32_t *>(src);
		uint32_t *d1 = reinterpret_cast<uint32_t *>(dst);
		for (uint32_t i = g.lane; i < nw; i += G)
			d1[i] = __ldg(s1 + i);
		for (uint32_t i = (nw << 2) + g.lane; i < n; i += G)
			dst[i] = src[i];
	} else {
		for (uint32_t i = g.lane; i < n; i += G)
			dst[i] = src[i];
	}
}

// explicit shared-space accesses on 32-bit shared addresses (no generic-pointer arithmetic in hot loops)
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v)
{
	asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// little-endian 32-bit load at an arbitrary shared BYTE address
__device__ __forceinline__ uint32_t lds32u_a(uint32_t a)
{
	const uint32_t w = a & ~3u;
	return __funnelshift_r(lds_u32(w), lds_u32(w + 4), (a & 3u) * 8u);
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a)
{
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
	asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v)
{
	asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v)
{
	asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t a, uint32_t x, uint32_t y)
{
	asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

// Stage n bytes of global memory at ANY alignment into shared memory: the 16-byte aligned interior
// travels as one bulk async copy completing on `bar`, the (< 16 byte) head and tail as byte loads.
// `area` is 16-byte aligned with room for n + 16 bytes; the data lands at area + (src & 15) -- the
// returned shift -- so that global and shared addresses agree modulo 16.  Reads nothing outside
// [src, src + n).  *bulk tells whether a bulk copy is in flight (wait for `bar` before reading).
template <int G>
__device__ __forceinline__ uint32_t stage_block(const Group<G> &g, uint8_t *area, const uint8_t *src, uint32_t n,
						uint32_t bar, bool *bulk)
{
	const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
	uint32_t head = (16u - s) & 15u;
	if (head > n)
		head = n;
	const uint32_t mid = (n - head) & ~15u;
	*bulk = mid != 0;
	if (mid && g.lane == 0) {
		fence_proxy_async();  // earlier generic-proxy accesses to this area precede the async-proxy write
		mbar_expect_tx(bar, mid);
		bulk_g2s(smem_u32(area + s + head), src + head, mid, bar);
	}
	for (uint32_t i = g.lane; i < head; i += G)
		area[s + i] = src[i];
	for (uint32_t i = head + mid + g.lane; i < n; i += G)
		area[s + i] = src[i];
	return s;
}

struct DeviceInfo {
	int sm_count;
	int smem_per_block_optin;  // max dynamic shared memory per CTA (232448 on B200)
	int smem_per_sm;	   // 233472 on B200
};

// cached per-device query; returns a cudaError_t
int device_info(DeviceInfo *out);
void count_launch();
uint32_t *next_counter();  // a device word for a kernel's block-claim counter (nullptr on failure)

}  // namespace csb
