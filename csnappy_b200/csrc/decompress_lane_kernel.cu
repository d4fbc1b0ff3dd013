// decompress_lane_kernel.cu -- batched Snappy raw-stream decoder, ONE LANE PER BLOCK.
//
// Second decoder family next to decompress_kernel.cu, for blocks that do not fit shared memory in useful
// numbers (32 KiB fragments and anything larger: 3 staged fragments per SM, or 26 warp-owned blocks on
// the global path).  A block's tag stream is a serial chain whose step costs an L2 round trip once the
// block lives in global memory; the only way to hide that is MANY chains, and a lane is the cheapest
// owner a chain can have: 1024 blocks in flight per SM instead of 26.  Each lane runs the reference's
// loop (csnappy_decompress.c:345-382) literally -- tag, length, checks, copy -- on its own block:
//
//   * ONE load path for the tag stream, literal payloads and back-references: the two aligned 8-byte
//     words around the (arbitrarily aligned) source and a funnel shift; output through an 8-byte
//     accumulator flushed with aligned 8-byte stores, so every memory instruction moves 8 bytes per
//     lane whatever the alignment of the tag;
//   * back-references read the lane's own earlier stores (same thread, same address: ordered); the
//     not yet stored tail of the accumulator is flushed first when a copy reaches into it
//     (offset < 15), and offsets 1, 2 and 4 are expanded in registers (pattern fill);
//   * the loop is FLAT and WARP-UNIFORM: one iteration = claim (free lanes take the next block from
//     the global counter) / tag (lanes whose tag is used up decode the next) / move (every lane moves
//     up to 8 bytes), with a reconvergence point between the phases -- without them the lanes of a
//     warp drift apart and run the loop body one lane at a time (measured: 6.7 of 32 lanes active).
//
// Semantics per block are those of decompress_kernel.cu: first failing tag in stream order decides;
// literal: input shortage (-5) before space (-3); copy: offset validity (-5) before space (-3); end of
// input at a tag boundary is success; a tag header cut off by the end of input is -5 (defined here,
// SURVEY.md 0.5); nothing is ever written at or past the block's capacity.
// The kernel is bound by the LSU (a fully divergent 8-byte access costs ~2 cycles per lane), not by
// HBM; see DESIGN.md 4.3.
#include "device_common.cuh"
#include "kernels.h"

namespace csb {

constexpr int L_OK = 0, L_HEADER_BAD = -1, L_OUTPUT_INSUF = -2, L_OUTPUT_OVERRUN = -3, L_DATA_MALFORMED = -5;
constexpr int kLaneThreads = 256;

__device__ __forceinline__ uint64_t ldg64(uintptr_t a)
{
	uint64_t v;
	asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
	return v;
}
__device__ __forceinline__ void stg64(uintptr_t a, uint64_t v)
{
	asm volatile("st.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}

// `need` (1..8) bytes at byte address a, little endian, without touching a byte at or past `lim`.
// One code path for the tag stream, literal payloads (lim = end of the block) and back-references
// (lim = end of the output slot): two aligned words and a funnel shift.  A word that starts before
// the block still lies inside the caller's allocation (allocations are 256-byte aligned).
__device__ __forceinline__ uint64_t load8(uintptr_t a, uint32_t need, uintptr_t lim)
{
	const uintptr_t w = a & ~(uintptr_t)7;
	const uint32_t k = (uint32_t)a & 7u;
	if (w + 16 <= lim) {
		const uint64_t lo = ldg64(w);
		if (k + need <= 8)
			return lo >> (8u * k);
		return (lo >> (8u * k)) | (ldg64(w + 8) << (64u - 8u * k));
	}
	uint64_t v = 0;	 // the last words of the block / slot: byte by byte
	for (uint32_t i = 0; i < need && a + i < lim; ++i)
		v |= (uint64_t) * reinterpret_cast<const volatile uint8_t *>(a + i) << (8 * i);
	return v;
}

struct LaneParams {
	csb_decompress_args a;
	uint32_t *counter;
};

enum : uint32_t { M_LIT = 0, M_COPY = 1, M_PATTERN = 2, M_NEAR = 3 };

// The loop is warp-uniform: every iteration has a CLAIM phase (lanes without a block take the next one
// from the global counter), a TAG phase (lanes whose tag is used up decode the next one; a lane whose
// block ends here stores its result and becomes free) and a MOVE phase (every lane moves up to 8 bytes
// of its current tag), with the lanes reconverging between the phases.  Literals, back-references and
// the tag stream share ONE load path (load8), so the MOVE phase is free of mode-dependent branches
// except for the rare pattern / near-offset cases.
__global__ void __launch_bounds__(kLaneThreads) decompress_lane_kernel(const LaneParams p)
{
	const csb_decompress_args &a = p.a;
	const unsigned full = 0xffffffffu;
	bool have = false, more = true;
	uint32_t blk = 0, cap = 0, ip = 0, op = 0, rem = 0, mode = M_LIT, off = 0;
	uintptr_t src = 0, src_end = 0, dst = 0;
	uint64_t acc = 0, pat = 0;  // acc: bytes [op & ~7, op) of the output, not yet stored

	for (;;) {
		// ---- CLAIM ----
		if (!have && more) {
			blk = atomicAdd(p.counter, 1u);
			if (blk >= a.n_blocks) {
				more = false;
			} else {
				const uint8_t *s = a.in + (a.in_off ? a.in_off[blk] : (uint64_t)blk * a.in_stride);
				uint32_t ilen = a.in_len[blk];
				cap = a.out_cap ? a.out_cap[blk] : a.uniform_cap;
				dst = reinterpret_cast<uintptr_t>(a.out + (uint64_t)blk * a.out_stride);
				int rc = L_OK;
				if (a.flags & 2u) {  // varint32 length prefix, csnappy_decompress.c:45-71, 404-409
					uint32_t shift = 0, used = 0, value = 0;
					for (;;) {
						if (shift >= 32 || used == ilen) {
							rc = L_HEADER_BAD;
							break;
						}
						const uint32_t c = s[used++];
						value |= (c & 0x7fu) << shift;
						if (c < 128)
							break;
						shift += 7;
					}
					if (rc == L_OK) {
						if (value > cap)
							rc = L_OUTPUT_INSUF;
						cap = value;
						s += used;
						ilen -= used;
					}
				}
				if (rc != L_OK) {
					a.status[blk] = rc;
					a.out_len[blk] = 0u;
				} else {
					have = true;
					src = reinterpret_cast<uintptr_t>(s);
					src_end = src + ilen;
					ip = op = rem = 0;
					acc = 0;
					mode = M_LIT;
					if ((a.flags & 4u) && ilen == cap)
						rem = ilen;  // stored block (block_compressor.c:378): the whole input is one literal payload
				}
			}
		}
		__syncwarp(full);
		if (!__any_sync(full, have || more))
			break;

		// ---- TAG: next tag of the reference's loop (csnappy_decompress.c:345-382) ----
		if (have && rem == 0) {
			int rc = L_OK;
			bool fin = false;
			const uint32_t ilen = (uint32_t)(src_end - src);
			if (ip >= ilen) {
				fin = true;  // end of input at a tag boundary
			} else {
				const uint32_t left = ilen - ip;
				const uint64_t x = load8(src + ip, left < 8 ? left : 8u, src_end);
				const uint32_t tag = (uint32_t)x & 0xffu, kind = tag & 3u;
				uint32_t len = (tag >> 2) + 1;
				if (kind == 0) {
					uint32_t hdr = 1;
					if (len > 60) {
						const uint32_t nb = len - 60;
						if (left - 1 < nb) {
							rc = L_DATA_MALFORMED;	// length bytes cut off (reference: UB)
						} else {
							const uint32_t v = (uint32_t)(x >> 8) & (nb == 4 ? 0xffffffffu : ((1u << (8 * nb)) - 1u));
							len = v + 1;  // 0xffffffff wraps to a zero-length literal (csnappy_decompress.c:370)
							hdr = 1 + nb;
						}
					}
					if (rc == L_OK) {
						ip += hdr;
						// (a length of 2^31 or more passes the reference's signed input check and fails on space, :374)
						if ((int32_t)len >= 0 ? (ilen - ip < len) : (cap - op >= len))
							rc = L_DATA_MALFORMED;
						else if (cap - op < len)
							rc = L_OUTPUT_OVERRUN;
						mode = M_LIT;
						rem = len;
					}
				} else {
					const uint32_t hdr = kind == 1 ? 2u : (kind == 2 ? 3u : 5u);
					if (left < hdr) {
						rc = L_DATA_MALFORMED;	// offset bytes cut off (reference: UB)
					} else {
						if (kind == 1) {
							len = ((tag >> 2) & 7u) + 4;
							off = ((tag >> 5) << 8) | ((uint32_t)(x >> 8) & 0xffu);
						} else if (kind == 2) {
							off = (uint32_t)(x >> 8) & 0xffffu;
						} else {
							off = (uint32_t)(x >> 8);
						}
						ip += hdr;
						if (off - 1u >= op)  // off == 0 or off > produced, csnappy_decompress.c:302
							rc = L_DATA_MALFORMED;
						else if (cap - op < len)
							rc = L_OUTPUT_OVERRUN;
						rem = len;
						// off >= 15: the source never reaches into the unstored accumulator
						mode = off >= 15 ? M_COPY : ((off == 1 || off == 2 || off == 4) ? M_PATTERN : M_NEAR);
					}
				}
			}
			if (rc == L_OK && !fin && mode >= M_PATTERN && rem) {
				// make bytes [op & ~7, op) visible to this lane's loads
				const uint32_t k = op & 7u, w = op & ~7u;
				if (k) {
					if (w + 8 <= cap) {
						stg64(dst + w, acc);  // upper bytes: zeros inside the capacity, rewritten later
					} else {
						for (uint32_t i = 0; i < k; ++i)
							*reinterpret_cast<volatile uint8_t *>(dst + w + i) = (uint8_t)(acc >> (8 * i));
					}
				}
				if (mode == M_PATTERN) {
					pat = load8(dst + op - off, off, dst + cap);
					if (off == 1)
						pat = (pat & 0xffull) * 0x0101010101010101ull;
					else if (off == 2)
						pat = (pat & 0xffffull) * 0x0001000100010001ull;
					else
						pat = (pat & 0xffffffffull) * 0x0000000100000001ull;
				}
			}
			if (rc != L_OK || fin) {
				if (rc == L_OK) {
					const uint32_t k = op & 7u, w = op & ~7u;
					if (k) {
						if (w + 8 <= cap) {
							stg64(dst + w, acc);
						} else {
							for (uint32_t i = 0; i < k; ++i)
								*reinterpret_cast<volatile uint8_t *>(dst + w + i) = (uint8_t)(acc >> (8 * i));
						}
					}
				}
				a.status[blk] = rc;
				a.out_len[blk] = rc == L_OK ? op : 0u;
				have = false;
				rem = 0;
			}
		}
		__syncwarp(full);

		// ---- MOVE: up to 8 bytes of the current tag ----
		if (have && rem) {
			uint32_t n = rem < 8 ? rem : 8;
			uint64_t v = pat;
			if (mode == M_NEAR) {
				// offsets 3, 5..14: rounds of at most `off` bytes read only what is already there; the
				// accumulator's bytes must be in memory first
				if (n > off)
					n = off;
				const uint32_t k = op & 7u, w = op & ~7u;
				if (k) {
					if (w + 8 <= cap) {
						stg64(dst + w, acc);
					} else {
						for (uint32_t i = 0; i < k; ++i)
							*reinterpret_cast<volatile uint8_t *>(dst + w + i) = (uint8_t)(acc >> (8 * i));
					}
				}
			}
			if (mode != M_PATTERN) {
				const bool lit = mode == M_LIT;
				v = load8(lit ? src + ip : dst + op - off, n, lit ? src_end : dst + cap);
				if (lit)
					ip += n;
			}
			// append the n low bytes of v (op + n <= cap was checked with the tag)
			const uint32_t k = op & 7u, sh = 8u * k;
			if (n < 8)
				v &= (1ull << (8u * n)) - 1ull;
			acc |= v << sh;
			if (k + n >= 8) {
				stg64(dst + (op & ~7u), acc);  // a complete word lies below op + n <= cap
				acc = k ? v >> (64u - sh) : 0ull;
			}
			op += n;
			rem -= n;
		}
		__syncwarp(full);
	}
}

}  // namespace csb

using namespace csb;

extern "C" int csb_launch_decompress_lane(const struct csb_decompress_args *a, csb_stream_t s)
{
	if (a->n_blocks == 0)
		return 0;
	DeviceInfo di;
	int e = device_info(&di);
	if (e)
		return e;
	LaneParams p;
	p.a = *a;
	uint32_t *counter = a->counter ? a->counter : next_counter();
	if (!counter)
		return (int)cudaErrorMemoryAllocation;
	cudaError_t ce = cudaMemsetAsync(counter, 0, sizeof(uint32_t), s);
	if (ce != cudaSuccess)
		return (int)ce;
	p.counter = counter;
	long ctas = ((long)a->n_blocks + kLaneThreads - 1) / kLaneThreads;
	const long max_ctas = (long)di.sm_count * 6;  // persistent: lanes claim blocks until none is left
	if (ctas > max_ctas)
		ctas = max_ctas;
	decompress_lane_kernel<<<(int)ctas, kLaneThreads, 0, s>>>(p);
	count_launch();
	return (int)cudaGetLastError();
}
