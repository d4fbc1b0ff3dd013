// decompress_lane_kernel.cu -- batched Snappy raw-stream decoder, ONE LANE PER BLOCK.
//
// Second decoder family next to decompress_kernel.cu, for blocks that do not fit shared memory in useful
// numbers (32 KiB fragments and anything larger: 3 staged fragments per SM, or 26 warp-owned blocks on
// the global path) and for very large batches of small blocks.  A block's tag stream is a serial chain
// whose step costs an L2 round trip once the block lives in global memory; the only way to hide that
// is MANY chains, and a lane is the cheapest owner a chain can have: 1280 blocks in flight per SM
// instead of 26.  Each lane runs the reference's loop (csnappy_decompress.c:345-382) -- tag, length,
// checks, copy -- on its own block:
//
//   * ONE load path for the tag stream, literal payloads and back-references: the two aligned 16-byte
//     vectors around the (arbitrarily aligned) source and a funnel shift; output through a 16-byte
//     accumulator flushed with aligned 16-byte stores, so every memory instruction moves 16 bytes per
//     lane whatever the alignment of the tag, and most tags (mean length 12-14 bytes) need ONE move;
//   * back-references read the lane's own earlier stores (same thread, same address: ordered); the
//     not yet stored tail of the accumulator is flushed first when a copy reaches into it
//     (offset < 31), and offsets 1, 2, 4 and 8 are expanded in registers (pattern fill);
//   * the loop is FLAT and WARP-UNIFORM: one iteration = claim (free lanes take the next block from
//     the global counter) / tag (lanes whose tag is used up decode the next, with selects instead of
//     branches between the tag kinds) / bulk (long literals are copied by the whole warp) / move
//     (every lane moves up to 16 bytes), with a reconvergence point between the phases -- without
//     them the lanes of a warp drift apart and run the loop body one lane at a time (measured: 6.7
//     of 32 lanes active per instruction, 121 GB/s instead of 272 on 32 KiB fragments).
//
// Semantics per block are those of decompress_kernel.cu: first failing tag in stream order decides;
// literal: input shortage (-5) before space (-3); copy: offset validity (-5) before space (-3); end of
// input at a tag boundary is success; a tag header cut off by the end of input is -5 (defined here,
// SURVEY.md 0.5); nothing is ever written at or past the block's capacity.
// Needs 16-byte aligned output slots (out, out_stride).  Bound by instruction issue and the LSU (a
// fully divergent access costs ~2 cycles per lane), not by HBM; see DESIGN.md 4.3.
#include "device_common.cuh"
#include "kernels.h"

namespace csb {

constexpr int L_OK = 0, L_HEADER_BAD = -1, L_OUTPUT_INSUF = -2, L_OUTPUT_OVERRUN = -3, L_DATA_MALFORMED = -5;
constexpr int kLaneThreads = 256;
constexpr int kLaneWarpsDefault = 32;
constexpr uint32_t kBulkMin = 528;  // literals at least this long leave the lane loop for a warp-wide copy

struct U128 {
	uint64_t lo, hi;
};

__device__ __forceinline__ U128 ldg128(uintptr_t a)
{
	U128 v;
	asm volatile("ld.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "l"(a) : "memory");
	return v;
}
__device__ __forceinline__ void stg128(uintptr_t a, uint64_t lo, uint64_t hi)
{
	asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(a), "l"(lo), "l"(hi) : "memory");
}
// (hi:lo) >> 8 r for r in 0..7, without the undefined shift by 64
__device__ __forceinline__ uint64_t shr_pair(uint64_t lo, uint64_t hi, uint32_t r)
{
	return (lo >> (8u * r)) | ((hi << 1) << (63u - 8u * r));
}

// `need` (1..16) bytes at byte address a, little endian, without touching a byte at or past `lim`.
// One code path for the tag stream, literal payloads (lim = end of the block) and back-references
// (lim = end of the output slot): the two aligned vectors around a and a funnel shift.  A vector that
// starts before the block still lies inside the caller's allocation (allocations are 256-byte aligned).
__device__ __forceinline__ U128 load16(uintptr_t a, uint32_t need, uintptr_t lim)
{
	const uintptr_t w = a & ~(uintptr_t)15;
	const uint32_t k = (uint32_t)a & 15u;
	const bool two = k + need > 16;
	U128 r;
	if (w + (two ? 32 : 16) <= lim) {
		const U128 A = ldg128(w);
		U128 B;
		B.lo = B.hi = 0;
		if (two)
			B = ldg128(w + 16);
		const bool q = k >= 8;
		const uint32_t s = k & 7u;
		const uint64_t x0 = q ? A.hi : A.lo, x1 = q ? B.lo : A.hi, x2 = q ? B.hi : B.lo;
		r.lo = shr_pair(x0, x1, s);
		r.hi = shr_pair(x1, x2, s);
		return r;
	}
	r.lo = r.hi = 0;  // the last bytes of the block / slot: byte by byte
	for (uint32_t i = 0; i < need && a + i < lim; ++i) {
		const uint64_t b = *reinterpret_cast<const volatile uint8_t *>(a + i);
		if (i < 8)
			r.lo |= b << (8 * i);
		else
			r.hi |= b << (8 * (i - 8));
	}
	return r;
}

// ---- per-lane input ring in shared memory ------------------------------------------------------------
// Every lane reads its own compressed block strictly forwards, 1..21 bytes per step.  Through L1 / L2 that is a
// dependent long-latency load per tag on a cache thousands of lanes compete for (measured: the tag load alone was
// 26-31 % of all stall samples, DRAM reads 5-7x the input).  Instead each lane owns a small ring of 64-byte lines
// -- global byte address g lives at ring + (g & (kRing - 1)) -- and fetches ahead with cp.async (LDGSTS, 16 bytes
// x 4, L2 only), so every input byte crosses L2 once and the tag / literal loads become LDS.  One step reads at
// most 36 bytes ahead of its start (5 header + 16 payload bytes through two aligned 16-byte vectors), i.e. the
// line of the read position and the next one.
//   kRing 128 (default): two lines.  Entering a line requests the next one; the step that enters a line is at most
//       20 bytes into it and cannot reach the next, so the request has a whole step to land: the loop waits for
//       all cp.async groups but the newest.  16 KiB per 128-lane CTA: up to 52 warps per SM.
//   kRing 256: four lines, three requested ahead, the loop waits for all but the two newest groups.  More slack,
//       but only 24 warps per SM fit (measured: 405 vs 462 GB/s on URL text pages).
#ifndef CSB_LANE_RING
#define CSB_LANE_RING 128
#endif
constexpr uint32_t kRing = CSB_LANE_RING, kLine = 64, kAhead = kRing / kLine - 1;

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, uintptr_t gsrc)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ring_fetch(uint32_t ring, uintptr_t line)
{
	const uint32_t d = ring + ((uint32_t)line & (kRing - 1));
#pragma unroll
	for (uint32_t i = 0; i < kLine; i += 16)
		cp_async16(d + i, line + i);
}
// The ring's invariant for read position g: the line holding g and the kAhead behind it are requested (lines that
// start at or past `lim` hold no input byte and are never touched).  `have` = line of the previous read position.
// ring_kind: 0 = nothing to do, 1 = the position entered the next line (request one line ahead), 2 = a jump (new block,
// behind a bulk literal): every line is new and needed at once.  cp.async operations are NOT ordered among themselves,
// so before a jump refills the slots nothing may still be in flight towards them: the caller drains (wait_group 0)
// between ring_kind and ring_request whenever a lane of the warp jumps.
__device__ __forceinline__ uint32_t ring_kind(uintptr_t g, uintptr_t lim, uintptr_t have)
{
	const uintptr_t cur = g & ~(uintptr_t)(kLine - 1);
	if (cur == have || cur >= lim)
		return 0;
	return cur == have + kLine ? 1u : 2u;
}
__device__ __forceinline__ void ring_request(uint32_t ring, uintptr_t g, uintptr_t lim, uintptr_t &have, uint32_t kind)
{
	const uintptr_t cur = g & ~(uintptr_t)(kLine - 1);
	have = cur;
	if (kind == 2) {
#pragma unroll
		for (uint32_t i = 0; i < kAhead; ++i)
			if (cur + i * kLine < lim)
				ring_fetch(ring, cur + i * kLine);
	}
	if (cur + kAhead * kLine < lim)
		ring_fetch(ring, cur + kAhead * kLine);
}
// `need` (1..16) bytes at global byte address g out of the ring, little endian
__device__ __forceinline__ U128 ring_load16(uint32_t ring, uintptr_t g, uint32_t need)
{
	const uint32_t w = (uint32_t)g & ~15u, k = (uint32_t)g & 15u;
	const uint4 a = lds_v4(ring + (w & (kRing - 1)));
	uint4 b = make_uint4(0, 0, 0, 0);
	if (k + need > 16)
		b = lds_v4(ring + ((w + 16) & (kRing - 1)));
	const uint64_t alo = ((uint64_t)a.y << 32) | a.x, ahi = ((uint64_t)a.w << 32) | a.z;
	const uint64_t blo = ((uint64_t)b.y << 32) | b.x, bhi = ((uint64_t)b.w << 32) | b.z;
	const bool q = k >= 8;
	const uint32_t s = k & 7u;
	const uint64_t x0 = q ? ahi : alo, x1 = q ? blo : ahi, x2 = q ? bhi : blo;
	U128 r;
	r.lo = shr_pair(x0, x1, s);
	r.hi = shr_pair(x1, x2, s);
	return r;
}

struct LaneParams {
	csb_decompress_args a;
	uint32_t *counter;
};

enum : uint32_t { M_LIT = 0, M_COPY = 1, M_PATTERN = 2, M_NEAR = 3 };

// make the accumulator's bytes [op & ~15, op) visible to this lane's loads
__device__ __forceinline__ void flush_partial(uintptr_t dst, uint32_t cap, uint32_t op, uint64_t acc0, uint64_t acc1)
{
	const uint32_t k = op & 15u, w = op & ~15u;
	if (!k)
		return;
	if (w + 16 <= cap) {
		stg128(dst + w, acc0, acc1);  // upper bytes: zeros inside the capacity, rewritten later
	} else {
		for (uint32_t i = 0; i < k; ++i)
			*reinterpret_cast<volatile uint8_t *>(dst + w + i) = (uint8_t)((i < 8 ? acc0 : acc1) >> (8 * (i & 7u)));
	}
}

__global__ void __launch_bounds__(kLaneThreads) decompress_lane_kernel(const LaneParams p)
{
	const csb_decompress_args &a = p.a;
	const unsigned full = 0xffffffffu;
	const uint32_t lane = threadIdx.x & 31u;
	bool have = false, more = true;
	uint32_t blk = 0, cap = 0, ip = 0, op = 0, rem = 0, mode = M_LIT, off = 0;
	uintptr_t src = 0, src_end = 0, dst = 0;
	uint64_t acc0 = 0, acc1 = 0, pat = 0;  // acc: bytes [op & ~15, op) of the output, not yet stored
	uint64_t prv0 = 0, prv1 = 0;  // the 16 bytes in front of the accumulator, as this lane stored them ...
	bool prv_ok = false;          // ... unless the warp-wide bulk copy wrote them
	extern __shared__ __align__(128) uint8_t lane_smem[];
	const uint32_t ring = smem_u32(lane_smem) + threadIdx.x * kRing;
	uintptr_t have_line = 1;  // never a line address

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
					acc0 = acc1 = 0;
					prv_ok = false;
					mode = M_LIT;
					if ((a.flags & 4u) && ilen == cap)
						rem = ilen;  // stored block (block_compressor.c:378): the whole input is one literal payload
					have_line = 1;
				}
			}
		}
		// ---- INPUT RING: request what the new read position needs; wait for what this step reads ----
		const uint32_t kind = have ? ring_kind(src + ip, src_end, have_line) : 0u;
		const bool jump = __any_sync(full, kind == 2u);
		if (jump)
			asm volatile("cp.async.wait_group 0;" ::: "memory");  // drain before any slot is refilled
		if (kind)
			ring_request(ring, src + ip, src_end, have_line, kind);
		asm volatile("cp.async.commit_group;" ::: "memory");
		if (jump) {
			asm volatile("cp.async.wait_group 0;" ::: "memory");
		} else {
			if (kAhead >= 3)
				asm volatile("cp.async.wait_group 2;" ::: "memory");
			else
				asm volatile("cp.async.wait_group 1;" ::: "memory");
		}
		__syncwarp(full);
		if (!__any_sync(full, have || more))
			break;

		// ---- TAG: next tag of the reference's loop (csnappy_decompress.c:345-382) ----
		if (have && rem == 0) {
			const uint32_t ilen = (uint32_t)(src_end - src);
			int rc = L_OK;
			const bool fin = ip >= ilen;  // end of input at a tag boundary
			if (!fin) {
				const uint32_t left = ilen - ip;
				const uint64_t x = ring_load16(ring, src + ip, 8u).lo;  // bytes past the end are never used (hdr <= left)
				const uint32_t tag = (uint32_t)x & 0xffu, kind = tag & 3u, lf = tag >> 2;
				const uint32_t extra = (uint32_t)(x >> 8);
				const bool lit = kind == 0, longlit = lit && lf >= 60;
				// header bytes: literal 1 (+ 1..4 length bytes), copy 2 / 3 / 5
				const uint32_t hdr = lit ? (longlit ? lf - 58u : 1u) : (kind == 1 ? 2u : (kind == 2 ? 3u : 5u));
				const uint32_t nb8 = 8u * (lf - 59u);  // bits of the long literal's length field
				const uint32_t lv = longlit ? (extra & (nb8 >= 32 ? 0xffffffffu : ((1u << (nb8 & 31u)) - 1u))) : lf;
				// literal: length field + 1 (0xffffffff wraps to a zero-length literal, csnappy_decompress.c:370)
				const uint32_t len = kind == 1 ? ((lf & 7u) + 4u) : (lv + 1u);
				off = kind == 1 ? (((tag >> 5) << 8) | (extra & 0xffu)) : (kind == 2 ? (extra & 0xffffu) : extra);
				const uint32_t ip2 = ip + hdr, room = cap - op;
				// check order of the reference: header / length bytes present (UB there, -5 here); literal: input
				// shortage (a length of 2^31 or more passes its signed test and fails on space, :374) then space;
				// copy: offset == 0 or > produced (:302) then space
				const bool cut = left < hdr;
				const bool bad = lit ? ((int32_t)len >= 0 ? (ilen - ip2 < len) : (room >= len)) : (off - 1u >= op);
				rc = (cut || bad) ? L_DATA_MALFORMED : (room < len ? L_OUTPUT_OVERRUN : L_OK);
				ip = ip2;
				rem = len;
				// copies: off >= 31 never reaches into the unstored accumulator
				mode = lit ? M_LIT : (off >= 31 ? M_COPY : ((off == 1 || off == 2 || off == 4 || off == 8) ? M_PATTERN : M_NEAR));
			}
			if (rc != L_OK || fin) {
				if (rc == L_OK)
					flush_partial(dst, cap, op, acc0, acc1);
				a.status[blk] = rc;
				a.out_len[blk] = rc == L_OK ? op : 0u;
				have = false;
				rem = 0;
			} else if (mode == M_PATTERN && rem) {
				// the last `off` (1, 2, 4, 8) bytes of the output.  They are still in registers -- the accumulator and the
				// vector stored before it -- so a run of pattern copies (a zero page is 64 of them) never waits for its
				// own stores to come back from L2; only behind a bulk copy they have to be read from memory.
				const uint32_t k = op & 15u;
				if (k >= off || prv_ok) {
					// 8 bytes ending at op: from (prv1, acc0) for k < 8, from (acc0, acc1) otherwise
					const uint64_t l8 = k < 8 ? shr_pair(prv1, acc0, k) : shr_pair(acc0, acc1, k - 8u);
					pat = l8 >> (8u * (8u - off));
				} else {
					flush_partial(dst, cap, op, acc0, acc1);
					pat = load16(dst + op - off, off, dst + cap).lo;
				}
				if (off == 1)
					pat = (pat & 0xffull) * 0x0101010101010101ull;
				else if (off == 2)
					pat = (pat & 0xffffull) * 0x0001000100010001ull;
				else if (off == 4)
					pat = (pat & 0xffffffffull) * 0x0000000100000001ull;
			}
		}
		__syncwarp(full);

		// ---- BULK: a long literal (incompressible data, stored blocks) is copied by the whole warp ----
		// once the owner's output position is 16-byte aligned (the MOVE phase below aligns it first)
		bool moved = false;
		unsigned bulk = __ballot_sync(full, have && mode == M_LIT && rem >= kBulkMin && (op & 15u) == 0);
		while (bulk) {
			const int o = __ffs(bulk) - 1;
			bulk &= bulk - 1;
			const uintptr_t s_ = __shfl_sync(full, (unsigned long long)(src + ip), o);
			const uintptr_t d_ = __shfl_sync(full, (unsigned long long)(dst + op), o);
			const uintptr_t lim = __shfl_sync(full, (unsigned long long)src_end, o);
			const uint32_t n_ = __shfl_sync(full, rem & ~15u, o);
			for (uint32_t i = 16 * lane; i < n_; i += 512) {
				const U128 v = load16(s_ + i, 16, lim);
				stg128(d_ + i, v.lo, v.hi);
			}
			if ((int)lane == o) {
				ip += n_;
				op += n_;
				rem -= n_;
				prv_ok = false;  // other lanes wrote the bytes in front of the accumulator
				moved = true;    // the ring does not hold the new read position yet
			}
		}
		__syncwarp(full);

		// ---- MOVE: up to 16 bytes of the current tag ----
		if (have && rem && !moved) {
			uint32_t n = rem < 16 ? rem : 16;
			if (mode == M_LIT && rem >= kBulkMin)
				n = 16 - (op & 15u);  // align the output for the bulk copy (16 when it already is: never here)
			U128 v;
			v.lo = v.hi = pat;
			bool from_regs = false;
			if (mode == M_NEAR) {
				// offsets 3, 5..7, 9..30: rounds of at most `off` bytes read only what is already there.  The last
				// 16 + (op & 15) bytes of the output are still in registers (the vector stored last + the accumulator):
				// a source inside that window is cut out of them (9 % of the copies of URL text; a store followed by a
				// load of the same bytes is a round trip through L2 that the whole warp waits for); a source further
				// back, or behind a bulk copy, needs the accumulator's bytes in memory first.
				if (n > off)
					n = off;
				const uint32_t k = op & 15u;
				from_regs = off <= k || (prv_ok && off <= 16u + k);
				if (from_regs) {
					const uint32_t b = 16u + k - off, w = b >> 3, sft = b & 7u;  // byte offset into prv0 prv1 acc0 acc1
					const uint64_t x0 = w == 0 ? prv0 : (w == 1 ? prv1 : (w == 2 ? acc0 : acc1));
					const uint64_t x1 = w == 0 ? prv1 : (w == 1 ? acc0 : (w == 2 ? acc1 : 0ull));
					const uint64_t x2 = w == 0 ? acc0 : (w == 1 ? acc1 : 0ull);
					v.lo = shr_pair(x0, x1, sft);
					v.hi = shr_pair(x1, x2, sft);
				} else {
					flush_partial(dst, cap, op, acc0, acc1);
				}
			}
			if (mode == M_LIT) {
				v = ring_load16(ring, src + ip, n);
				ip += n;
			} else if (mode != M_PATTERN && !from_regs) {
				v = load16(dst + op - off, n, dst + cap);
			}
			// append the n low bytes of v (op + n <= cap was checked with the tag)
			if (n < 16) {
				const uint64_t m = ~0ull >> (8u * (8u - (n & 7u)) & 63u);  // low (n & 7) bytes; all ones for n & 7 == 0
				if (n < 8) {
					v.lo &= m;
					v.hi = 0;
				} else if (n > 8) {
					v.hi &= m;
				} else {
					v.hi = 0;
				}
			}
			const uint32_t k = op & 15u, s = k & 7u;
			// v << 8 k as four words y0..y3 (k = 8 q + s)
			const uint64_t t0 = v.lo << (8u * s);
			const uint64_t t1 = (v.hi << (8u * s)) | ((v.lo >> 1) >> (63u - 8u * s));
			const uint64_t t2 = (v.hi >> 1) >> (63u - 8u * s);
			const bool q = k >= 8;
			acc0 |= q ? 0ull : t0;
			acc1 |= q ? t0 : t1;
			if (k + n >= 16) {
				stg128(dst + (op & ~15u), acc0, acc1);	// a complete vector lies below op + n <= cap
				prv0 = acc0;
				prv1 = acc1;
				prv_ok = true;
				acc0 = q ? t1 : t2;
				acc1 = q ? t2 : 0ull;
			}
			op += n;
			rem -= n;
			// A pattern run (offset 1, 2, 4, 8: zero pages, RLE) repeats every 16 bytes: once a full 16-byte move
			// has gone through the accumulator, every further vector of the run is the same value -- the k bytes
			// left over plus the first 16 - k bytes of the pattern -- and the accumulator does not change.
			if (mode == M_PATTERN && n == 16 && rem >= 16) {
				const uint64_t v0 = acc0 | (q ? 0ull : t0), v1 = acc1 | (q ? t0 : t1);
				uint32_t more = rem >> 4;
				if (more > 15)
					more = 15;
				const uintptr_t at = dst + (op & ~15u);  // the last of these vectors ends at or below op + rem <= cap
				for (uint32_t e = 0; e < more; ++e)
					stg128(at + 16 * e, v0, v1);
				if (more) {
					prv0 = v0;
					prv1 = v1;
				}
				op += 16 * more;
				rem -= 16 * more;
			}
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
	// Persistent: lanes claim blocks until none is left.  Blocks in flight per SM = 32 * warps: more lanes hide more
	// latency, but every lane's output page competes for L2, where the back-references read it back (measured on
	// 4 KiB text pages, 1 Mi blocks: 16 warps 324, 24 warps 430, 32 warps 462, 48 warps 461 GB/s; 32 KiB fragments,
	// 512 Ki blocks: 16 warps 226, 24 warps 273, 32 warps 282).
	int warps = a->lane_warps > 0 ? a->lane_warps : (a->ctas_per_sm > 0 ? 8 * a->ctas_per_sm : kLaneWarpsDefault);
	if (a->lane_warps <= 0 && a->ctas_per_sm <= 0) {
		// a batch that fits the machine in ONE round of at most 40 warps per SM runs as one round: no second-round tail
		const long need = ((long)a->n_blocks + 32L * di.sm_count - 1) / (32L * di.sm_count);
		if (need > warps && need <= 40)
			warps = (int)need;
	}
	const int threads = 128;  // 4 warps per CTA: fine granularity for the warps-per-SM knob
	long ctas = ((long)a->n_blocks + threads - 1) / threads;
	const long max_ctas = ((long)di.sm_count * warps + 3) / 4;
	if (ctas > max_ctas)
		ctas = max_ctas;
	const size_t smem = (size_t)threads * kRing;
	const int fit = (int)((di.smem_per_sm / (smem + 1024)) * (threads / 32));  // warps per SM the rings leave room for
	if (warps > fit) {
		warps = fit;
		ctas = ((long)a->n_blocks + threads - 1) / threads;
		const long mc = ((long)di.sm_count * warps + 3) / 4;
		if (ctas > mc)
			ctas = mc;
	}
	decompress_lane_kernel<<<(int)ctas, threads, smem, s>>>(p);
	count_launch();
	return (int)cudaGetLastError();
}
