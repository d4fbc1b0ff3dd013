// compress_kernel.cu -- batched Snappy fragment compressor for sm_100a.
//
// Produces, for every block, exactly the bytes csnappy_compress_fragment produces
// (/root/reference/csnappy_compress.c:469-606) for the same input and table size.
// The reference's greedy parse is a serial dependency chain (every probe reads and
// overwrites a hash slot), so parallelism comes from three places only:
//   * blocks are independent: one GROUP of G lanes per block, ~18 blocks resident per SM
//     (bounded by shared memory: u16 hash table of 1<<wm bytes + the staged input);
//   * inside a scan, the next G probe positions are data independent (they depend only on
//     the skip counter, csnappy_compress.c:535-542), so G lanes probe them speculatively,
//     resolve same-hash collisions with match.any, find the first hit with ballot/ffs and
//     commit only the table writes the serial code would have made;
//   * match extension compares 4*G bytes per step, copy tags of long matches and literal
//     payloads are emitted by all lanes.
// Output is staged in a small shared-memory ring and flushed to HBM in 16-byte units.
#include "device_common.cuh"
#include "kernels.h"

namespace csb {

constexpr uint32_t kRing = 512;	     // output staging ring per group (bytes, power of two)
constexpr uint32_t kAppendMax = 256; // largest single append into the ring
constexpr uint32_t kInPad = 32;	     // slack after the staged input for 4-byte over-reads
constexpr uint32_t kTailMargin = 15; // kInputMarginBytes, csnappy_compress.c:468

struct CompressParams {
	csb_compress_args a;
	uint32_t *counter;     // dynamic block claim; NULL => static striding
	uint32_t table_bytes;  // 1 << wm
	uint32_t in_area;      // bytes reserved for the staged input (multiple of 16)
	uint32_t group_smem;   // table_bytes + in_area + kRing
	uint32_t groups;       // groups per CTA that own shared memory
};

template <int G>
struct Emitter {
	const Group<G> &g;
	uint8_t *ring;
	uint8_t *dst;	   // block's output slot in HBM
	uint32_t op;	   // bytes emitted so far
	uint32_t flushed;  // bytes already written to HBM (multiple of 16 while vec is true)
	bool vec;	   // dst is 16-byte aligned

	__device__ __forceinline__ Emitter(const Group<G> &g_, uint8_t *ring_, uint8_t *dst_)
		: g(g_), ring(ring_), dst(dst_), op(0), flushed(0)
	{
		vec = (reinterpret_cast<uintptr_t>(dst_) & 15u) == 0;
	}

	// write out every complete 16-byte unit (vec) or every pending byte (!vec)
	__device__ __forceinline__ void flush()
	{
		g.sync();
		if (vec) {
			const uint32_t units = (op - flushed) >> 4;
			for (uint32_t u = g.lane; u < units; u += G) {
				const uint32_t at = flushed + (u << 4);
				stg_stream(reinterpret_cast<uint4 *>(dst + at),
					   *reinterpret_cast<const uint4 *>(ring + (at & (kRing - 1))));
			}
			flushed += units << 4;
		} else {
			for (uint32_t at = flushed + g.lane; at < op; at += G)
				dst[at] = ring[at & (kRing - 1)];
			flushed = op;
		}
		g.sync();
	}

	__device__ __forceinline__ void reserve(uint32_t k)
	{
		if (op - flushed + k > kRing)
			flush();
	}

	__device__ __forceinline__ void finish()
	{
		flush();
		for (uint32_t at = flushed + g.lane; at < op; at += G)
			dst[at] = ring[at & (kRing - 1)];
	}

	__device__ __forceinline__ void put(uint32_t at, uint32_t byte) { ring[at & (kRing - 1)] = (uint8_t)byte; }

	// literal tag + payload, csnappy_compress.c:332-371
	__device__ __forceinline__ void literal(const uint8_t *sin, uint32_t src, uint32_t len)
	{
		const uint32_t v = len - 1;
		uint32_t first = len < kAppendMax - 3 ? len : kAppendMax - 3;
		reserve(3 + first);
		uint32_t hb;
		if (v < 60) {
			hb = 1;
			if (g.lane == 0)
				put(op, v << 2);
		} else if (v < 256) {
			hb = 2;
			if (g.lane == 0) {
				put(op, 60u << 2);
				put(op + 1, v);
			}
		} else {
			hb = 3;
			if (g.lane == 0) {
				put(op, 61u << 2);
				put(op + 1, v & 0xff);
				put(op + 2, v >> 8);
			}
		}
		op += hb;
		for (;;) {
			for (uint32_t i = g.lane; i < first; i += G)
				put(op + i, sin[src + i]);
			op += first;
			src += first;
			len -= first;
			if (len == 0)
				break;
			first = len < kAppendMax ? len : kAppendMax;
			reserve(first);
		}
	}

	// one copy element at ring position `at`; returns its size (2 or 3), csnappy_compress.c:373-393
	__device__ __forceinline__ uint32_t copy_piece(uint32_t at, uint32_t offset, uint32_t len, bool write)
	{
		if (len < 12 && offset < 2048) {
			if (write) {
				put(at, 1u | ((len - 4) << 2) | ((offset >> 8) << 5));
				put(at + 1, offset & 0xff);
			}
			return 2;
		}
		if (write) {
			put(at, 2u | ((len - 1) << 2));
			put(at + 1, offset & 0xff);
			put(at + 2, offset >> 8);
		}
		return 3;
	}

	// split rule of csnappy_compress.c:395-415: 64s while len >= 68, one 60 if len > 64, the rest
	__device__ __forceinline__ void copy(uint32_t offset, uint32_t len)
	{
		if (len <= 64) {
			reserve(3);
			op += copy_piece(op, offset, len, g.lane == 0);
			return;
		}
		const uint32_t q = len >= 68 ? (len - 68) / 64 + 1 : 0;
		uint32_t rem = len - 64 * q;  // 4..67
		const uint32_t n60 = rem > 64 ? 1 : 0;
		rem -= 60 * n60;
		const uint32_t lead = q + n60;	// all 3-byte copy-2 elements
		for (uint32_t t0 = 0; t0 < lead; t0 += G) {
			reserve(3 * G);
			const uint32_t t = t0 + g.lane;
			if (t < lead)
				copy_piece(op + 3 * g.lane, offset, t < q ? 64 : 60, true);
			const uint32_t done = lead - t0 < (uint32_t)G ? lead - t0 : (uint32_t)G;
			op += 3 * done;
		}
		reserve(3);
		op += copy_piece(op, offset, rem, g.lane == 0);
	}
};

template <int G>
__device__ __forceinline__ void compress_block(const Group<G> &g, const CompressParams &p, uint32_t blk,
					       uint16_t *tab, uint8_t *sin, uint8_t *ring)
{
	const csb_compress_args &a = p.a;
	const uint64_t in_at = a.in_off ? a.in_off[blk] : (uint64_t)blk * a.in_stride;
	uint32_t n = a.in_len ? a.in_len[blk] : a.uniform_len;
	if (a.total_len) {
		const uint64_t left = a.total_len > in_at ? a.total_len - in_at : 0;
		if (left < n)
			n = (uint32_t)left;
	}
	if (n > CSB_FRAGMENT_MAX)
		n = CSB_FRAGMENT_MAX;  // REQUIRES of the reference (csnappy.h:38); launcher rejects uniform_len above it
	const uint8_t *src = a.in + in_at;
	Emitter<G> em(g, ring, a.out + (uint64_t)blk * a.out_stride);

	// table size for this block (csnappy_compress.c:638-646 when SHRINK_TABLE is set)
	int ws = a.wm;
	if ((a.flags & 1u) && n < CSB_FRAGMENT_MAX) {
		for (ws = 9; ws < a.wm; ++ws)
			if ((1u << (ws - 1)) >= n)
				break;
	}
	const int shift = 33 - ws;

	g.sync();  // previous block's readers of sin/tab/ring are done
	load_block_to_smem<G>(g, sin, src, n);
	if (n >= kTailMargin) {	 // zero the table, csnappy_compress.c:501
		uint4 *t4 = reinterpret_cast<uint4 *>(tab);
		const uint32_t nv = (1u << ws) >> 4;
		const uint4 z = make_uint4(0, 0, 0, 0);
		for (uint32_t i = g.lane; i < nv; i += G)
			t4[i] = z;
	}
	g.sync();

	uint32_t next_emit = 0;
	if (n >= kTailMargin) {

		const uint32_t ip_limit = n - kTailMargin;
		uint32_t ip = 1;
		for (;;) {
			// ---- scan: G speculative probes per step (csnappy_compress.c:535-552) ----
			uint32_t base = ip, j = 0, cand = 0;
			bool found = false;
			for (;;) {
				const uint32_t stride = (32u + j) >> 5;
				const uint32_t pp = base + g.lane * stride;
				const bool valid = pp + stride <= ip_limit;
				uint32_t bytes = 0, h = 0x80000000u | g.lane, cb = 1;
				if (valid) {
					bytes = lds32u(sin, pp);
					h = (bytes * kHashMul) >> shift;
				}
				const unsigned same = g.match(h);
				const unsigned lower = same & ((1u << g.lane) - 1u);
				if (valid) {
					cand = lower ? base + (31 - __clz(lower)) * stride : tab[h];
					cb = lds32u(sin, cand);
				}
				const unsigned hits = g.ballot(valid && cb == bytes);
				const unsigned valids = g.ballot(valid);
				const unsigned f = hits ? __ffs(hits) - 1 : (unsigned)G;
				if (valid && g.lane <= f) {
					// highest lane of an equal-hash run (up to the hit) owns the slot
					const unsigned above = (same >> g.lane) >> 1;
					const unsigned span = f - g.lane;
					const unsigned rivals = span >= 32 ? above : (above & ((1u << span) - 1u));
					if (!rivals)
						tab[h] = (uint16_t)pp;
				}
				g.sync();
				if (hits) {
					ip = base + f * stride;
					cand = g.bcast(cand, (int)f);
					found = true;
					break;
				}
				if (valids != ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)))
					break;	// ran into ip_limit without a hit
				base += G * stride;
				j += G;
			}
			if (!found)
				break;

			// ---- emit literal, then copies while the next position matches (:560-594) ----
			em.literal(sin, next_emit, ip - next_emit);
			bool again;
			do {
				// match extension: 4*G bytes per step, bounded by n (csnappy_compress.c:252-295)
				uint32_t m = 4;
				for (;;) {
					const uint32_t at = ip + m + 4 * g.lane;
					uint32_t mb = 0;
					if (at < n) {
						const uint32_t room = n - at;
						const uint32_t x = lds32u(sin, cand + m + 4 * g.lane) ^ lds32u(sin, at);
						mb = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
						if (mb > room)
							mb = room;
					}
					const unsigned stop = g.ballot(mb < 4);
					if (stop) {
						const int fl = __ffs(stop) - 1;
						m += 4 * fl + g.bcast(mb, fl);
						break;
					}
					m += 4 * G;
				}
				em.copy(ip - cand, m);
				ip += m;
				next_emit = ip;
				if (ip >= ip_limit)
					goto remainder;
				// insert ip-1, then probe ip (csnappy_compress.c:587-593); every lane does it
				// redundantly: identical values to identical addresses
				const uint32_t prev = lds32u(sin, ip - 1);
				tab[(prev * kHashMul) >> shift] = (uint16_t)(ip - 1);
				g.sync();
				const uint32_t cur = lds32u(sin, ip);
				const uint32_t hc = (cur * kHashMul) >> shift;
				cand = tab[hc];
				g.sync();
				tab[hc] = (uint16_t)ip;
				again = lds32u(sin, cand) == cur;
				g.sync();
			} while (again);
			ip += 1;
		}
	}
remainder:
	if (next_emit < n)
		em.literal(sin, next_emit, n - next_emit);
	em.finish();
	if (g.lane == 0)
		a.out_len[blk] = em.op;
}

template <int G>
__global__ void __launch_bounds__(1024) compress_kernel(const CompressParams p)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const Group<G> g;
	const uint32_t groups_per_cta = p.groups;
	const uint32_t gid = threadIdx.x / G;
	if (gid >= p.groups)
		return;	 // padding lanes of the last warp (no block-wide barriers in this kernel)
	uint8_t *gs = smem + (size_t)gid * p.group_smem;
	uint16_t *tab = reinterpret_cast<uint16_t *>(gs);
	uint8_t *sin = gs + p.table_bytes;
	uint8_t *ring = sin + p.in_area;

	if (p.counter) {
		for (;;) {
			uint32_t blk = 0;
			if (g.lane == 0)
				blk = atomicAdd(p.counter, 1u);
			blk = g.bcast(blk, 0);
			if (blk >= p.a.n_blocks)
				break;
			compress_block<G>(g, p, blk, tab, sin, ring);
		}
	} else {
		const uint32_t total = gridDim.x * groups_per_cta;
		for (uint32_t blk = blockIdx.x * groups_per_cta + gid; blk < p.a.n_blocks; blk += total)
			compress_block<G>(g, p, blk, tab, sin, ring);
	}
}

}  // namespace csb

using namespace csb;

template <int G>
static int launch_compress_g(const CompressParams &p, int threads, int ctas, size_t smem, cudaStream_t s)
{
	cudaError_t e = cudaFuncSetAttribute(compress_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess)
		return (int)e;
	compress_kernel<G><<<ctas, threads, smem, s>>>(p);
	count_launch();
	return (int)cudaGetLastError();
}

extern "C" int csb_launch_compress(const struct csb_compress_args *a, csb_stream_t s)
{
	if (a->n_blocks == 0)
		return 0;
	DeviceInfo di;
	int e = device_info(&di);
	if (e)
		return e;

	CompressParams p;
	p.a = *a;
	p.table_bytes = 1u << a->wm;
	uint32_t in_cap = a->uniform_len;
	if (a->in_len)	// per-block lengths: bounded by the stride when strided, else by the format
		in_cap = (!a->in_off && a->in_stride && a->in_stride < CSB_FRAGMENT_MAX) ? (uint32_t)a->in_stride : CSB_FRAGMENT_MAX;
	if (in_cap > CSB_FRAGMENT_MAX)
		in_cap = CSB_FRAGMENT_MAX;
	p.in_area = ((in_cap + 15u) & ~15u) + kInPad;
	p.group_smem = p.table_bytes + p.in_area + kRing;

	const int G = a->lanes ? a->lanes : 32;
	const int ctas_per_sm = a->ctas_per_sm > 0 ? a->ctas_per_sm : 1;
	// shared memory per CTA: the SM's carve-out divided among resident CTAs (1 KiB reserved each)
	long budget = (long)di.smem_per_sm / ctas_per_sm - 1024;
	if (budget > di.smem_per_block_optin)
		budget = di.smem_per_block_optin;
	int groups = (int)(budget / p.group_smem);
	const int max_groups = 1024 / G;
	if (groups > max_groups)
		groups = max_groups;
	if (groups < 1)
		return (int)cudaErrorInvalidConfiguration;
	p.groups = (uint32_t)groups;
	const int threads = (groups * G + 31) / 32 * 32;  // whole warps; surplus lanes exit at once
	const size_t smem = (size_t)groups * p.group_smem;

	long want = ((long)a->n_blocks + groups - 1) / groups;
	long ctas = (long)di.sm_count * ctas_per_sm;
	if (ctas > want)
		ctas = want;

	p.counter = nullptr;
	uint32_t *counter = nullptr;
	if ((long)a->n_blocks > ctas * groups) {
		cudaError_t ce = cudaMallocAsync((void **)&counter, sizeof(uint32_t), s);
		if (ce != cudaSuccess)
			return (int)ce;
		ce = cudaMemsetAsync(counter, 0, sizeof(uint32_t), s);
		if (ce != cudaSuccess)
			return (int)ce;
		p.counter = counter;
	}
	switch (G) {
	case 32: e = launch_compress_g<32>(p, threads, (int)ctas, smem, s); break;
	case 16: e = launch_compress_g<16>(p, threads, (int)ctas, smem, s); break;
	case 8: e = launch_compress_g<8>(p, threads, (int)ctas, smem, s); break;
	default: e = (int)cudaErrorInvalidValue; break;
	}
	if (counter)
		cudaFreeAsync(counter, s);
	return e;
}
