// decompress_kernel.cu -- batched Snappy raw-stream decoder for sm_100a.
//
// Per block: the semantics of csnappy_decompress_noheader
// (/root/reference/csnappy_decompress.c:319-387) -- first failing tag in stream order
// decides the code; literal: input shortage (-5) before space (-3); copy: offset validity
// (-5) before space (-3); end of input at a tag boundary is success -- or, with
// CSNAPPY_BATCH_WITH_HEADER, of csnappy_decompress (csnappy_decompress.c:394-411).
//
// A GROUP of G lanes owns one block.  Two paths inside one kernel:
//   staged     compressed block and output both fit the group's shared memory: the input is
//              loaded with 16-byte coalesced loads, tags are interpreted against shared
//              memory (back-references never touch HBM), the finished block is written to
//              HBM with 16-byte coalesced stores.  This is the 4 KiB page / 32 KiB fragment path.
//   streaming  anything larger (whole multi-chunk streams through the drop-in API): input and
//              output stay in global memory, back-references are read through L2 (ld.cg).
#include "device_common.cuh"
#include "kernels.h"

namespace csb {

constexpr int E_OK = 0, E_HEADER_BAD = -1, E_OUTPUT_INSUF = -2, E_OUTPUT_OVERRUN = -3, E_DATA_MALFORMED = -5;

struct DecompressParams {
	csb_decompress_args a;
	uint32_t *counter;
	uint32_t in_area;   // staged input capacity (multiple of 16), 0 => streaming only
	uint32_t out_area;  // staged output capacity (multiple of 16)
	uint32_t group_smem;
	uint32_t groups;  // groups per CTA that own shared memory
};

template <bool STAGED>
__device__ __forceinline__ uint32_t load_out(const uint8_t *p)
{
	if (STAGED)
		return *p;
	return __ldcg(p);  // streaming path: written by other lanes of this group, read via L2
}

// Tag interpreter over ib[0..ilen) into ob[0..cap).  Every lane of the group walks the tags
// redundantly (uniform control flow); payload bytes are moved G at a time.
template <int G, bool STAGED>
__device__ __forceinline__ int decode_core(const Group<G> &g, const uint8_t *ib, uint32_t ilen, uint8_t *ob,
					   uint32_t cap, uint32_t *produced_out)
{
	uint32_t pos = 0, produced = 0;
	while (pos < ilen) {
		const uint32_t tag = ib[pos++];
		const uint32_t kind = tag & 3u;
		uint32_t len;
		if (kind == 0) {
			len = (tag >> 2) + 1;
			if (len > 60) {
				const uint32_t nb = len - 60;
				if (ilen - pos < nb)
					return E_DATA_MALFORMED;
				uint32_t v = 0;
				for (uint32_t b = 0; b < nb; ++b)
					v |= (uint32_t)ib[pos + b] << (8 * b);
				pos += nb;
				len = v + 1;  // 0xffffffff wraps to a zero-length literal (csnappy_decompress.c:370)
			}
			if ((int32_t)len >= 0) {
				if (ilen - pos < len)
					return E_DATA_MALFORMED;
			} else if (cap - produced >= len) {
				return E_DATA_MALFORMED;
			}
			if (cap - produced < len)
				return E_OUTPUT_OVERRUN;
			for (uint32_t i = g.lane; i < len; i += G)
				ob[produced + i] = ib[pos + i];
			pos += len;
		} else {
			const uint32_t nb = kind == 3 ? 4u : kind;
			if (ilen - pos < nb)
				return E_DATA_MALFORMED;
			uint32_t off = ib[pos];
			if (nb >= 2)
				off |= (uint32_t)ib[pos + 1] << 8;
			if (nb == 4)
				off |= ((uint32_t)ib[pos + 2] << 16) | ((uint32_t)ib[pos + 3] << 24);
			pos += nb;
			if (kind == 1) {
				len = ((tag >> 2) & 7u) + 4;
				off |= (tag >> 5) << 8;
			} else {
				len = (tag >> 2) + 1;
			}
			if (off - 1u >= produced)  // off == 0 or off > produced, csnappy_decompress.c:302
				return E_DATA_MALFORMED;
			if (cap - produced < len)
				return E_OUTPUT_OVERRUN;
			const uint8_t *from = ob + produced - off;
			if (off >= len) {
				// disjoint: len <= 64, at most two rounds
				for (uint32_t i = g.lane; i < len; i += G)
					ob[produced + i] = load_out<STAGED>(from + i);
			} else if (off >= (uint32_t)G) {
				// overlapping but a round of G bytes only reads what earlier rounds wrote
				for (uint32_t c = 0; c < len; c += G) {
					const uint32_t i = c + g.lane;
					if (i < len)
						ob[produced + i] = load_out<STAGED>(from + i);
					g.sync();
				}
			} else {
				// short period: byte i repeats the pattern ob[produced-off .. produced)
				for (uint32_t i = g.lane; i < len; i += G)
					ob[produced + i] = load_out<STAGED>(from + (i % off));
			}
		}
		produced += len;
		g.sync();
	}
	*produced_out = produced;
	return E_OK;
}

template <int G>
__device__ __forceinline__ void decompress_block(const Group<G> &g, const DecompressParams &p, uint32_t blk,
						 uint8_t *sin, uint8_t *sout)
{
	const csb_decompress_args &a = p.a;
	const uint8_t *src = a.in + (a.in_off ? a.in_off[blk] : (uint64_t)blk * a.in_stride);
	uint32_t ilen = a.in_len[blk];
	uint32_t cap = a.out_cap ? a.out_cap[blk] : a.uniform_cap;
	uint8_t *dst = a.out + (uint64_t)blk * a.out_stride;
	int rc = E_OK;
	uint32_t produced = 0;

	if (a.flags & 2u) {  // varint32 length prefix, csnappy_decompress.c:45-71, 404-409
		uint32_t shift = 0, used = 0, value = 0;
		for (;;) {
			if (shift >= 32 || used == ilen) {
				rc = E_HEADER_BAD;
				break;
			}
			const uint32_t c = src[used++];
			value |= (c & 0x7fu) << shift;
			if (c < 128)
				break;
			shift += 7;
		}
		if (rc == E_OK) {
			if (value > cap)
				rc = E_OUTPUT_INSUF;
			cap = value;
			src += used;
			ilen -= used;
		}
	}

	if (rc == E_OK) {
		if (ilen <= p.in_area && cap <= p.out_area) {
			g.sync();  // previous block's readers of the staging areas are done
			load_block_to_smem<G>(g, sin, src, ilen);
			g.sync();
			rc = decode_core<G, true>(g, sin, ilen, sout, cap, &produced);
			if (rc == E_OK) {
				if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
					const uint32_t nv = produced >> 4;
					const uint4 *s4 = reinterpret_cast<const uint4 *>(sout);
					uint4 *d4 = reinterpret_cast<uint4 *>(dst);
					for (uint32_t i = g.lane; i < nv; i += G)
						stg_stream(d4 + i, s4[i]);
					for (uint32_t i = (nv << 4) + g.lane; i < produced; i += G)
						dst[i] = sout[i];
				} else {
					for (uint32_t i = g.lane; i < produced; i += G)
						dst[i] = sout[i];
				}
			}
		} else {
			rc = decode_core<G, false>(g, src, ilen, dst, cap, &produced);
		}
	}
	if (g.lane == 0) {
		a.status[blk] = rc;
		a.out_len[blk] = rc == E_OK ? produced : 0u;
	}
}

template <int G>
__global__ void __launch_bounds__(1024) decompress_kernel(const DecompressParams p)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const Group<G> g;
	const uint32_t groups_per_cta = p.groups;
	const uint32_t gid = threadIdx.x / G;
	if (gid >= p.groups)
		return;	 // padding lanes of the last warp (no block-wide barriers in this kernel)
	uint8_t *sin = smem + (size_t)gid * p.group_smem;
	uint8_t *sout = sin + p.in_area;

	if (p.counter) {
		for (;;) {
			uint32_t blk = 0;
			if (g.lane == 0)
				blk = atomicAdd(p.counter, 1u);
			blk = g.bcast(blk, 0);
			if (blk >= p.a.n_blocks)
				break;
			decompress_block<G>(g, p, blk, sin, sout);
		}
	} else {
		const uint32_t total = gridDim.x * groups_per_cta;
		for (uint32_t blk = blockIdx.x * groups_per_cta + gid; blk < p.a.n_blocks; blk += total)
			decompress_block<G>(g, p, blk, sin, sout);
	}
}

}  // namespace csb

using namespace csb;

template <int G>
static int launch_decompress_g(const DecompressParams &p, int threads, int ctas, size_t smem, cudaStream_t s)
{
	cudaError_t e = cudaFuncSetAttribute(decompress_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess)
		return (int)e;
	decompress_kernel<G><<<ctas, threads, smem, s>>>(p);
	count_launch();
	return (int)cudaGetLastError();
}

extern "C" int csb_launch_decompress(const struct csb_decompress_args *a, csb_stream_t s)
{
	if (a->n_blocks == 0)
		return 0;
	DeviceInfo di;
	int e = device_info(&di);
	if (e)
		return e;

	DecompressParams p;
	p.a = *a;
	// staging capacities: output from the (uniform) capacity or the stride, input from the hint,
	// the stride, or the format's worst case for that output size; both capped at fragment scale
	uint64_t out_cap = a->out_cap ? a->out_stride : a->uniform_cap;
	if (out_cap > CSB_FRAGMENT_MAX)
		out_cap = CSB_FRAGMENT_MAX;
	uint64_t in_cap = a->max_in_len;
	if (!in_cap) {
		in_cap = 32 + out_cap + out_cap / 6;
		if (!a->in_off && a->in_stride && a->in_stride < in_cap)
			in_cap = a->in_stride;
	}
	if (in_cap > 32 + CSB_FRAGMENT_MAX + CSB_FRAGMENT_MAX / 6)
		in_cap = 32 + CSB_FRAGMENT_MAX + CSB_FRAGMENT_MAX / 6;
	p.in_area = (uint32_t)((in_cap + 15) & ~15ull);
	p.out_area = (uint32_t)((out_cap + 15) & ~15ull);
	p.group_smem = p.in_area + p.out_area;
	if (p.group_smem == 0)
		p.group_smem = 16;

	const int G = a->lanes ? a->lanes : 32;
	const int ctas_per_sm = a->ctas_per_sm > 0 ? a->ctas_per_sm : 1;
	long budget = (long)di.smem_per_sm / ctas_per_sm - 1024;
	if (budget > di.smem_per_block_optin)
		budget = di.smem_per_block_optin;
	int groups = (int)(budget / p.group_smem);
	const int max_groups = 1024 / G;
	if (groups > max_groups)
		groups = max_groups;
	if (groups < 1) {
		// cannot stage even one block: run streaming only
		p.in_area = p.out_area = 0;
		p.group_smem = 16;
		groups = 256 / G;
	}
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
	case 32: e = launch_decompress_g<32>(p, threads, (int)ctas, smem, s); break;
	case 16: e = launch_decompress_g<16>(p, threads, (int)ctas, smem, s); break;
	case 8: e = launch_decompress_g<8>(p, threads, (int)ctas, smem, s); break;
	default: e = (int)cudaErrorInvalidValue; break;
	}
	if (counter)
		cudaFreeAsync(counter, s);
	return e;
}
