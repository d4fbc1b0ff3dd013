/*
 * kernels.h -- C-linkage launchers of the sm_100a kernels, called by the plain-C
 * shim (csnappy_shim.c).  Internal to the library; the public ABI is in include/.
 */
#ifndef CSNAPPY_B200_KERNELS_H_
#define CSNAPPY_B200_KERNELS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *csb_stream_t; /* == cudaStream_t */

#define CSB_FRAGMENT_MAX 32768u
#define CSB_LEN_REFUSED 0xffffffffu /* == CSNAPPY_BATCH_LEN_REFUSED */

struct csb_compress_args {
	const uint8_t *in;
	const uint64_t *in_off; /* NULL => i * in_stride */
	uint64_t in_stride;
	const uint32_t *in_len; /* NULL => uniform_len (clipped by total_len if nonzero); an entry above 32768 or above
				   the stride is refused: out_len[i] = CSB_LEN_REFUSED, nothing written */
	uint32_t uniform_len;
	uint64_t total_len;	/* nonzero: block i has min(uniform_len, total_len - i*in_stride) bytes */
	uint32_t n_blocks;
	uint8_t *out;
	uint64_t out_stride;
	uint32_t *out_len;
	int wm;			/* log2 of hash-table bytes, 9..16 */
	uint32_t flags;		/* CSNAPPY_BATCH_SHRINK_TABLE */
	int lanes;		/* lanes cooperating on one block: 8, 16, 32 (0 = default) */
	int ctas_per_sm;	/* 0 = default */
	uint32_t *counter;	/* device word for the block claim counter (NULL: a word of the per-device ring, pack_kernel.cu next_counter) */
	int stage_input;	/* 0: choose; 1: always stage the block in shared memory; 2: read it from global memory (32-lane groups) */
};

struct csb_decompress_args {
	const uint8_t *in;
	const uint64_t *in_off;
	uint64_t in_stride;
	const uint32_t *in_len;
	uint32_t n_blocks;
	uint8_t *out;
	uint64_t out_stride;
	const uint32_t *out_cap; /* NULL => uniform_cap */
	uint32_t uniform_cap;
	uint32_t *out_len;
	int32_t *status;
	uint32_t flags;		/* CSNAPPY_BATCH_WITH_HEADER | CSNAPPY_BATCH_RAW_IF_FULL */
	uint32_t max_in_len;	/* staging hint: longest input block (0 = derive) */
	int lanes;
	int ctas_per_sm;
	uint32_t *counter;	/* device word for the block claim counter (NULL: a word of the per-device ring, pack_kernel.cu next_counter) */
	int stage_input;	/* 0: choose; 1: always stage in shared memory; 2: stage only the output; 3: warp per block against
				   global memory; 4: one lane per block (decompress_lane_kernel.cu) */
	int smem_kb;		/* unstaged mode: shared memory to use per SM, rest stays L1 (0 = default) */
	int lane_warps;		/* lane-per-block decoder: warps (of 32 blocks) in flight per SM (0 = default) */
};

/* all return 0 or a cudaError_t value (> 0) */
int csb_launch_compress(const struct csb_compress_args *a, csb_stream_t s);
int csb_launch_decompress(const struct csb_decompress_args *a, csb_stream_t s);
int csb_launch_decompress_lane(const struct csb_decompress_args *a, csb_stream_t s); /* one lane per block; out and out_stride 16-byte aligned */
int csb_launch_pack(const uint8_t *slots, uint64_t slot_stride, const uint32_t *len,
		    uint32_t n_blocks, uint8_t *packed, uint64_t *off, csb_stream_t s);
int csb_launch_pack_stored(const uint8_t *slots, uint64_t slot_stride, const uint32_t *len, uint32_t n_blocks,
			   const uint8_t *in, uint32_t page_len, uint64_t total_in, uint32_t *clen,
			   uint8_t *packed, uint64_t *off, csb_stream_t s);
int csb_launch_scan(const uint32_t *len, uint32_t n_blocks, uint64_t *off, csb_stream_t s);
/* batched csnappy_compress framing: header + fragments of every buffer into its output slot (pack_kernel.cu) */
int csb_launch_frame(const uint8_t *slots, uint64_t slot_stride, const uint32_t *len, const uint64_t *off,
		     const uint32_t *fbuf, const uint32_t *bfirst, const uint32_t *blen, uint32_t n_frag,
		     uint32_t n_buffers, uint8_t *out, uint64_t out_stride, uint32_t *out_len, csb_stream_t s);
/* Parallel decoder of ONE long raw stream (stream_kernel.cu).  aux tables: which = 0 -> per input byte, 1 -> per output
 * byte.  Returns 0, 1 ("not for this stream": caller falls back to the warp-per-stream path), or a cudaError_t (> 1). */
size_t csb_stream_aux_bytes(uint32_t src_len, uint32_t cap, int which);
int csb_launch_decompress_stream(const uint8_t *d_in, uint32_t src_len, uint8_t *d_out, uint32_t cap, uint32_t *d_out_len,
				 int32_t *d_status, void *d_aux_in, void *d_aux_out, csb_stream_t s);
uint64_t csb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
