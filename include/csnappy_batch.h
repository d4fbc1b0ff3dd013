/*
 * csnappy_batch.h -- batched entry points of the B200-native Snappy codec.
 *
 * These are the additions BASELINE.json's north_star asks for next to the
 * csnappy.h drop-in: many independent blocks (zram 4 KiB pages,
 * kernel_3_2_10.patch:1348-1372; block_compressor pages, block_compressor.c:
 * 113-134, 307-335, 365-387; the 32 KiB fragments of csnappy_compress,
 * csnappy_compress.c:636-654) handed to the GPU in ONE call.  Per block the
 * semantics are exactly those of csnappy_compress_fragment /
 * csnappy_decompress_noheader / csnappy_decompress.
 *
 * Pointers named d_* are DEVICE pointers on the current CUDA device, pointers
 * named h_* are HOST pointers.  `stream` is a cudaStream_t passed as void*
 * (NULL = the legacy default stream).  Device-pointer calls are asynchronous
 * with respect to the host: they enqueue work on `stream` and return.
 *
 * Block i of a batch lives at
 *     base + (off ? off[i] : i * stride)
 * Strided outputs need  out_stride >= csnappy_max_compressed_length(len)  for
 * compression (the compressor never bounds-checks its output, exactly like the
 * reference, cl_tester.c:120-165) and >= capacity for decompression.
 * Fast paths engage when base, stride and offsets are multiples of 16.
 *
 * Every function returns 0, CSNAPPY_E_DEVICE or CSNAPPY_E_BAD_ARG; per-block
 * decoder results go to d_status[] / h_status[].
 */
#ifndef CSNAPPY_B200_CSNAPPY_BATCH_H_
#define CSNAPPY_B200_CSNAPPY_BATCH_H_

#include <stddef.h>
#include <stdint.h>

#include "csnappy.h"

#ifdef __cplusplus
extern "C" {
#endif

/* flags */
#define CSNAPPY_BATCH_SHRINK_TABLE 0x1u /* compress: per block, apply csnappy_compress()'s
					   short-chunk table rule (csnappy_compress.c:638-646) */
#define CSNAPPY_BATCH_WITH_HEADER 0x2u  /* decompress: blocks start with a varint32 length;
					   csnappy_decompress() semantics incl. -1 / -2
					   (csnappy_decompress.c:394-411) */
#define CSNAPPY_BATCH_RAW_IF_FULL 0x4u  /* decompress: a block whose input length equals its
					   output capacity is a stored block and is copied
					   (block_compressor.c:378) */

/*
 * Batched csnappy_compress_fragment (csnappy_compress.c:469-606).
 *   d_in_len == NULL  => every block is uniform_in_len bytes.
 *   d_out_len[i]      <- compressed size of block i.
 * Block lengths must be <= 32768 (and <= in_stride when strided); 9 <= wm <= 16.
 * The reference has no error channel on this side (csnappy.h:38 only REQUIRES it); here a
 * d_in_len[i] that breaks the rule is refused: d_out_len[i] = CSNAPPY_BATCH_LEN_REFUSED and
 * nothing is written to the block's slot.
 */
#define CSNAPPY_BATCH_LEN_REFUSED 0xffffffffu
int csnappy_batch_compress_fragments(const void *d_in, const uint64_t *d_in_off,
				     uint64_t in_stride, const uint32_t *d_in_len,
				     uint32_t uniform_in_len, uint32_t n_blocks,
				     void *d_out, uint64_t out_stride,
				     uint32_t *d_out_len,
				     int workmem_bytes_power_of_two, uint32_t flags,
				     void *stream);

/*
 * Batched csnappy_decompress_noheader (csnappy_decompress.c:319-387), or
 * csnappy_decompress with CSNAPPY_BATCH_WITH_HEADER.
 *   d_out_cap == NULL => every block has capacity uniform_out_cap.
 *   d_status[i]       <- 0 / -1 / -2 / -3 / -5
 *   d_out_len[i]      <- bytes produced when d_status[i] == 0, else 0.
 */
int csnappy_batch_decompress(const void *d_in, const uint64_t *d_in_off,
			     uint64_t in_stride, const uint32_t *d_in_len,
			     uint32_t n_blocks, void *d_out, uint64_t out_stride,
			     const uint32_t *d_out_cap, uint32_t uniform_out_cap,
			     uint32_t *d_out_len, int32_t *d_status, uint32_t flags,
			     void *stream);

/*
 * Batched csnappy_compress (csnappy_compress.c:621-656): n_buffers WHOLE buffers, each framed exactly like one
 * csnappy_compress() call -- varint32 length, one fragment per 32 KiB chunk, the short last chunk with the
 * reference's smaller-table rule (:638-646) -- in one call: all fragments of all buffers are compressed by one
 * launch, then headers and fragments are packed into each buffer's output slot.
 *   buffer i        d_in + (h_in_off ? h_in_off[i] : i * in_stride), h_in_len[i] bytes (NULL: uniform_in_len)
 *   output i        d_out + i * out_stride, d_out_len[i] <- its length;
 *                   out_stride >= 5 + n + n/6 + 32 * ceil(n / 32768) for the longest buffer
 * The lengths (and offsets) are HOST arrays: the fragment table is laid out on the host.  d_workspace must hold
 * csnappy_batch_compress_workspace(...) bytes (fragment slots, sizes, tables) and stay untouched until the
 * work on `stream` has finished.
 */
uint64_t csnappy_batch_compress_workspace(const uint32_t *h_in_len, uint32_t uniform_in_len, uint32_t n_buffers);
int csnappy_batch_compress(const void *d_in, const uint64_t *h_in_off, uint64_t in_stride,
			   const uint32_t *h_in_len, uint32_t uniform_in_len, uint32_t n_buffers,
			   void *d_out, uint64_t out_stride, uint32_t *d_out_len,
			   int workmem_bytes_power_of_two, void *d_workspace,
			   uint64_t workspace_bytes, void *stream);

/*
 * ONE long raw stream, device resident (csnappy_decompress_noheader semantics, csnappy_decompress.c:319-387):
 * tag starts by speculative parsing + pointer jumping, back-references by pointer jumping over the output
 * (stream_kernel.cu) instead of a serial tag walk.  This is what the host-pointer csnappy_decompress /
 * csnappy_decompress_noheader use for streams above 64 KiB.  d_workspace: csnappy_stream_decompress_workspace()
 * bytes (8 per input byte + 4 per output byte).  *d_status <- 0 / -3 / -5, *d_out_len <- bytes produced (0 on error).
 */
uint64_t csnappy_stream_decompress_workspace(uint32_t src_len, uint32_t out_cap);
int csnappy_stream_decompress(const void *d_src, uint32_t src_len, void *d_dst, uint32_t out_cap,
			      uint32_t *d_out_len, int32_t *d_status, void *d_workspace,
			      uint64_t workspace_bytes, void *stream);

/*
 * Exclusive scan of d_len[0..n) into d_off[0..n] (d_off[n] = total) and
 * gather of the strided slots into one contiguous payload:
 *     d_packed[d_off[i] .. d_off[i]+d_len[i]) = d_slots[i*slot_stride ..)
 * This is the size-gather + prefix-sum index of block_compressor.c:298-335 and
 * the fragment concatenation of csnappy_compress.c:647-653, done on the device.
 * d_packed may be NULL to compute offsets only.
 */
int csnappy_batch_pack(const void *d_slots, uint64_t slot_stride,
		       const uint32_t *d_len, uint32_t n_blocks, void *d_packed,
		       uint64_t *d_off, void *stream);

/*
 * Host-buffer variants: same semantics, HOST pointers, synchronous.  Pages are
 * streamed through the device in chunks on several CUDA streams so that H2D,
 * kernels and D2H overlap.  These are what a zram / block_compressor style
 * caller with host memory uses, and what bench.py times as "e2e".
 */
int csnappy_batch_compress_fragments_host(const void *h_in, uint64_t in_stride,
					  uint32_t uniform_in_len, uint32_t n_blocks,
					  void *h_out, uint64_t out_stride,
					  uint32_t *h_out_len,
					  int workmem_bytes_power_of_two);

int csnappy_batch_decompress_host(const void *h_in, uint64_t in_stride,
				  const uint32_t *h_in_len, uint32_t n_blocks,
				  void *h_out, uint64_t out_stride,
				  uint32_t uniform_out_cap, uint32_t *h_out_len,
				  int32_t *h_status, uint32_t flags);

/*
 * block_compressor-style page container on HOST buffers (reference block_compressor.c:275-394):
 *     [u32 nr_pages][u32 clen[nr_pages]][payload_0]...[payload_{nr_pages-1}]
 * payload_i = csnappy_compress_fragment(page_i, wm), or the page itself when that is not smaller
 * (clen_i = page length, :316-318); a reader takes clen_i == page_size as "stored" (:378) -- the
 * reference's wart that a stored PARTIAL last page is not recognised is reproduced, not fixed.
 * The bytes written are identical to what block_compressor's writer loop produces with csnappy
 * and WMSIZE_ORDER = workmem_bytes_power_of_two (:99).  Chunks of pages are pipelined through the
 * device (H2D / kernels / D2H overlap); only compressed bytes cross the bus on the compressed side.
 *
 * csnappy_bc_compress_host    container_capacity >= csnappy_bc_max_container_length(); returns 0,
 *                             CSNAPPY_E_DEVICE or CSNAPPY_E_BAD_ARG.
 * csnappy_bc_decompress_host  out_capacity >= nr_pages * page_size (CSNAPPY_E_OUTPUT_INSUF else);
 *                             returns 0 or the code of the first failing page (its index goes to
 *                             *failed_page when given); *out_length = bytes produced.
 */
uint64_t csnappy_bc_max_container_length(uint64_t input_length, uint32_t page_size);
int csnappy_bc_compress_host(const void *h_in, uint64_t input_length, uint32_t page_size,
			     void *h_container, uint64_t container_capacity,
			     uint64_t *container_length, int workmem_bytes_power_of_two);
int csnappy_bc_decompress_host(const void *h_container, uint64_t container_length,
			       uint32_t page_size, void *h_out, uint64_t out_capacity,
			       uint64_t *out_length, uint32_t *failed_page);

/*
 * The same container over SEVERAL devices from one process (static sharding, per-device streams,
 * the host gathers the u32 size arrays and prefix-sums them into one container): the page range is cut into chunks,
 * chunk c is handled by devices[c mod n_devices] on that device's own worker thread and streams; there is no
 * data-path collective -- the only thing that crosses devices is the running payload position (8 bytes per chunk).
 * devices == NULL: devices 0 .. n_devices-1; n_devices == 0: every visible device.  The bytes produced are
 * identical to the single-device call.  The caller's current device is not changed.
 */
int csnappy_bc_compress_host_multi(const void *h_in, uint64_t input_length, uint32_t page_size,
				   void *h_container, uint64_t container_capacity,
				   uint64_t *container_length, int workmem_bytes_power_of_two,
				   const int *devices, int n_devices);
int csnappy_bc_decompress_host_multi(const void *h_container, uint64_t container_length,
				     uint32_t page_size, void *h_out, uint64_t out_capacity,
				     uint64_t *out_length, uint32_t *failed_page,
				     const int *devices, int n_devices);

/*
 * Library / device introspection and kernel tuning knobs (used by bench.py and
 * the tests; not needed by drop-in callers).
 */
int csnappy_b200_device_ok(void);	   /* 1 if a CUDA device is usable */
int csnappy_b200_device_count(void);	   /* visible CUDA devices (0 on error) */
const char *csnappy_b200_last_error(void); /* text of the last device error (thread local) */
uint64_t csnappy_b200_kernel_launches(void); /* kernels launched by this library so far */
/* key (value 0 restores the default; returns 0 or CSNAPPY_E_BAD_ARG):
 *   "compress_lanes" | "decompress_lanes"  lanes cooperating on one block in the warp-per-block kernels: 8 / 16 / 32
 *   "compress_stage_input"    1: always stage blocks in shared memory; 2: read them from global memory (32-lane
 *                             groups only; the default for large batches of fragments above ~14 KB)
 *   "decompress_stage_input"  1: always stage blocks in shared memory; 2: stage only the output and read the
 *                             compressed block through L1; 3: decode against global memory, a warp per block;
 *                             4: one lane per block (the default for very large batches)
 *   "decompress_lane_warps"   lane-per-block decoder: warps of 32 blocks in flight per SM (default 32)
 *   "decompress_smem_kb"      shared memory per SM in mode 2
 *   "ctas_per_sm"             resident CTAs per SM of the warp-per-block kernels
 *   "stream_decode_min"       bytes from which ONE single stream takes the parallel stream decoder; -1: never
 *   "copy_threads"            helper threads that stage PAGEABLE caller memory through pinned buffers in the
 *                             host-buffer pipelines (read when the pool starts; default 7 on >= 16 cores)
 *   "chunk_mb"                MiB per chunk of the host-buffer pipelines (default 32; measured flat from 32 to 128)
 *   "no_bounce"               1: hand pageable caller memory straight to cudaMemcpyAsync (round-1 behaviour)
 *   "host_register"           1: page-lock the caller's buffers for the duration of a host-buffer call (measured
 *                             slower than the staging above unless the same buffers are registered once by the caller) */
int csnappy_b200_set_tuning(const char *key, int value);

#ifdef __cplusplus
}
#endif
#endif /* CSNAPPY_B200_CSNAPPY_BATCH_H_ */
